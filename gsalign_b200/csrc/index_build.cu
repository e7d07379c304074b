// index_build.cu -- GPU construction of the BWA-format index files (.pac .ann .amb .bwt .sa).
//
// Replaces bwa_idx_build (reference src/BWT_Index/bwtindex.c:77-149): bns_fasta2bntseq (bntseq.c:158-211),
// the BWT-SW incremental BWT construction (bwt_gen.c), bwt_bwtupdate_core (bwtindex.c:53-75) and
// bwt_cal_sa (bwt.c:101-123).  The BWT, Occ table and sampled SA are canonical functions of the text
// T = F . revcomp(F), so any suffix sorter yields byte-identical files provided the conventions are kept:
//   * non-ACGT bases become lrand48() & 3 after srand48(11), consumed in FASTA order (bntseq.c:144,173-174);
//   * rows of the sorted suffix matrix of T$ with row 0 = "$"; primary = the row of suffix 0;
//   * .bwt = primary, L2[1..4], then per 128 symbols of the '$'-less BWT 4 x u64 running counts + 8 words,
//     plus one trailing count block; .sa = primary, L2[1..4], 32, |T|, then SA[32k] for k >= 1.
// Suffix sorting is prefix doubling on the device: one radix sort on (16-mer, length) keys, then rounds
// that re-sort only the still-tied suffixes by (rank[i], rank[i+h]) with h = 16, 32, 64, ...
// This is SURVEY.md's "next" row N1; it is not on the timed path (the index is an offline artefact), but it
// makes `GSAlign -r`, `GSAlign index` and the benchmark self-contained at 100 Mbp+ (reference: ~1.1 s/Mbp).
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <algorithm>
#include <string>
#include <vector>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include "gsa_internal.cuh"

#define IB_CHECK(call)                                                                                     \
	do {                                                                                                   \
		cudaError_t _e = (call);                                                                           \
		if (_e != cudaSuccess) { fprintf(stderr, "[gsa_index] %s:%d %s: %s\n", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); return -2; } \
	} while (0)

// ---------------------------------------------------------------------------------------------------
// FASTA -> forward pac + annotations (host; kseq/bntseq semantics)
// ---------------------------------------------------------------------------------------------------
struct Hole { int64_t offset; int32_t len; char amb; };
struct SeqAnn { std::string name, anno; int64_t offset; int32_t len, n_ambs; };

static inline int nt4_host(unsigned char ch)
{
	switch (ch) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': return 3;
	default: return 4;
	}
}

static int pack_fasta(const char *path, std::vector<uint8_t> &pac, int64_t &l_pac, std::vector<SeqAnn> &anns, std::vector<Hole> &holes)
{
	gzFile fp = gzopen(path, "r");
	if (!fp) { fprintf(stderr, "[gsa_index] cannot open %s\n", path); return -1; }
	std::vector<char> buf;
	{
		char tmp[1 << 16]; int got;
		while ((got = gzread(fp, tmp, sizeof(tmp))) > 0) buf.insert(buf.end(), tmp, tmp + got);
		gzclose(fp);
	}
	srand48(11); // bns->seed, bntseq.c:173-174
	l_pac = 0; pac.clear(); anns.clear(); holes.clear();
	size_t p = 0, n = buf.size();
	long qi = -1; // the hole being extended (index: the vector may reallocate)
	while (p < n && buf[p] != '>' && buf[p] != '@') p++; // kseq_read: jump to the first header
	while (p < n) {
		p++; // header char
		SeqAnn a; a.n_ambs = 0;
		size_t s = p;
		while (p < n && !isspace((unsigned char)buf[p])) p++;
		a.name.assign(buf.data() + s, p - s);
		if (p < n && buf[p] != '\n') { // comment = rest of the line
			s = ++p;
			while (p < n && buf[p] != '\n') p++;
			size_t e = p;
			if (e - s > 1 && buf[e - 1] == '\r') e--;
			a.anno.assign(buf.data() + s, e - s);
		}
		if (p < n) p++; // the newline
		if (a.anno.empty()) a.anno = "(null)"; // bntseq.c:120
		a.offset = l_pac; a.len = 0;
		int lasts = 0;
		while (p < n && buf[p] != '>' && buf[p] != '+' && buf[p] != '@') { // one sequence line
			if (buf[p] == '\n') { p++; continue; }
			size_t ls = p;
			while (p < n && buf[p] != '\n') p++;
			size_t le = p;
			if (le - ls > 0 && buf[le - 1] == '\r' && (a.len + (le - ls)) > 1) le--;
			for (size_t k = ls; k < le; k++) {
				unsigned char ch = (unsigned char)buf[k];
				int c = nt4_host(ch);
				if (c >= 4) { // bntseq.c:124-141: one hole per run of the SAME character
					if (lasts == ch && qi >= 0) holes[(size_t)qi].len++;
					else { Hole h; h.offset = a.offset + a.len; h.len = 1; h.amb = (char)ch; holes.push_back(h); qi = (long)holes.size() - 1; a.n_ambs++; }
					c = (int)(lrand48() & 3);
				}
				lasts = ch;
				if ((l_pac & 3) == 0) pac.push_back(0);
				pac[(size_t)(l_pac >> 2)] |= (uint8_t)(c << ((~l_pac & 3) << 1));
				l_pac++; a.len++;
			}
			if (p < n) p++;
		}
		anns.push_back(a);
		if (p < n && buf[p] == '+') { fprintf(stderr, "[gsa_index] FASTQ input is not supported\n"); return -1; }
	}
	if (anns.empty() || l_pac == 0) { fprintf(stderr, "[gsa_index] no sequence in %s\n", path); return -1; }
	return 0;
}

static int write_pac_ann_amb(const std::string &prefix, const std::vector<uint8_t> &pac, int64_t l_pac, const std::vector<SeqAnn> &anns, const std::vector<Hole> &holes)
{
	FILE *fp = fopen((prefix + ".pac").c_str(), "wb");
	if (!fp) return -1;
	fwrite(pac.data(), 1, (size_t)((l_pac >> 2) + ((l_pac & 3) == 0 ? 0 : 1)), fp); // bntseq.c:192-201
	unsigned char ct;
	if (l_pac % 4 == 0) { ct = 0; fwrite(&ct, 1, 1, fp); }
	ct = (unsigned char)(l_pac % 4); fwrite(&ct, 1, 1, fp);
	fclose(fp);
	fp = fopen((prefix + ".ann").c_str(), "w"); // bns_dump, bntseq.c:59-89
	if (!fp) return -1;
	fprintf(fp, "%lld %d %u\n", (long long)l_pac, (int)anns.size(), 11u);
	for (const SeqAnn &a : anns) {
		fprintf(fp, "%d %s", 0, a.name.c_str());
		if (!a.anno.empty()) fprintf(fp, " %s\n", a.anno.c_str()); else fprintf(fp, "\n");
		fprintf(fp, "%lld %d %d\n", (long long)a.offset, a.len, a.n_ambs);
	}
	fclose(fp);
	fp = fopen((prefix + ".amb").c_str(), "w");
	if (!fp) return -1;
	fprintf(fp, "%lld %d %u\n", (long long)l_pac, (int)anns.size(), (unsigned)holes.size());
	for (const Hole &h : holes) fprintf(fp, "%lld %d %c\n", (long long)h.offset, h.len, h.amb);
	fclose(fp);
	return 0;
}

// ---------------------------------------------------------------------------------------------------
// device kernels
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ib_base(const uint32_t *txt, uint32_t i) { return (int)(txt[i >> 4] >> ((~i & 15) << 1)) & 3; }

__global__ void k_ib_text(const uint8_t *pac, int64_t N, uint32_t *txt, uint64_t nwords)
{
	uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= nwords) return;
	uint32_t v = 0;
	for (int i = 0; i < 16; i++) {
		int64_t pos = (int64_t)(w << 4) + i;
		int s = 0;
		if (pos < N) s = pac[pos >> 2] >> ((~pos & 3) << 1) & 3;
		else if (pos < 2 * N) { int64_t f = 2 * N - 1 - pos; s = 3 - (pac[f >> 2] >> ((~f & 3) << 1) & 3); }
		v |= (uint32_t)s << ((15 - i) << 1);
	}
	txt[w] = v;
}

// round 0 key: (16-mer padded with A, number of real bases) -- '$' sorts before A, so the shorter suffix wins a tie
__global__ void k_ib_key0(const uint32_t *txt, uint32_t n, uint64_t *key, uint32_t *val)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t w0 = txt[i >> 4], w1 = txt[(i >> 4) + 1];
	uint32_t kmer = __funnelshift_l(w1, w0, (i & 15) << 1);
	uint32_t len = n - i < 16 ? n - i : 16;
	if (len < 16) kmer &= ~0u << ((16 - len) << 1);
	key[i] = ((uint64_t)kmer << 5) | len;
	val[i] = i;
}

// rank of the suffix in sorted slot k = (slot of the head of its run of equal keys) + 1
__global__ void k_ib_heads(const uint64_t *key, const uint32_t *slot_of, uint32_t m, uint32_t *head)
{
	uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= m) return;
	uint32_t slot = slot_of ? slot_of[k] : k;
	head[k] = (k == 0 || key[k] != key[k - 1]) ? slot + 1 : 0;
}

__global__ void k_ib_set_rank(const uint32_t *sorted_sfx, const uint32_t *newrank, uint32_t m, uint32_t *rank)
{
	uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < m) rank[sorted_sfx[k]] = newrank[k];
}

__global__ void k_ib_scatter(const uint32_t *slot_of, const uint32_t *sorted_sfx, uint32_t m, uint32_t *sa)
{
	uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < m) sa[slot_of[k]] = sorted_sfx[k];
}

__global__ void k_ib_unresolved(const uint32_t *sa, const uint32_t *rank, uint32_t n, uint8_t *flag)
{
	uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	uint32_t r = rank[sa[j]];
	flag[j] = (j > 0 && rank[sa[j - 1]] == r) || (j + 1 < n && rank[sa[j + 1]] == r);
}

__global__ void k_ib_keys(const uint32_t *slots, uint32_t m, const uint32_t *sa, const uint32_t *rank, uint32_t n, uint32_t h, uint64_t *key, uint32_t *val)
{
	uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= m) return;
	uint32_t s = sa[slots[k]];
	uint64_t r2 = (uint64_t)s + h < n ? rank[s + h] : 0;
	key[k] = ((uint64_t)rank[s] << 32) | r2;
	val[k] = s;
}

struct MaxOp { __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } };

// BWT symbol x of the '$'-less BWT: rows 0..n without `primary`; row 0 is the suffix "$" (SA = n)
__device__ __forceinline__ int ib_bwt_sym(const uint32_t *txt, const uint32_t *sa, uint32_t n, uint32_t primary, uint32_t x)
{
	uint32_t row = x + (x >= primary);
	uint32_t pos = row == 0 ? n : sa[row - 1];
	return ib_base(txt, pos - 1);
}

// one thread per 128-symbol block: the 8 symbol words go straight to their interleaved place, counts to cnt[4][nblk]
__global__ void k_ib_bwt_block(const uint32_t *txt, const uint32_t *sa, uint32_t n, uint32_t primary, uint32_t *out, unsigned long long *cnt, uint32_t nblk)
{
	uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nblk) return;
	uint32_t c[4] = {0, 0, 0, 0};
	for (int w = 0; w < 8; w++) {
		uint32_t x0 = b * 128 + w * 16;
		if (x0 >= n) break;
		uint32_t v = 0;
		for (int i = 0; i < 16 && x0 + i < n; i++) { int s = ib_bwt_sym(txt, sa, n, primary, x0 + i); c[s]++; v |= (uint32_t)s << ((15 - i) << 1); }
		out[(size_t)b * 16 + 8 + w] = v;
	}
	for (int s = 0; s < 4; s++) cnt[(size_t)s * (nblk + 1) + b] = c[s];
}

__global__ void k_ib_bwt_counts(const unsigned long long *cum, uint32_t nblk, uint32_t n, uint32_t *out)
{ // cumulative counts in front of every block + the trailing block (bwtindex.c:61-71)
	uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b > nblk) return;
	size_t base = b < nblk ? (size_t)b * 16 : (size_t)nblk * 8 + (n + 15) / 16;
	unsigned long long *o = (unsigned long long *)(out + base);
	for (int s = 0; s < 4; s++) o[s] = cum[(size_t)s * (nblk + 1) + b];
}

__global__ void k_ib_samples(const uint32_t *sa, uint32_t n_sa, int intv, unsigned long long *out)
{
	uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k == 0 || k >= n_sa) return;
	out[k - 1] = sa[(size_t)k * intv - 1]; // SA[row 32k]
}

__global__ void k_ib_find_primary(const uint32_t *sa, uint32_t n, uint32_t *primary)
{
	uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j < n && sa[j] == 0) *primary = j + 1;
}

// ---------------------------------------------------------------------------------------------------
static int sort_pairs(uint64_t *k_in, uint64_t *k_out, uint32_t *v_in, uint32_t *v_out, uint32_t m, int end_bit, void **tmp, size_t *tmp_cap)
{
	size_t bytes = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, bytes, k_in, k_out, v_in, v_out, (int)m, 0, end_bit);
	if (bytes > *tmp_cap) { if (*tmp) cudaFree(*tmp); IB_CHECK(cudaMalloc(tmp, bytes)); *tmp_cap = bytes; }
	IB_CHECK(cub::DeviceRadixSort::SortPairs(*tmp, bytes, k_in, k_out, v_in, v_out, (int)m, 0, end_bit));
	return 0;
}

extern "C" int gsa_build_index_files(const char *fasta, const char *prefix_c, int device)
{
	std::string prefix = prefix_c;
	std::vector<uint8_t> pac; int64_t N = 0; std::vector<SeqAnn> anns; std::vector<Hole> holes;
	if (pack_fasta(fasta, pac, N, anns, holes) != 0) return -1;
	if (2 * (uint64_t)N >= 0xFFFFFF00ull) { fprintf(stderr, "[gsa_index] text of %lld symbols needs the 64-bit build\n", (long long)(2 * N)); return -3; }
	if (write_pac_ann_amb(prefix, pac, N, anns, holes) != 0) { fprintf(stderr, "[gsa_index] cannot write %s.*\n", prefix_c); return -1; }
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) { fprintf(stderr, "[gsa_index] no CUDA device (index construction has no CPU path in this build)\n"); return -2; }
	IB_CHECK(cudaSetDevice(device));
	const uint32_t n = (uint32_t)(2 * N);
	const uint64_t nwords = ((uint64_t)n >> 4) + 3;
	uint8_t *d_pac; uint32_t *txt, *sa, *rank, *v_in, *v_out, *head, *slots; uint64_t *k_in, *k_out; uint8_t *flag; uint32_t *d_cnt;
	pac.resize((size_t)(N / 4 + 1), 0);
	IB_CHECK(cudaMalloc(&d_pac, pac.size())); IB_CHECK(cudaMemcpy(d_pac, pac.data(), pac.size(), cudaMemcpyHostToDevice));
	IB_CHECK(cudaMalloc(&txt, nwords * 4));
	IB_CHECK(cudaMalloc(&sa, (size_t)n * 4)); IB_CHECK(cudaMalloc(&rank, (size_t)n * 4));
	IB_CHECK(cudaMalloc(&v_in, (size_t)n * 4)); IB_CHECK(cudaMalloc(&v_out, (size_t)n * 4));
	IB_CHECK(cudaMalloc(&head, (size_t)n * 4)); IB_CHECK(cudaMalloc(&slots, (size_t)n * 4));
	IB_CHECK(cudaMalloc(&k_in, (size_t)n * 8)); IB_CHECK(cudaMalloc(&k_out, (size_t)n * 8));
	IB_CHECK(cudaMalloc(&flag, (size_t)n)); IB_CHECK(cudaMalloc(&d_cnt, 64));
	void *tmp = nullptr; size_t tmp_cap = 0;
	const int T = 256;
	k_ib_text<<<(unsigned)((nwords + T - 1) / T), T>>>(d_pac, N, txt, nwords);
	// ---- round 0: sort all suffixes by their first 16 bases ---------------------------------------------------------
	k_ib_key0<<<(n + T - 1) / T, T>>>(txt, n, k_in, v_in);
	if (sort_pairs(k_in, k_out, v_in, sa, n, 37, &tmp, &tmp_cap)) return -2;
	auto assign_ranks = [&](const uint64_t *keys, const uint32_t *slot_of, const uint32_t *sfx, uint32_t m) -> int {
		k_ib_heads<<<(m + T - 1) / T, T>>>(keys, slot_of, m, head);
		size_t bytes = 0;
		cub::DeviceScan::InclusiveScan(nullptr, bytes, head, v_in, MaxOp(), (int)m);
		if (bytes > tmp_cap) { if (tmp) cudaFree(tmp); IB_CHECK(cudaMalloc(&tmp, bytes)); tmp_cap = bytes; }
		IB_CHECK(cub::DeviceScan::InclusiveScan(tmp, bytes, head, v_in, MaxOp(), (int)m));
		k_ib_set_rank<<<(m + T - 1) / T, T>>>(sfx, v_in, m, rank);
		return 0;
	};
	if (assign_ranks(k_out, nullptr, sa, n)) return -2;
	// ---- doubling rounds over the still-tied suffixes ---------------------------------------------------------------
	int rounds = 0;
	for (uint64_t h = 16; h < (uint64_t)n * 2; h <<= 1, rounds++) {
		k_ib_unresolved<<<(n + T - 1) / T, T>>>(sa, rank, n, flag);
		size_t bytes = 0;
		thrust::counting_iterator<uint32_t> it(0);
		cub::DeviceSelect::Flagged(nullptr, bytes, it, flag, slots, d_cnt, (int)n);
		if (bytes > tmp_cap) { if (tmp) cudaFree(tmp); IB_CHECK(cudaMalloc(&tmp, bytes)); tmp_cap = bytes; }
		IB_CHECK(cub::DeviceSelect::Flagged(tmp, bytes, it, flag, slots, d_cnt, (int)n));
		uint32_t m = 0;
		IB_CHECK(cudaMemcpy(&m, d_cnt, 4, cudaMemcpyDeviceToHost));
		if (m == 0) break;
		k_ib_keys<<<(m + T - 1) / T, T>>>(slots, m, sa, rank, n, (uint32_t)std::min<uint64_t>(h, n), k_in, v_in);
		if (sort_pairs(k_in, k_out, v_in, v_out, m, 64, &tmp, &tmp_cap)) return -2;
		k_ib_scatter<<<(m + T - 1) / T, T>>>(slots, v_out, m, sa);
		if (assign_ranks(k_out, slots, v_out, m)) return -2;
	}
	IB_CHECK(cudaDeviceSynchronize());
	// ---- BWT + Occ in the BWA layout, sampled SA ------------------------------------------------------------------------
	uint32_t primary = 0;
	IB_CHECK(cudaMemset(d_cnt, 0, 4));
	k_ib_find_primary<<<(n + T - 1) / T, T>>>(sa, n, d_cnt);
	IB_CHECK(cudaMemcpy(&primary, d_cnt, 4, cudaMemcpyDeviceToHost));
	const uint32_t nblk = (n + 127) / 128;
	const size_t bwt_words = (size_t)nblk * 8 + ((size_t)n + 15) / 16 + 8;
	uint32_t *d_bwt; unsigned long long *d_c, *d_cum;
	IB_CHECK(cudaMalloc(&d_bwt, bwt_words * 4)); IB_CHECK(cudaMemset(d_bwt, 0, bwt_words * 4));
	IB_CHECK(cudaMalloc(&d_c, (size_t)4 * (nblk + 1) * 8)); IB_CHECK(cudaMalloc(&d_cum, (size_t)4 * (nblk + 1) * 8));
	IB_CHECK(cudaMemset(d_c, 0, (size_t)4 * (nblk + 1) * 8));
	k_ib_bwt_block<<<(nblk + 127) / 128, 128>>>(txt, sa, n, primary, d_bwt, d_c, nblk);
	for (int s = 0; s < 4; s++) {
		size_t bytes = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_c + (size_t)s * (nblk + 1), d_cum + (size_t)s * (nblk + 1), (int)(nblk + 1));
		if (bytes > tmp_cap) { if (tmp) cudaFree(tmp); IB_CHECK(cudaMalloc(&tmp, bytes)); tmp_cap = bytes; }
		IB_CHECK(cub::DeviceScan::ExclusiveSum(tmp, bytes, d_c + (size_t)s * (nblk + 1), d_cum + (size_t)s * (nblk + 1), (int)(nblk + 1)));
	}
	k_ib_bwt_counts<<<(nblk + 1 + T - 1) / T, T>>>(d_cum, nblk, n, d_bwt);
	const int intv = 32;
	const uint32_t n_sa = (uint32_t)(((uint64_t)n + intv) / intv);
	unsigned long long *d_samples;
	IB_CHECK(cudaMalloc(&d_samples, (size_t)n_sa * 8));
	k_ib_samples<<<(n_sa + T - 1) / T, T>>>(sa, n_sa, intv, d_samples);
	std::vector<uint32_t> h_bwt(bwt_words);
	std::vector<unsigned long long> h_samples(n_sa > 1 ? n_sa - 1 : 0), h_tot(4);
	IB_CHECK(cudaMemcpy(h_bwt.data(), d_bwt, bwt_words * 4, cudaMemcpyDeviceToHost));
	if (n_sa > 1) IB_CHECK(cudaMemcpy(h_samples.data(), d_samples, (size_t)(n_sa - 1) * 8, cudaMemcpyDeviceToHost));
	for (int s = 0; s < 4; s++) IB_CHECK(cudaMemcpy(&h_tot[s], d_cum + (size_t)s * (nblk + 1) + nblk, 8, cudaMemcpyDeviceToHost));
	unsigned long long L2[5] = {0, 0, 0, 0, 0};
	for (int s = 0; s < 4; s++) L2[s + 1] = L2[s] + h_tot[s];
	unsigned long long prim64 = primary, seq_len = n, intv64 = intv;
	FILE *fp = fopen((prefix + ".bwt").c_str(), "wb");
	if (!fp) return -1;
	fwrite(&prim64, 8, 1, fp); fwrite(L2 + 1, 8, 4, fp); fwrite(h_bwt.data(), 4, bwt_words, fp);
	fclose(fp);
	fp = fopen((prefix + ".sa").c_str(), "wb");
	if (!fp) return -1;
	fwrite(&prim64, 8, 1, fp); fwrite(L2 + 1, 8, 4, fp); fwrite(&intv64, 8, 1, fp); fwrite(&seq_len, 8, 1, fp);
	fwrite(h_samples.data(), 8, h_samples.size(), fp);
	fclose(fp);
	cudaFree(d_pac); cudaFree(txt); cudaFree(sa); cudaFree(rank); cudaFree(v_in); cudaFree(v_out); cudaFree(head); cudaFree(slots);
	cudaFree(k_in); cudaFree(k_out); cudaFree(flag); cudaFree(d_cnt); cudaFree(d_bwt); cudaFree(d_c); cudaFree(d_cum); cudaFree(d_samples);
	if (tmp) cudaFree(tmp);
	fprintf(stderr, "[gsa_index] %lld bp, %d sequence(s), %d doubling round(s) -> %s.{pac,ann,amb,bwt,sa}\n", (long long)N, (int)anns.size(), rounds, prefix_c);
	return 0;
}
