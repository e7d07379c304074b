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
#include <sys/stat.h>
#include <algorithm>
#include <string>
#include <vector>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/device/device_segmented_sort.cuh>
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
		struct stat sb;
		size_t cap = stat(path, &sb) == 0 && sb.st_size > 0 ? (size_t)sb.st_size + (1 << 20) : (size_t)1 << 20; // exact for plain files, a start for .gz
		buf.resize(cap);
		size_t used = 0; int got;
		gzbuffer(fp, 1 << 20);
		for (;;) {
			if (used == buf.size()) buf.resize(buf.size() * 2);
			size_t want = std::min<size_t>(buf.size() - used, (size_t)1 << 30);
			if ((got = gzread(fp, buf.data() + used, (unsigned)want)) <= 0) break;
			used += (size_t)got;
		}
		gzclose(fp);
		buf.resize(used);
	}
	srand48(11); // bns->seed, bntseq.c:173-174
	l_pac = 0; anns.clear(); holes.clear();
	pac.assign(buf.size() / 4 + 2, 0); // upper bound: every byte of the file a base
	size_t p = 0, n = buf.size();
	static uint8_t nt4_tab[256];
	for (int c = 0; c < 256; c++) nt4_tab[c] = (uint8_t)nt4_host((unsigned char)c);
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
				int c = nt4_tab[ch];
				if (c >= 4) { // bntseq.c:124-141: one hole per run of the SAME character
					if (lasts == ch && qi >= 0) holes[(size_t)qi].len++;
					else { Hole h; h.offset = a.offset + a.len; h.len = 1; h.amb = (char)ch; holes.push_back(h); qi = (long)holes.size() - 1; a.n_ambs++; }
					c = (int)(lrand48() & 3);
				}
				lasts = ch;
				pac[(size_t)(l_pac >> 2)] |= (uint8_t)(c << ((~l_pac & 3) << 1));
				l_pac++; a.len++;
			}
			if (p < n) p++;
		}
		anns.push_back(a);
		if (p < n && buf[p] == '+') { fprintf(stderr, "[gsa_index] FASTQ input is not supported\n"); return -1; }
	}
	if (anns.empty() || l_pac == 0) { fprintf(stderr, "[gsa_index] no sequence in %s\n", path); return -1; }
	pac.resize((size_t)((l_pac + 3) >> 2));
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


static int write_bwt_sa(const std::string &prefix, unsigned long long primary, const unsigned long long L2[5], unsigned long long seq_len, unsigned long long intv,
                        const std::vector<uint32_t> &bwt, const std::vector<unsigned long long> &samples)
{ // bwt_dump_bwt / bwt_dump_sa (reference src/BWT_Index/bwt.c:174-196)
	FILE *fp = fopen((prefix + ".bwt").c_str(), "wb");
	if (!fp) return -1;
	bool ok = fwrite(&primary, 8, 1, fp) == 1 && fwrite(L2 + 1, 8, 4, fp) == 4 && fwrite(bwt.data(), 4, bwt.size(), fp) == bwt.size();
	ok = fclose(fp) == 0 && ok;
	fp = fopen((prefix + ".sa").c_str(), "wb");
	if (!fp) return -1;
	ok = fwrite(&primary, 8, 1, fp) == 1 && fwrite(L2 + 1, 8, 4, fp) == 4 && fwrite(&intv, 8, 1, fp) == 1 && fwrite(&seq_len, 8, 1, fp) == 1 &&
	     fwrite(samples.data(), 8, samples.size(), fp) == samples.size() && ok;
	ok = fclose(fp) == 0 && ok;
	if (!ok) fprintf(stderr, "[gsa_index] short write on %s.bwt / .sa\n", prefix.c_str());
	return ok ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------------
// Blockwise suffix sorting for texts of any size (n >= 2^31 symbols, e.g. the 6.2 G symbols of a human-size pair).
// The full suffix array is never held: suffixes are cut into blocks by their first 6 bases (a histogram picks key ranges
// of at most `cap` suffixes), every block is sorted on its own and leaves only its BWT characters (2 bit per row) and its
// share of the SA samples.  Inside a block: one radix sort on (29 bases, length) keys, then the still-tied suffixes are
// refined 29 bases at a time with a segmented sort over the tie groups (depth-wise extension: rounds ~ longest repeat / 29,
// each over the tied suffixes only).  '$' < A is honoured by padding with A and breaking ties on the number of real bases.
// ---------------------------------------------------------------------------------------------------
#define IBW_KBASES 29
#define IBW_BIN_BASES 6
#define IBW_NBINS (1 << (2 * IBW_BIN_BASES))

__device__ __forceinline__ uint64_t ibw_key(const uint32_t *txt, uint64_t n, uint64_t p)
{ // (29 bases from p, padded with A) << 5 | number of real bases; the text is padded with 3 zero words
	if (p >= n) return 0;
	const uint64_t w = p >> 4; const uint32_t sh = ((uint32_t)p & 15) << 1;
	const uint32_t w0 = txt[w], w1 = txt[w + 1], w2 = txt[w + 2];
	uint64_t x = ((uint64_t)w0 << 32) | w1;
	if (sh) x = (x << sh) | ((uint64_t)w2 >> (32 - sh));
	x >>= 64 - 2 * IBW_KBASES;
	const uint64_t rem = n - p; const uint32_t len = rem < IBW_KBASES ? (uint32_t)rem : IBW_KBASES;
	if (len < IBW_KBASES) x &= ~0ull << ((IBW_KBASES - len) << 1);
	return (x << 5) | len;
}
__device__ __forceinline__ uint32_t ibw_bin(uint64_t key) { return (uint32_t)(key >> (5 + 2 * (IBW_KBASES - IBW_BIN_BASES))); }

// suffixes per 6-mer bin; shared-memory histogram per CTA, flushed with 64-bit atomics
__global__ void __launch_bounds__(512) k_ibw_hist(const uint32_t *txt, uint64_t n, unsigned long long *hist)
{
	__shared__ unsigned int sh[IBW_NBINS];
	for (int i = threadIdx.x; i < IBW_NBINS; i += blockDim.x) sh[i] = 0;
	__syncthreads();
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	// a CTA adds at most 2^32 - 1 to one bin: n / gridDim.x positions per CTA, and the grid is sized for that
	for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) atomicAdd(&sh[ibw_bin(ibw_key(txt, n, p))], 1u);
	__syncthreads();
	for (int i = threadIdx.x; i < IBW_NBINS; i += blockDim.x) if (sh[i]) atomicAdd(hist + i, (unsigned long long)sh[i]);
}

// (key, position) of every suffix whose bin lies in [bin_lo, bin_hi), in any order (warp-aggregated append)
__global__ void __launch_bounds__(256) k_ibw_select(const uint32_t *txt, uint64_t n, uint32_t bin_lo, uint32_t bin_hi, uint64_t *key, uint64_t *pos,
                                                    unsigned long long *count)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	const int lane = threadIdx.x & 31;
	for (uint64_t p0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; p0 < n; p0 += stride) {
		const uint64_t p = p0 + lane;
		uint64_t k = 0; bool take = false;
		if (p < n) { k = ibw_key(txt, n, p); const uint32_t b = ibw_bin(k); take = b >= bin_lo && b < bin_hi; }
		const unsigned m = __ballot_sync(0xffffffffu, take);
		if (!m) continue;
		unsigned long long base = 0;
		if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(m));
		base = __shfl_sync(0xffffffffu, base, 0);
		if (take) { const unsigned long long o = base + __popc(m & ((1u << lane) - 1)); key[o] = k; pos[o] = p; }
	}
}

__global__ void k_ibw_tied(const uint64_t *key, uint32_t m, uint8_t *flag)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	const uint64_t k = key[i];
	flag[i] = (i > 0 && key[i - 1] == k) || (i + 1 < m && key[i + 1] == k);
}

// the tied suffixes of a freshly sorted block: their positions and the heads of their tie groups
__global__ void k_ibw_tie_init(const uint32_t *tix, uint32_t t, const uint64_t *key, const uint64_t *pos, uint64_t *tp, uint8_t *head)
{
	uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= t) return;
	const uint32_t s = tix[j];
	tp[j] = pos[s];
	head[j] = j == 0 || key[s] != key[tix[j - 1]];
}

__global__ void k_ibw_depth_keys(const uint32_t *txt, uint64_t n, const uint64_t *tp, uint32_t t, uint64_t depth, uint64_t *key)
{
	uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j < t) key[j] = ibw_key(txt, n, tp[j] + depth);
}

// after the segmented sort of a round: sorted positions back to their slots, sub-group heads, still-tied flags
__global__ void k_ibw_round_close(const uint32_t *tix, uint32_t t, const uint64_t *k2, const uint64_t *tps, const uint8_t *head, uint64_t *pos, uint8_t *head2, uint8_t *tied)
{
	uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= t) return;
	pos[tix[j]] = tps[j];
	const bool h = head[j] || k2[j] != k2[j - 1];
	const bool next_same = j + 1 < t && !head[j + 1] && k2[j + 1] == k2[j];
	head2[j] = h;
	tied[j] = !h || next_same;
}

__global__ void k_ibw_compact(const uint32_t *sel, uint32_t t2, const uint32_t *tix, const uint64_t *tps, const uint8_t *head2, uint32_t *tix2, uint64_t *tp2, uint8_t *head3)
{
	uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= t2) return;
	const uint32_t s = sel[j];
	tix2[j] = tix[s]; tp2[j] = tps[s]; head3[j] = head2[s];
}

// BWT characters of the block's rows (row = row_base + slot; row 0 is the suffix "$") into the row-indexed 2-bit array
// (zeroed beforehand; neighbouring blocks share edge words, hence the atomic), SA samples of rows that are multiples of
// sa_intv, and the row of suffix 0 (primary)
__global__ void k_ibw_emit(const uint32_t *txt, const uint64_t *pos, uint32_t m, uint64_t row_base, int sa_intv, uint32_t *brow, unsigned long long *samples,
                           unsigned long long *primary)
{
	const uint64_t nrows = m;
	const uint64_t first_word = row_base >> 4;
	const uint64_t w = first_word + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if ((w << 4) >= row_base + nrows) return;
	uint32_t v = 0;
	for (int i = 0; i < 16; i++) {
		const uint64_t row = (w << 4) + i;
		if (row < row_base || row >= row_base + nrows) continue;
		const uint64_t p = pos[row - row_base];
		int c = 0;
		if (p == 0) *primary = row;
		else c = (int)(txt[(p - 1) >> 4] >> ((~(uint32_t)(p - 1) & 15) << 1)) & 3;
		v |= (uint32_t)c << ((15 - i) << 1);
		if ((row & (uint64_t)(sa_intv - 1)) == 0) samples[row / (uint64_t)sa_intv] = p;
	}
	if (v) atomicOr(brow + w, v);
}

// row 0 is the suffix "$" (SA = n): its BWT character is the last base of the text
__global__ void k_ibw_row0(const uint32_t *txt, uint64_t n, uint32_t *brow)
{
	const uint64_t p = n - 1;
	brow[0] |= ((txt[p >> 4] >> ((~(uint32_t)p & 15) << 1)) & 3u) << 30;
}

// the '$'-less BWT in the BWA layout: symbol x = character of row x + (x >= primary); one thread per 128-symbol block
// writes its 8 words in place and its per-base counts
__global__ void k_ibw_bwt_block(const uint32_t *brow, uint64_t n, uint64_t primary, uint32_t *out, unsigned long long *cnt, uint64_t nblk)
{
	const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nblk) return;
	unsigned long long c[4] = {0, 0, 0, 0};
	for (int w = 0; w < 8; w++) {
		const uint64_t x0 = b * 128 + (uint64_t)w * 16;
		if (x0 >= n) break;
		uint32_t v = 0;
		for (int i = 0; i < 16 && x0 + i < n; i++) {
			const uint64_t row = x0 + i + (x0 + i >= primary);
			const int s = (int)(brow[row >> 4] >> ((~(uint32_t)row & 15) << 1)) & 3;
			c[s]++; v |= (uint32_t)s << ((15 - i) << 1);
		}
		out[b * 16 + 8 + w] = v;
	}
	for (int s = 0; s < 4; s++) cnt[(size_t)s * (nblk + 1) + b] = c[s];
}

__global__ void k_ibw_bwt_counts(const unsigned long long *cum, uint64_t nblk, uint64_t n, uint32_t *out)
{ // cumulative counts in front of every block + the trailing block (bwtindex.c:61-71)
	const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (b > nblk) return;
	const size_t base = b < nblk ? (size_t)b * 16 : (size_t)nblk * 8 + (n + 15) / 16;
	unsigned long long *o = (unsigned long long *)(out + base);
	for (int s = 0; s < 4; s++) o[s] = cum[(size_t)s * (nblk + 1) + b];
}

struct IbwBufs { // freed on every exit path
	std::vector<void *> ptrs;
	~IbwBufs() { for (void *p : ptrs) if (p) cudaFree(p); }
	template <typename T> int get(T **out, size_t count)
	{
		void *p = nullptr;
		cudaError_t e = cudaMalloc(&p, (count ? count : 1) * sizeof(T));
		if (e != cudaSuccess) { fprintf(stderr, "[gsa_index] cudaMalloc(%zu): %s\n", count * sizeof(T), cudaGetErrorString(e)); return -2; }
		ptrs.push_back(p); *out = (T *)p;
		return 0;
	}
};

#define IB_LAUNCH_CHECK() IB_CHECK(cudaGetLastError())

// cap: most suffixes per block (0 = from the free device memory).  Returns the BWA-layout BWT words, SA samples, primary and
// per-base totals on the host.
static int build_blockwise(const uint8_t *h_pac, int64_t N, uint64_t cap, std::vector<uint32_t> &h_bwt, std::vector<unsigned long long> &h_samples,
                           unsigned long long &primary_out, unsigned long long L2[5], int &n_blocks_out, int &n_rounds_out)
{
	const uint64_t n = 2 * (uint64_t)N;
	const uint64_t nwords = (n >> 4) + 4;
	const int T = 256, intv = 32;
	IbwBufs B;
	uint8_t *d_pac; uint32_t *txt, *brow; unsigned long long *hist, *d_cnt, *samples;
	const uint64_t n_sa = (n + intv) / intv;
	if (B.get(&d_pac, (size_t)(N / 4 + 1)) || B.get(&txt, nwords) || B.get(&brow, ((n + 1) >> 4) + 2) || B.get(&hist, IBW_NBINS) || B.get(&d_cnt, 8) || B.get(&samples, n_sa + 1)) return -2;
	IB_CHECK(cudaMemcpy(d_pac, h_pac, (size_t)(N / 4 + 1), cudaMemcpyHostToDevice));
	IB_CHECK(cudaMemset(txt, 0, nwords * 4));
	k_ib_text<<<(unsigned)((nwords + T - 1) / T), T>>>(d_pac, N, txt, nwords);
	IB_LAUNCH_CHECK();
	IB_CHECK(cudaMemset(brow, 0, (((n + 1) >> 4) + 2) * 4));
	IB_CHECK(cudaMemset(hist, 0, IBW_NBINS * 8));
	IB_CHECK(cudaMemset(d_cnt, 0, 64));
	k_ibw_hist<<<148 * 4, 512>>>(txt, n, hist);
	IB_LAUNCH_CHECK();
	std::vector<unsigned long long> h_hist(IBW_NBINS);
	IB_CHECK(cudaMemcpy(h_hist.data(), hist, IBW_NBINS * 8, cudaMemcpyDeviceToHost));
	if (cap == 0) {
		size_t fr = 0, tot = 0;
		IB_CHECK(cudaMemGetInfo(&fr, &tot));
		cap = std::min<uint64_t>((uint64_t)1 << 30, fr / 96); // ~64 B per suffix in the sort buffers + tie-refinement scratch
	}
	unsigned long long biggest = 0;
	for (unsigned long long h : h_hist) biggest = std::max(biggest, h);
	if (biggest > cap) {
		if (biggest >= ((uint64_t)1 << 31)) { fprintf(stderr, "[gsa_index] one 6-mer starts %llu suffixes: text too repetitive for this build\n", biggest); return -3; }
		cap = biggest; // tests force tiny caps
	}
	uint64_t *k_in, *k_out, *p_in, *p_out, *tp, *tps, *k2, *k2s; uint32_t *tix, *tix2, *sel, *segs; uint8_t *flag, *head, *head2, *tied;
	if (B.get(&k_in, cap) || B.get(&k_out, cap) || B.get(&p_in, cap) || B.get(&p_out, cap)) return -2;
	// tie-refinement scratch grows on demand (a random text leaves next to nothing tied after 29 bases)
	size_t tcap = 0;
	tp = tps = k2 = k2s = nullptr; tix = tix2 = sel = segs = nullptr; head = head2 = tied = nullptr;
	if (B.get(&flag, cap)) return -2;
	void *tmp = nullptr; size_t tmp_cap = 0;
	struct TmpGuard { void **p; ~TmpGuard() { if (*p) cudaFree(*p); } } tg{&tmp};
	auto need_tmp = [&](size_t bytes) -> int { if (bytes > tmp_cap) { if (tmp) cudaFree(tmp); tmp = nullptr; IB_CHECK(cudaMalloc(&tmp, bytes)); tmp_cap = bytes; } return 0; };
	auto select_flagged = [&](const uint8_t *fl, uint32_t *out, uint32_t count, uint32_t *n_out) -> int {
		size_t bytes = 0;
		thrust::counting_iterator<uint32_t> it(0);
		cub::DeviceSelect::Flagged(nullptr, bytes, it, fl, out, (unsigned int *)d_cnt + 2, (int)count);
		if (need_tmp(bytes)) return -2;
		IB_CHECK(cub::DeviceSelect::Flagged(tmp, bytes, it, fl, out, (unsigned int *)d_cnt + 2, (int)count));
		IB_CHECK(cudaMemcpy(n_out, (unsigned int *)d_cnt + 2, 4, cudaMemcpyDeviceToHost));
		return 0;
	};
	uint64_t row_base = 1; // row 0 is the suffix "$"
	int n_blocks = 0, n_rounds = 0;
	for (uint32_t bin = 0; bin < IBW_NBINS;) {
		uint32_t hi = bin; uint64_t m64 = 0;
		while (hi < IBW_NBINS && m64 + h_hist[hi] <= cap) m64 += h_hist[hi++];
		if (m64 == 0) { bin = hi; continue; }
		const uint32_t m = (uint32_t)m64;
		n_blocks++;
		IB_CHECK(cudaMemset(d_cnt, 0, 8));
		k_ibw_select<<<148 * 8, 256>>>(txt, n, bin, hi, k_in, p_in, d_cnt);
		IB_LAUNCH_CHECK();
		unsigned long long got = 0;
		IB_CHECK(cudaMemcpy(&got, d_cnt, 8, cudaMemcpyDeviceToHost));
		if (got != m64) { fprintf(stderr, "[gsa_index] internal error: block holds %llu suffixes, histogram said %llu\n", got, (unsigned long long)m64); return -2; }
		{
			size_t bytes = 0;
			cub::DeviceRadixSort::SortPairs(nullptr, bytes, k_in, k_out, p_in, p_out, (int64_t)m, 0, 63);
			if (need_tmp(bytes)) return -2;
			IB_CHECK(cub::DeviceRadixSort::SortPairs(tmp, bytes, k_in, k_out, p_in, p_out, (int64_t)m, 0, 63));
		}
		// ---- refine the ties, 29 bases per round ----
		k_ibw_tied<<<(m + T - 1) / T, T>>>(k_out, m, flag);
		IB_LAUNCH_CHECK();
		// the tie scratch is sized by the first count of this block
		uint32_t t = 0;
		{
			uint32_t *probe = (uint32_t *)k_in; // k_in is free after the sort: room for m slot indices
			if (select_flagged(flag, probe, m, &t)) return -2;
			if (t > 0) {
				if (t > tcap) {
					tcap = (size_t)t + t / 8 + 1024;
					if (B.get(&tp, tcap) || B.get(&tps, tcap) || B.get(&k2, tcap) || B.get(&k2s, tcap) || B.get(&tix, tcap) || B.get(&tix2, tcap) || B.get(&sel, tcap) ||
					    B.get(&segs, tcap + 2) || B.get(&head, tcap + 1) || B.get(&head2, tcap + 1) || B.get(&tied, tcap + 1)) return -2;
				}
				IB_CHECK(cudaMemcpy(tix, probe, (size_t)t * 4, cudaMemcpyDeviceToDevice));
				k_ibw_tie_init<<<(t + T - 1) / T, T>>>(tix, t, k_out, p_out, tp, head);
				IB_LAUNCH_CHECK();
			}
		}
		for (uint64_t depth = IBW_KBASES; t > 0; depth += IBW_KBASES) {
			n_rounds++;
			uint32_t nseg = 0;
			if (select_flagged(head, segs, t, &nseg)) return -2;
			IB_CHECK(cudaMemcpy(segs + nseg, &t, 4, cudaMemcpyHostToDevice));
			k_ibw_depth_keys<<<(t + T - 1) / T, T>>>(txt, n, tp, t, depth, k2);
			IB_LAUNCH_CHECK();
			size_t bytes = 0;
			cub::DeviceSegmentedSort::SortPairs(nullptr, bytes, k2, k2s, tp, tps, (int)t, (int)nseg, segs, segs + 1);
			if (need_tmp(bytes)) return -2;
			IB_CHECK(cub::DeviceSegmentedSort::SortPairs(tmp, bytes, k2, k2s, tp, tps, (int)t, (int)nseg, segs, segs + 1));
			k_ibw_round_close<<<(t + T - 1) / T, T>>>(tix, t, k2s, tps, head, p_out, head2, tied);
			IB_LAUNCH_CHECK();
			uint32_t t2 = 0;
			if (select_flagged(tied, sel, t, &t2)) return -2;
			if (t2 > 0) {
				k_ibw_compact<<<(t2 + T - 1) / T, T>>>(sel, t2, tix, tps, head2, tix2, tp, head);
				IB_LAUNCH_CHECK();
				std::swap(tix, tix2);
			}
			t = t2;
		}
		k_ibw_emit<<<(unsigned)(((uint64_t)m + 32) / 16 / T + 1), T>>>(txt, p_out, m, row_base, intv, brow, samples, d_cnt + 4);
		IB_LAUNCH_CHECK();
		row_base += m;
		bin = hi;
	}
	k_ibw_row0<<<1, 1>>>(txt, n, brow);
	IB_LAUNCH_CHECK();
	IB_CHECK(cudaDeviceSynchronize());
	if (row_base != n + 1) { fprintf(stderr, "[gsa_index] internal error: %llu rows sorted, expected %llu\n", (unsigned long long)row_base, (unsigned long long)(n + 1)); return -2; }
	unsigned long long primary = 0;
	IB_CHECK(cudaMemcpy(&primary, d_cnt + 4, 8, cudaMemcpyDeviceToHost));
	// the sort buffers are no longer needed: free them before the output arrays are allocated
	for (void *&q : B.ptrs) if (q == k_in || q == k_out || q == p_in || q == p_out) { cudaFree(q); q = nullptr; }
	// ---- BWT + Occ in the BWA layout ----
	const uint64_t nblk = (n + 127) / 128;
	const size_t bwt_words = (size_t)nblk * 8 + (size_t)((n + 15) / 16) + 8;
	uint32_t *d_bwt; unsigned long long *d_c, *d_cum;
	if (B.get(&d_bwt, bwt_words) || B.get(&d_c, (size_t)4 * (nblk + 1)) || B.get(&d_cum, (size_t)4 * (nblk + 1))) return -2;
	IB_CHECK(cudaMemset(d_bwt, 0, bwt_words * 4));
	IB_CHECK(cudaMemset(d_c, 0, (size_t)4 * (nblk + 1) * 8));
	k_ibw_bwt_block<<<(unsigned)((nblk + 127) / 128), 128>>>(brow, n, primary, d_bwt, d_c, nblk);
	IB_LAUNCH_CHECK();
	for (int s = 0; s < 4; s++) {
		size_t bytes = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_c + (size_t)s * (nblk + 1), d_cum + (size_t)s * (nblk + 1), (int64_t)(nblk + 1));
		if (need_tmp(bytes)) return -2;
		IB_CHECK(cub::DeviceScan::ExclusiveSum(tmp, bytes, d_c + (size_t)s * (nblk + 1), d_cum + (size_t)s * (nblk + 1), (int64_t)(nblk + 1)));
	}
	k_ibw_bwt_counts<<<(unsigned)((nblk + 1 + T - 1) / T), T>>>(d_cum, nblk, n, d_bwt);
	IB_LAUNCH_CHECK();
	h_bwt.resize(bwt_words);
	h_samples.resize(n_sa > 1 ? n_sa - 1 : 0);
	IB_CHECK(cudaMemcpy(h_bwt.data(), d_bwt, bwt_words * 4, cudaMemcpyDeviceToHost));
	if (n_sa > 1) IB_CHECK(cudaMemcpy(h_samples.data(), samples + 1, (size_t)(n_sa - 1) * 8, cudaMemcpyDeviceToHost));
	unsigned long long tot[4];
	for (int s = 0; s < 4; s++) IB_CHECK(cudaMemcpy(&tot[s], d_cum + (size_t)s * (nblk + 1) + nblk, 8, cudaMemcpyDeviceToHost));
	L2[0] = 0;
	for (int s = 0; s < 4; s++) L2[s + 1] = L2[s] + tot[s];
	primary_out = primary; n_blocks_out = n_blocks; n_rounds_out = n_rounds;
	return 0;
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
	if (write_pac_ann_amb(prefix, pac, N, anns, holes) != 0) { fprintf(stderr, "[gsa_index] cannot write %s.*\n", prefix_c); return -1; }
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) { fprintf(stderr, "[gsa_index] no CUDA device (index construction has no CPU path in this build)\n"); return -2; }
	IB_CHECK(cudaSetDevice(device));
	pac.resize((size_t)(N / 4 + 1), 0);
	// texts of 2^31 symbols and more (and tests, GSA_INDEX_BLOCK=<suffixes per block>) take the blockwise sorter
	const char *force = getenv("GSA_INDEX_BLOCK");
	if (2 * (uint64_t)N >= 0x7FFFFF00ull || force) {
		std::vector<uint32_t> h_bwt; std::vector<unsigned long long> h_samples; unsigned long long prim64 = 0, L2[5]; int nb = 0, nr = 0;
		int rc = build_blockwise(pac.data(), N, force ? strtoull(force, nullptr, 10) : 0, h_bwt, h_samples, prim64, L2, nb, nr);
		if (rc != 0) return rc;
		if (write_bwt_sa(prefix, prim64, L2, 2 * (unsigned long long)N, 32, h_bwt, h_samples) != 0) return -1;
		fprintf(stderr, "[gsa_index] %lld bp, %d sequence(s), %d block(s), %d refinement round(s) -> %s.{pac,ann,amb,bwt,sa}\n", (long long)N, (int)anns.size(), nb, nr, prefix_c);
		return 0;
	}
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
	IB_CHECK(cudaGetLastError());
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
	if (write_bwt_sa(prefix, primary, L2, n, intv, h_bwt, h_samples) != 0) return -1;
	cudaFree(d_pac); cudaFree(txt); cudaFree(sa); cudaFree(rank); cudaFree(v_in); cudaFree(v_out); cudaFree(head); cudaFree(slots);
	cudaFree(k_in); cudaFree(k_out); cudaFree(flag); cudaFree(d_cnt); cudaFree(d_bwt); cudaFree(d_c); cudaFree(d_cum); cudaFree(d_samples);
	if (tmp) cudaFree(tmp);
	fprintf(stderr, "[gsa_index] %lld bp, %d sequence(s), %d doubling round(s) -> %s.{pac,ann,amb,bwt,sa}\n", (long long)N, (int)anns.size(), rounds, prefix_c);
	return 0;
}
