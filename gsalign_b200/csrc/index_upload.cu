// index_upload.cu -- re-lays the BWA-format FM-index out for HBM (DevIndex, gsa_internal.cuh).
//
// Replaces, for the device side, bwa_idx_load() + RestoreReferenceInfo() (reference
// src/bwt_index.cpp:147-159,229-264).  Every structure built here is a canonical function of the
// text T, so search results are identical to walking the BWA layout:
//   * 32-byte rank blocks (one sector per Occ query instead of a 64-byte block)
//   * the text itself, 2 bit/base, so that a search whose interval has shrunk to one row can finish
//     by comparing against T directly instead of one dependent rank query per base
//   * the full suffix array (no LF-walk at locate time), expanded on the device from the 1/32
//     row-sampled SA by walking LF from every sample
//   * a k-mer prefix table replacing the first k backward-search steps
#include "fm.cuh"

// ---- views of the BWA layout on the device (only used while uploading) --------------------------
struct BwaView {
	const uint32_t *bwt;   // Occ-interleaved words: per 128 symbols 4 x u64 counts + 8 x u32 symbols
	uint64_t primary;
	uint64_t n;
};

__device__ __forceinline__ int bwa_sym(const BwaView &b, uint64_t x)
{ // symbol x of the '$'-less BWT (reference bwt_B0, src/bwt_search.cpp:32-34)
	uint32_t w = b.bwt[((x >> 7) << 4) + 8 + ((x & 127) >> 4)];
	return (int)(w >> ((~x & 15) << 1)) & 3;
}

// one thread per 32-byte rank block
__global__ void k_build_occ(BwaView b, uint4 *occ, uint64_t nblocks)
{
	uint64_t blk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (blk >= nblocks) return;
	uint64_t r0 = blk << 6;
	// symbols of rows r0 .. r0+63
	uint32_t sym[4] = {0, 0, 0, 0};
	for (int i = 0; i < 64; i++) {
		uint64_t r = r0 + i;
		int s = 0;
		if (r <= b.n && r != b.primary) s = bwa_sym(b, r - (r > b.primary));
		sym[i >> 4] |= (uint32_t)s << ((15 - (i & 15)) << 1);
	}
	// counts of rows [0, r0): x0 real symbols precede row r0, plus the placeholder A of the '$' row
	uint32_t cnt[4] = {0, 0, 0, 0};
	if (r0 <= b.n) {
		uint64_t x0 = r0 - (r0 > b.primary);
		const uint64_t *cum = (const uint64_t *)(b.bwt + ((x0 >> 7) << 4));
		for (int c = 0; c < 4; c++) cnt[c] = (uint32_t)cum[c];
		for (uint64_t x = x0 & ~127ull; x < x0; x++) cnt[bwa_sym(b, x)]++;
		if (b.primary < r0) cnt[0]++;
	}
	occ[2 * blk] = make_uint4(cnt[0], cnt[1], cnt[2], cnt[3]);
	occ[2 * blk + 1] = make_uint4(sym[0], sym[1], sym[2], sym[3]);
}

// one thread per 16 bases of T = F . revcomp(F) (reference IdvLoadReferenceSequences, src/bwt_index.cpp:193-212)
__global__ void k_build_text(const uint8_t *pac, int64_t N, uint32_t *txt, uint64_t nwords)
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

// Full SA from the row-sampled one: the sample at row 32s holds SA = v; LF maps the row with SA = v
// to the row with SA = v-1, so walking LF until the next sampled row fills every row exactly once.
__global__ void k_fill_sa(DevIndex ix, const uint64_t *samples, uint64_t n_sa, int sa_intv, uint32_t *sa)
{
	uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n_sa) return;
	uint32_t r = (uint32_t)(s * (uint64_t)sa_intv);
	uint32_t v = s == 0 ? ix.n : (uint32_t)samples[s];   // samples[0] = -1 stands for SA[0] = n
	sa[r] = v;
	uint32_t mask = (uint32_t)sa_intv - 1;
	for (;;) {
		if (r == ix.primary) break;                      // LF(primary) = row 0, which is sampled
		int c = gsa_bwt_char(ix, r);
		r = ix.L2[c] + gsa_occ(ix, c, r);
		v--;
		if ((r & mask) == 0) break;
		sa[r] = v;
	}
}

// one thread per k-mer: interval of revcomp(kmer) by k backward-search steps
__global__ void k_build_ktab(DevIndex ix, int k, uint2 *ktab, uint32_t ncodes)
{
	uint32_t code = blockIdx.x * blockDim.x + threadIdx.x;
	if (code >= ncodes) return;
	int b0 = (code >> ((k - 1) << 1)) & 3, c = 3 - b0;
	uint32_t lo = ix.L2[c] + 1, size = ix.L2[c + 1] - ix.L2[c];
	for (int j = 1; j < k && size > 0; j++) {
		int b = (code >> ((k - 1 - j) << 1)) & 3;
		c = 3 - b;
		uint32_t o1, o2;
		gsa_occ2(ix, c, lo - 1, lo + size - 1, o1, o2);
		lo = ix.L2[c] + o1 + 1; size = o2 - o1;
	}
	// a k-mer that occurs once needs no rank step, only its position: the entry carries SA[lo] itself and the search
	// goes from the table straight to the text (size == 1 <=> .x is a suffix-array value, not a row)
	if (size == 1) lo = __ldg(ix.sa + lo);
	ktab[code] = make_uint2(lo, size);
}

int gsa_impl_build_ktab(gsa_ctx *ctx, int k)
{
	if (k > GSA_KTAB_MAX_K) k = GSA_KTAB_MAX_K;
	if (k < 1) k = 1;
	if (ctx->ix.ktab_k == k && ctx->ix.ktab) return GSA_OK;
	uint32_t ncodes = 1u << (2 * k);
	GSA_TRY(gsa_ensure(ctx, ctx->d_ktab, (size_t)ncodes * sizeof(uint2)));
	ctx->ix.ktab = nullptr; ctx->ix.ktab_k = 0;
	k_build_ktab<<<gsa_grid(ncodes, 256), 256, 0, ctx->stream>>>(ctx->ix, k, (uint2 *)ctx->d_ktab.p, ncodes);
	KERNEL_CHECK(ctx);
	ctx->ix.ktab = (const uint2 *)ctx->d_ktab.p; ctx->ix.ktab_k = k;
	return GSA_OK;
}

// one thread per text position: marks the k-mer starting there (every window of T counts, junction-spanning ones included:
// BWT_Search does not filter them either, SURVEY.md hazard H12)
__global__ void k_build_kbits(DevIndex ix, int k, uint32_t *bits)
{
	uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p + (uint64_t)k > ix.n) return;
	uint32_t code = gsa_pk_window(ix.txt, (uint32_t)p) >> (32 - 2 * k);
	uint32_t bit = 1u << (code & 31);
	uint32_t *w = bits + (code >> 5);
	if (!(*w & bit)) atomicOr(w, bit); // most k-mers are already marked after the first pass over a repeat-free genome's 2 strands
}

int gsa_impl_build_kbits(gsa_ctx *ctx, int k)
{
	if (k > GSA_KBITS_MAX_K) k = GSA_KBITS_MAX_K;
	if (k < 1) k = 1;
	// worth its accesses only while most k-mers are absent from T: with |T| >= 4^k / 2 nearly every k-mer occurs
	if ((double)ctx->ix.n >= 0.5 * (double)(1ull << (2 * k))) { ctx->ix.kbits = nullptr; ctx->ix.kbits_k = k; return GSA_OK; }
	if (ctx->ix.kbits_k == k && ctx->ix.kbits) return GSA_OK;
	size_t words = k >= 3 ? ((size_t)1 << (2 * k - 5)) : 1;
	GSA_TRY(gsa_ensure(ctx, ctx->d_kbits, words * 4));
	ctx->ix.kbits = nullptr; ctx->ix.kbits_k = 0;
	CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_kbits.p, 0, words * 4, ctx->stream));
	k_build_kbits<<<gsa_grid((int64_t)ctx->ix.n, 256), 256, 0, ctx->stream>>>(ctx->ix, k, (uint32_t *)ctx->d_kbits.p);
	KERNEL_CHECK(ctx);
	ctx->ix.kbits = (const uint32_t *)ctx->d_kbits.p; ctx->ix.kbits_k = k;
	return GSA_OK;
}

int gsa_impl_index_upload(gsa_ctx *ctx, const gsa_index_view *v)
{
	if (!v || !v->bwt || !v->sa || !v->pac || v->n_contigs <= 0) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_index_upload: incomplete view");
	if (v->seq_len != 2 * (uint64_t)v->l_pac) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_index_upload: seq_len != 2*l_pac");
	if (v->seq_len >= 0xFFFFFF00ull) return gsa_fail(ctx, GSA_ERR_LIMIT, "gsa_index_upload: text of %llu symbols needs the 64-bit row build (this build: < 2^32)", (unsigned long long)v->seq_len);
	if (v->sa_intv <= 0 || (v->sa_intv & (v->sa_intv - 1))) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_index_upload: sa_intv must be a power of two");
	ctx->have_index = false;
	uint64_t n = v->seq_len;
	ctx->N = v->l_pac;
	ctx->ix.n = (uint32_t)n; ctx->ix.primary = (uint32_t)v->primary;
	for (int i = 0; i < 5; i++) ctx->ix.L2[i] = (uint32_t)v->L2[i];
	ctx->ix.ktab = nullptr; ctx->ix.ktab_k = 0; ctx->ix.kbits = nullptr; ctx->ix.kbits_k = 0;

	uint64_t nblocks = (n >> 6) + 2, nwords = (n >> 4) + 3;
	GSA_TRY(gsa_ensure(ctx, ctx->d_occ, nblocks * 32));
	GSA_TRY(gsa_ensure(ctx, ctx->d_txt, nwords * 4));
	GSA_TRY(gsa_ensure(ctx, ctx->d_sa, (n + 1) * 4));

	// staging copies of the BWA arrays (freed before returning)
	uint32_t *d_bwt = nullptr; uint64_t *d_samples = nullptr; uint8_t *d_pac = nullptr;
	size_t pac_bytes = (size_t)(v->l_pac / 4 + 1);
	CUDA_TRY(ctx, cudaMalloc(&d_bwt, v->bwt_size * 4));
	CUDA_TRY(ctx, cudaMalloc(&d_samples, v->n_sa * 8));
	CUDA_TRY(ctx, cudaMalloc(&d_pac, pac_bytes));
	CUDA_TRY(ctx, cudaMemcpyAsync(d_bwt, v->bwt, v->bwt_size * 4, cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(ctx, cudaMemcpyAsync(d_samples, v->sa, v->n_sa * 8, cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(ctx, cudaMemcpyAsync(d_pac, v->pac, pac_bytes, cudaMemcpyHostToDevice, ctx->stream));

	BwaView bv; bv.bwt = d_bwt; bv.primary = v->primary; bv.n = n;
	k_build_occ<<<gsa_grid((int64_t)nblocks, 128), 128, 0, ctx->stream>>>(bv, (uint4 *)ctx->d_occ.p, nblocks);
	KERNEL_CHECK(ctx);
	k_build_text<<<gsa_grid((int64_t)nwords, 256), 256, 0, ctx->stream>>>(d_pac, v->l_pac, (uint32_t *)ctx->d_txt.p, nwords);
	KERNEL_CHECK(ctx);
	ctx->ix.occ = (const uint4 *)ctx->d_occ.p; ctx->ix.txt = (const uint32_t *)ctx->d_txt.p; ctx->ix.sa = (const uint32_t *)ctx->d_sa.p;
	k_fill_sa<<<gsa_grid((int64_t)v->n_sa, 128), 128, 0, ctx->stream>>>(ctx->ix, d_samples, v->n_sa, v->sa_intv, (uint32_t *)ctx->d_sa.p);
	KERNEL_CHECK(ctx);
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	cudaFree(d_bwt); cudaFree(d_samples); cudaFree(d_pac);

	// ChrLocMap (reference src/bwt_index.cpp:241-253): inclusive end of every contig on both strands
	ctx->cend.clear(); ctx->contig_off.assign(v->contig_off, v->contig_off + v->n_contigs);
	ctx->contig_len.assign(v->contig_len, v->contig_len + v->n_contigs);
	int64_t total = 0;
	for (int i = 0; i < v->n_contigs; i++) {
		ContigEnd f; f.end = total + v->contig_len[i] - 1; f.idx = i; f.pad = 0;
		total += v->contig_len[i];
		ContigEnd r; r.end = ((int64_t)n - total) + v->contig_len[i] - 1; r.idx = i; r.pad = 0;
		ctx->cend.push_back(f); ctx->cend.push_back(r);
	}
	if (total != v->l_pac) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_index_upload: contig lengths do not sum to l_pac");
	for (size_t i = 1; i < ctx->cend.size(); i++) // insertion sort by end (2K entries)
		for (size_t j = i; j > 0 && ctx->cend[j].end < ctx->cend[j - 1].end; j--) std::swap(ctx->cend[j], ctx->cend[j - 1]);
	GSA_TRY(gsa_ensure(ctx, ctx->d_cend, ctx->cend.size() * sizeof(ContigEnd)));
	CUDA_TRY(ctx, cudaMemcpy(ctx->d_cend.p, ctx->cend.data(), ctx->cend.size() * sizeof(ContigEnd), cudaMemcpyHostToDevice));
	ctx->have_index = true;
	return GSA_OK;
}
