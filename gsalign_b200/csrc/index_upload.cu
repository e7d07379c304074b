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
#include <stdlib.h>
#include <algorithm>

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
template <bool W>
__global__ void k_fill_sa(DevIndex ix, const uint64_t *samples, uint64_t n_sa, int sa_intv, void *sa)
{
	typedef typename RowT<W>::t row_t;
	uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n_sa) return;
	row_t r = (row_t)(s * (uint64_t)sa_intv);
	row_t v = s == 0 ? (row_t)ix.n : (row_t)samples[s];   // samples[0] = -1 stands for SA[0] = n
	gsa_sa_write<W>(sa, r, v);
	const row_t mask = (row_t)sa_intv - 1, primary = (row_t)ix.primary;
	for (;;) {
		if (r == primary) break;                         // LF(primary) = row 0, which is sampled
		int c = gsa_bwt_char(ix, r);
		r = (row_t)ix.L2[c] + gsa_occ<W>(ix, c, r);
		v--;
		if ((r & mask) == 0) break;
		gsa_sa_write<W>(sa, r, v);
	}
}

// one thread per k-mer: interval of revcomp(kmer) by k backward-search steps
template <bool W>
__global__ void k_build_ktab(DevIndex ix, int k, void *ktab_v, uint32_t ncodes)
{
	typedef typename RowT<W>::t row_t;
	uint32_t code = blockIdx.x * blockDim.x + threadIdx.x;
	if (code >= ncodes) return;
	int b0 = (code >> ((k - 1) << 1)) & 3, c = 3 - b0;
	row_t lo = (row_t)ix.L2[c] + 1, size = (row_t)(ix.L2[c + 1] - ix.L2[c]);
	for (int j = 1; j < k && size > 0; j++) {
		int b = (code >> ((k - 1 - j) << 1)) & 3;
		c = 3 - b;
		uint32_t o1, o2;
		gsa_occ2<W>(ix, c, lo - 1, lo + size - 1, o1, o2);
		lo = (row_t)ix.L2[c] + o1 + 1; size = o2 - o1;
	}
	// a k-mer that occurs once needs no rank step, only its position: the entry carries SA[lo] itself and the search
	// goes from the table straight to the text (size == 1 <=> .x is a suffix-array value, not a row)
	if (size == 1) lo = gsa_sa_read<W>(ix, lo);
	typename RowT<W>::ktab_t e; e.x = lo; e.y = size;
	((typename RowT<W>::ktab_t *)ktab_v)[code] = e;
}

int gsa_impl_build_ktab(gsa_ctx *ctx, int k)
{
	if (k > GSA_KTAB_MAX_K) k = GSA_KTAB_MAX_K;
	if (k < 1) k = 1;
	if (ctx->ix.ktab_k == k && ctx->ix.ktab) return GSA_OK;
	uint32_t ncodes = 1u << (2 * k);
	GSA_TRY(gsa_ensure(ctx, ctx->d_ktab, (size_t)ncodes * (ctx->ix.wide ? sizeof(ulonglong2) : sizeof(uint2))));
	ctx->ix.ktab = nullptr; ctx->ix.ktab_k = 0;
	if (ctx->ix.wide) k_build_ktab<true><<<gsa_grid(ncodes, 256), 256, 0, ctx->stream>>>(ctx->ix, k, ctx->d_ktab.p, ncodes);
	else k_build_ktab<false><<<gsa_grid(ncodes, 256), 256, 0, ctx->stream>>>(ctx->ix, k, ctx->d_ktab.p, ncodes);
	KERNEL_CHECK(ctx);
	ctx->ix.ktab = ctx->d_ktab.p; ctx->ix.ktab_k = k;
	return GSA_OK;
}

// one thread per text position: marks the k-mer starting there (every window of T counts, junction-spanning ones included:
// BWT_Search does not filter them either, SURVEY.md hazard H12)
__global__ void k_build_kbits(DevIndex ix, int k, uint32_t *bits)
{
	uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p + (uint64_t)k > ix.n) return;
	uint32_t code = gsa_pk_window(ix.txt, p) >> (32 - 2 * k);
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

// bytes of the full suffix array of n + 1 rows in the layout of the given width
static size_t sa_bytes(uint64_t n, bool wide)
{
	return wide ? ((n + 1 + GSA_SA_GROUP - 1) / GSA_SA_GROUP + 1) * 32 : (size_t)(n + 1) * 4;
}

// ChrLocMap (reference src/bwt_index.cpp:241-253): inclusive end of every contig on both strands, sorted by end
static int build_contig_table(gsa_ctx *ctx, uint64_t n, int64_t l_pac, int n_contigs, const int64_t *off, const int32_t *len)
{
	ctx->cend.clear(); ctx->contig_off.assign(off, off + n_contigs); ctx->contig_len.assign(len, len + n_contigs);
	int64_t total = 0;
	for (int i = 0; i < n_contigs; i++) {
		ContigEnd f; f.end = total + len[i] - 1; f.idx = i; f.pad = len[i];
		total += len[i];
		ContigEnd r; r.end = ((int64_t)n - total) + len[i] - 1; r.idx = i; r.pad = len[i];
		ctx->cend.push_back(f); ctx->cend.push_back(r);
	}
	if (total != l_pac) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_index_upload: contig lengths do not sum to l_pac");
	std::sort(ctx->cend.begin(), ctx->cend.end(), [](const ContigEnd &a, const ContigEnd &b) { return a.end < b.end; });
	GSA_TRY(gsa_ensure(ctx, ctx->d_cend, ctx->cend.size() * sizeof(ContigEnd)));
	CUDA_TRY(ctx, cudaMemcpy(ctx->d_cend.p, ctx->cend.data(), ctx->cend.size() * sizeof(ContigEnd), cudaMemcpyHostToDevice));
	return GSA_OK;
}

struct StagingGuard { // the BWA arrays on the device while the index is re-laid out
	void *p[3] = {nullptr, nullptr, nullptr};
	~StagingGuard() { for (void *q : p) if (q) cudaFree(q); }
};

int gsa_impl_index_upload(gsa_ctx *ctx, const gsa_index_view *v)
{
	if (!v || !v->bwt || !v->sa || !v->pac || v->n_contigs <= 0) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_index_upload: incomplete view");
	if (v->seq_len != 2 * (uint64_t)v->l_pac) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_index_upload: seq_len != 2*l_pac");
	if (v->sa_intv <= 0 || (v->sa_intv & (v->sa_intv - 1))) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_index_upload: sa_intv must be a power of two");
	const uint64_t n = v->seq_len;
	const bool wide = ctx->force_wide || getenv("GSA_FORCE_WIDE") != nullptr || n >= 0xFFFFFF00ull;
	if (n >= (1ull << 40) - 256) return gsa_fail(ctx, GSA_ERR_LIMIT, "gsa_index_upload: text of %llu symbols exceeds the 40-bit suffix array of this build", (unsigned long long)n);
	for (int c = 0; c < 4; c++)
		if (v->L2[c + 1] - v->L2[c] >= 0xFFFFFFF0ull) return gsa_fail(ctx, GSA_ERR_LIMIT, "gsa_index_upload: more than 2^32 occurrences of one base (rank blocks hold u32 counts)");
	ctx->have_index = false;
	ctx->N = v->l_pac;
	ctx->ix.n = n; ctx->ix.primary = v->primary; ctx->ix.wide = wide ? 1 : 0;
	for (int i = 0; i < 5; i++) ctx->ix.L2[i] = v->L2[i];
	ctx->ix.ktab = nullptr; ctx->ix.ktab_k = 0; ctx->ix.kbits = nullptr; ctx->ix.kbits_k = 0;

	uint64_t nblocks = (n >> 6) + 2, nwords = (n >> 4) + 3;
	GSA_TRY(gsa_ensure(ctx, ctx->d_occ, nblocks * 32));
	GSA_TRY(gsa_ensure(ctx, ctx->d_txt, nwords * 4));
	GSA_TRY(gsa_ensure(ctx, ctx->d_sa, sa_bytes(n, wide)));

	// staging copies of the BWA arrays (freed when the guard goes out of scope, error paths included)
	StagingGuard sg;
	size_t pac_bytes = (size_t)(v->l_pac / 4 + 1);
	CUDA_TRY(ctx, cudaMalloc(&sg.p[0], v->bwt_size * 4));
	CUDA_TRY(ctx, cudaMalloc(&sg.p[1], v->n_sa * 8));
	CUDA_TRY(ctx, cudaMalloc(&sg.p[2], pac_bytes));
	uint32_t *d_bwt = (uint32_t *)sg.p[0]; uint64_t *d_samples = (uint64_t *)sg.p[1]; uint8_t *d_pac = (uint8_t *)sg.p[2];
	CUDA_TRY(ctx, cudaMemcpyAsync(d_bwt, v->bwt, v->bwt_size * 4, cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(ctx, cudaMemcpyAsync(d_samples, v->sa, v->n_sa * 8, cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(ctx, cudaMemcpyAsync(d_pac, v->pac, pac_bytes, cudaMemcpyHostToDevice, ctx->stream));

	BwaView bv; bv.bwt = d_bwt; bv.primary = v->primary; bv.n = n;
	k_build_occ<<<gsa_grid((int64_t)nblocks, 128), 128, 0, ctx->stream>>>(bv, (uint4 *)ctx->d_occ.p, nblocks);
	KERNEL_CHECK(ctx);
	k_build_text<<<gsa_grid((int64_t)nwords, 256), 256, 0, ctx->stream>>>(d_pac, v->l_pac, (uint32_t *)ctx->d_txt.p, nwords);
	KERNEL_CHECK(ctx);
	ctx->ix.occ = (const uint4 *)ctx->d_occ.p; ctx->ix.txt = (const uint32_t *)ctx->d_txt.p; ctx->ix.sa = ctx->d_sa.p;
	if (wide) k_fill_sa<true><<<gsa_grid((int64_t)v->n_sa, 128), 128, 0, ctx->stream>>>(ctx->ix, d_samples, v->n_sa, v->sa_intv, ctx->d_sa.p);
	else k_fill_sa<false><<<gsa_grid((int64_t)v->n_sa, 128), 128, 0, ctx->stream>>>(ctx->ix, d_samples, v->n_sa, v->sa_intv, ctx->d_sa.p);
	KERNEL_CHECK(ctx);
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	GSA_TRY(build_contig_table(ctx, n, v->l_pac, v->n_contigs, v->contig_off, v->contig_len));
	ctx->have_index = true;
	return GSA_OK;
}

// A replica of src's device index on dst's GPU, copied over NVLink (cudaMemcpyPeerAsync) instead of being uploaded and
// re-derived from the host files once per GPU: the derived structures (full SA, prefix table) are 6-9 x the size of the
// BWA files.  src must have its index, parameters and prefix table ready; dst gets its own copy of everything.
int gsa_impl_index_clone(gsa_ctx *dst, gsa_ctx *src)
{
	if (!src->have_index) return gsa_fail(dst, GSA_ERR_ARG, "gsa_index_clone: the source has no index");
	CUDA_TRY(dst, cudaSetDevice(dst->device));
	if (dst->device != src->device) { // direct NVLink copies; without peer access the runtime stages through the host
		int can = 0;
		if (cudaDeviceCanAccessPeer(&can, dst->device, src->device) == cudaSuccess && can) {
			cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return gsa_fail(dst, GSA_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
			cudaGetLastError();
		}
	}
	dst->have_index = false;
	dst->ix = src->ix; dst->N = src->N; dst->prm = src->prm;
	struct { DevBuf *d; const DevBuf *s; const void **slot; } parts[] = {
		{&dst->d_occ, &src->d_occ, (const void **)&dst->ix.occ}, {&dst->d_txt, &src->d_txt, (const void **)&dst->ix.txt},
		{&dst->d_sa, &src->d_sa, &dst->ix.sa}, {&dst->d_ktab, &src->d_ktab, &dst->ix.ktab}, {&dst->d_kbits, &src->d_kbits, (const void **)&dst->ix.kbits}};
	for (auto &pt : parts) {
		if (!*pt.slot) continue;                 // structure not built on the source (e.g. no presence bitmap)
		GSA_TRY(gsa_ensure(dst, *pt.d, pt.s->cap));
		CUDA_TRY(dst, cudaMemcpyPeerAsync(pt.d->p, dst->device, pt.s->p, src->device, pt.s->cap, dst->stream));
		*pt.slot = pt.d->p;
	}
	CUDA_TRY(dst, cudaStreamSynchronize(dst->stream));
	GSA_TRY(build_contig_table(dst, src->ix.n, src->N, (int)src->contig_len.size(), src->contig_off.data(), src->contig_len.data()));
	dst->have_index = true;
	return GSA_OK;
}

// ---- self-check of the device index (debug hook; used on texts too large for the reference's indexer to cross-check) ------
// For pseudo-random rows r: (1) suffix SA[r] sorts before suffix SA[r+1] (direct text comparison, '$' smallest),
// (2) the BWT character stored for r is T[SA[r] - 1], (3) SA[LF(r)] = SA[r] - 1 (rank counts + L2 + SA agree).
template <bool W>
__global__ void k_index_selfcheck(DevIndex ix, uint64_t n_samples, unsigned long long *bad)
{
	typedef typename RowT<W>::t row_t;
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_samples) return;
	uint64_t h = (i + 1) * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
	const row_t r = (row_t)(h % ix.n);                 // rows 0 .. n-1, so r + 1 exists
	const uint64_t a = gsa_sa_read<W>(ix, r), b = gsa_sa_read<W>(ix, r + 1);
	unsigned fail = 0;
	if (a > ix.n || b >= ix.n || a == b) fail |= 1;
	else {
		for (uint64_t k = 0; k < 4096; k++) {
			if (a + k >= ix.n) break;                    // a ran into '$' first: smaller, fine
			if (b + k >= ix.n) { fail |= 2; break; }
			int ca = gsa_pk_base(ix.txt, a + k), cb = gsa_pk_base(ix.txt, b + k);
			if (ca != cb) { if (ca > cb) fail |= 2; break; }
		}
		if (a > 0 && r != (row_t)ix.primary) {
			int c = gsa_bwt_char(ix, r);
			if (c != gsa_pk_base(ix.txt, a - 1)) fail |= 4;
			row_t lf = (row_t)ix.L2[c] + gsa_occ<W>(ix, c, r);
			if ((uint64_t)gsa_sa_read<W>(ix, lf) != a - 1) fail |= 8;
		}
		if ((a == 0) != (r == (row_t)ix.primary)) fail |= 16;
	}
	if (fail) atomicAdd(bad, 1ull);
}

extern "C" int gsa_index_selfcheck(gsa_ctx *ctx, int64_t n_samples, int64_t *n_bad)
{
	if (!ctx || !n_bad || n_samples <= 0) return GSA_ERR_ARG;
	if (!ctx->have_index) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_index_selfcheck: no index");
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	GSA_TRY(gsa_ensure(ctx, ctx->d_counter, 1024));
	unsigned long long *d_bad = (unsigned long long *)ctx->d_counter.p + 100;
	CUDA_TRY(ctx, cudaMemsetAsync(d_bad, 0, 8, ctx->stream));
	if (ctx->ix.wide) k_index_selfcheck<true><<<gsa_grid(n_samples, 256), 256, 0, ctx->stream>>>(ctx->ix, (uint64_t)n_samples, d_bad);
	else k_index_selfcheck<false><<<gsa_grid(n_samples, 256), 256, 0, ctx->stream>>>(ctx->ix, (uint64_t)n_samples, d_bad);
	KERNEL_CHECK(ctx);
	unsigned long long h = 0;
	CUDA_TRY(ctx, cudaMemcpyAsync(&h, d_bad, 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	*n_bad = (int64_t)h;
	return GSA_OK;
}
