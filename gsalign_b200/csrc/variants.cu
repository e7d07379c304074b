// variants.cu -- N3: VariantIdentification (reference src/SeqVariant.cpp:12-119) on the device.
//
// The reference walks every gap fragment of every block column by column on the host and pushes one Variant_t per
// substitution, insertion run and deletion run.  Here the walk runs where the rows already are: a count pass, an exclusive
// scan over the fragments (scan order = push order: fragments of a block are stored in order, columns left to right) and a
// write pass that emits one 24-byte record per variant.  Alleles are not materialised: they are substrings of the query and
// of the reference text at the record's coordinates, which the host formats straight into the VCF line.
//
// Per fragment (SeqVariant.cpp:25-110): seeds and empty fragments yield nothing; qLen == 0 / rLen == 0 one deletion /
// insertion; a 1 x 1 fragment one substitution if the two bases differ by nst_nt4_table class and the query base is ACGT;
// anything else is scanned: a run of '-' in the reference row is one insertion, a run of '-' in the query row one deletion,
// a column of two letters of different class with an ACGT query letter one substitution.  Threads take a fragment each; the
// fragments that need the scan are then worked off by the whole warp, 32 columns per step (positions from ballots + popc).
#include <cub/cub.cuh>
#include "fm.cuh"

#define VAR_FULL 0xffffffffu

struct VarArgs {
	const gsa_frag *frag; int64_t nfr;
	const char *a1, *a2;
	const ContigEnd *ce; int nce;
	int64_t N;
};

// GenCoordinateInfo(rPos).gPos (reference src/tools.cpp:120-140): ChrLocMap.lower_bound, then the 1-based offset from the
// contig's start (forward half of T) or from its end (mirrored half)
__device__ __forceinline__ int var_gpos(const VarArgs &A, int64_t rpos)
{
	int lo = 0, hi = A.nce;
	while (lo < hi) { int m = (lo + hi) >> 1; if (A.ce[m].end < rpos) lo = m + 1; else hi = m; }
	if (lo == A.nce) lo = A.nce - 1;
	const ContigEnd e = A.ce[lo];
	return rpos < A.N ? (int)(rpos - e.end + e.pad) : (int)(e.end - rpos + 1);
}

__device__ __forceinline__ void var_put(const VarArgs &A, gsa_variant *out, int64_t at, int kind, int64_t rpos, int qpos, int len)
{
	gsa_variant v;
	v.rPos = rpos; v.qPos = qpos; v.gPos = var_gpos(A, rpos); v.len = len; v.kind = kind;
	out[at] = v;
}

template <bool WRITE>
__global__ void __launch_bounds__(256) k_variants(VarArgs A, int64_t *cnt, const int64_t *off, gsa_variant *out)
{
	const int lane = threadIdx.x & 31;
	const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	gsa_frag f;
	f.rPos = 0; f.qPos = 0; f.qLen = 0; f.rLen = 0; f.bSeed = 1; f.aln_off = 0; f.aln_len = 0;
	if (t < A.nfr) f = A.frag[t];
	const int64_t o = (WRITE && t < A.nfr) ? off[t] : 0;
	int n = 0;
	bool walk = false;
	if (!f.bSeed && (f.qLen | f.rLen) != 0) {
		if (f.qLen == 0) { n = 1; if (WRITE) var_put(A, out, o, GSA_VAR_FRAG_DEL, f.rPos - 1, f.qPos - 1, f.rLen); }
		else if (f.rLen == 0) { n = 1; if (WRITE) var_put(A, out, o, GSA_VAR_FRAG_INS, f.rPos - 1, f.qPos - 1, f.qLen); }
		else if (f.qLen == 1 && f.rLen == 1) {
			const int x = gsa_nt4((unsigned char)A.a1[f.aln_off]), y = gsa_nt4((unsigned char)A.a2[f.aln_off]);
			if (x != y && y != 4) { n = 1; if (WRITE) var_put(A, out, o, GSA_VAR_SNV, f.rPos, f.qPos, 1); }
		} else walk = true;
	}
	unsigned todo = __ballot_sync(VAR_FULL, walk);
	while (todo) {
		const int src = __ffs(todo) - 1;
		todo &= todo - 1;
		const int64_t rPos = __shfl_sync(VAR_FULL, f.rPos, src), aoff = __shfl_sync(VAR_FULL, f.aln_off, src), obase = __shfl_sync(VAR_FULL, o, src);
		const int qPos = __shfl_sync(VAR_FULL, f.qPos, src), alen = __shfl_sync(VAR_FULL, f.aln_len, src);
		const char *r1 = A.a1 + aoff, *r2 = A.a2 + aoff;
		const unsigned lt = (1u << lane) - 1;
		int nv = 0, br = 0, bq = 0;   // variants, reference bases and query bases before this window
		for (int w = 0; w < alen; w += 32) {
			const int i = w + lane;
			const bool in = i < alen;
			const char c1 = in ? r1[i] : 'A', c2 = in ? r2[i] : 'A';
			char p1 = (char)__shfl_up_sync(VAR_FULL, (int)c1, 1), p2 = (char)__shfl_up_sync(VAR_FULL, (int)c2, 1);
			if (lane == 0) { p1 = w > 0 ? r1[w - 1] : 'A'; p2 = w > 0 ? r2[w - 1] : 'A'; }
			const bool g1 = in && c1 == '-', g2 = in && c2 == '-';
			const unsigned m1 = __ballot_sync(VAR_FULL, g1), m2 = __ballot_sync(VAR_FULL, g2), mi = __ballot_sync(VAR_FULL, in);
			int kind = -1;
			if (g1) { if (p1 != '-') kind = GSA_VAR_INS; }          // first column of a run of '-' in the reference row
			else if (g2) { if (p2 != '-') kind = GSA_VAR_DEL; }     // ... in the query row
			else if (in) { const int x = gsa_nt4((unsigned char)c1), y = gsa_nt4((unsigned char)c2); if (x != y && y != 4) kind = GSA_VAR_SNV; }
			const unsigned vm = __ballot_sync(VAR_FULL, kind >= 0);
			if (WRITE && kind >= 0) {
				const int r = br + __popc(mi & ~m1 & lt), q = bq + __popc(mi & ~m2 & lt);
				int len = 1;
				if (kind == GSA_VAR_INS) while (i + len < alen && r1[i + len] == '-') len++;
				if (kind == GSA_VAR_DEL) while (i + len < alen && r2[i + len] == '-') len++;
				const int back = kind == GSA_VAR_SNV ? 0 : 1;   // indels are anchored on the base before the gap
				var_put(A, out, obase + nv + __popc(vm & lt), kind, rPos + r - back, qPos + q - back, len);
			}
			nv += __popc(vm); br += __popc(mi & ~m1); bq += __popc(mi & ~m2);
		}
		if (lane == src) n = nv;
	}
	if (!WRITE && t < A.nfr) cnt[t] = n;
}

// first record and number of records of every output block: its fragments are one contiguous range
__global__ void k_var_ranges(const int64_t *off, const int64_t *beg_end, int nb, int64_t *out)
{
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= nb) return;
	const int64_t a = off[beg_end[2 * k]], b = off[beg_end[2 * k + 1]];
	out[k] = a; out[nb + k] = b - a;
}

int gsa_impl_variants(gsa_ctx *ctx, gsa_variant_list *out)
{
	memset(out, 0, sizeof(*out));
	const int nb = (int)ctx->out_blocks.size();
	const int64_t nfr = ctx->n_frags;
	if (nb == 0 || nfr == 0) return GSA_OK;
	// scratch: slots of the fill phase that are dead once the rows exist
	DevBuf &b_cnt = ctx->d_tmp[0], &b_off = ctx->d_tmp[1], &b_rng = ctx->d_tmp[2];
	GSA_TRY(gsa_ensure(ctx, b_cnt, (size_t)(nfr + 1) * 8));
	GSA_TRY(gsa_ensure(ctx, b_off, (size_t)(nfr + 1) * 8));
	GSA_TRY(gsa_ensure(ctx, b_rng, (size_t)nb * 32 + 16));
	int64_t *cnt = (int64_t *)b_cnt.p, *off = (int64_t *)b_off.p, *d_be = (int64_t *)b_rng.p, *d_rng = d_be + 2 * (size_t)nb;
	GSA_TRY(gsa_ensure_host(ctx, ctx->h_vrange, (size_t)nb * 32 + 16));
	int64_t *h_be = (int64_t *)ctx->h_vrange.p, *h_rng = h_be + 2 * (size_t)nb;
	for (int k = 0; k < nb; k++) { h_be[2 * k] = ctx->out_blocks[k].frag_beg; h_be[2 * k + 1] = ctx->out_blocks[k].frag_beg + ctx->out_blocks[k].n_frags; }
	VarArgs A;
	A.frag = (const gsa_frag *)ctx->d_frag.p; A.nfr = nfr; A.a1 = (const char *)ctx->d_aln1.p; A.a2 = (const char *)ctx->d_aln2.p;
	A.ce = (const ContigEnd *)ctx->d_cend.p; A.nce = (int)ctx->cend.size(); A.N = ctx->N;
	CUDA_TRY(ctx, cudaMemcpyAsync(d_be, h_be, (size_t)nb * 16, cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(cnt + nfr, 0, 8, ctx->stream));
	k_variants<false><<<gsa_grid(nfr, 256), 256, 0, ctx->stream>>>(A, cnt, nullptr, nullptr);
	KERNEL_CHECK(ctx);
	{
		size_t bytes = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt, off, nfr + 1, ctx->stream);
		GSA_TRY(gsa_ensure(ctx, ctx->d_cub, bytes));
		CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ctx->d_cub.p, bytes, cnt, off, nfr + 1, ctx->stream));
		ctx->tm.launches++;
	}
	k_var_ranges<<<gsa_grid(nb, 128), 128, 0, ctx->stream>>>(off, d_be, nb, d_rng);
	KERNEL_CHECK(ctx);
	int64_t *h_total = h_rng + 2 * (size_t)nb;   // the 16 spare bytes of h_vrange
	CUDA_TRY(ctx, cudaMemcpyAsync(h_rng, d_rng, (size_t)nb * 16, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaMemcpyAsync(h_total, off + nfr, 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	const int64_t total = *h_total;
	out->block_first = h_rng; out->block_count = h_rng + nb;
	out->n_variants = total;
	if (total == 0) return GSA_OK;
	GSA_TRY(gsa_ensure(ctx, ctx->d_var, (size_t)total * sizeof(gsa_variant)));
	GSA_TRY(gsa_ensure_host(ctx, ctx->h_var, (size_t)total * sizeof(gsa_variant)));
	k_variants<true><<<gsa_grid(nfr, 256), 256, 0, ctx->stream>>>(A, nullptr, off, (gsa_variant *)ctx->d_var.p);
	KERNEL_CHECK(ctx);
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_var.p, ctx->d_var.p, (size_t)total * sizeof(gsa_variant), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	out->variants = (const gsa_variant *)ctx->h_var.p;
	return GSA_OK;
}
