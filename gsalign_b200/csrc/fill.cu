// fill.cu -- K3: gapped fill of the fragments between seeds.
//
// Replaces GenerateFragAlignment (reference src/ProcessCandidateAlignment.cpp:290-351) and
// ksw2_alignment / ksw_extz2_sse / ksw_backtrack (src/ksw2_alignment.cpp:25-273).
//   * dispatch per non-seed fragment: pure gap rows, ungapped copy when qLen == rLen and <= 5 mismatches
//     (query-N positions skipped), otherwise a GLOBAL affine-gap alignment over the full matrix
//     (the reference passes w = -1, so no band is bit-exact).
//   * DP recurrence (absolute scores, validated against the reference by the oracle tests): rows over the
//     query fragment, columns over the reference fragment; match +1, mismatch -1, any non-ACGT 0;
//     E(i,j) = max(H(i-1,j)-3, E(i-1,j)-1), F(i,j) = max(H(i,j-1)-3, F(i,j-1)-1), ties open;
//     H = max(diag+s, E, F) preferring diag, then E, then F; scores fit int16 (|H| <= 2+m+n).
//   * the DP itself is the packed-int16 DPX wavefront kernel k_dpx (dpx.cu); the problems are binned by size class (and
//     by whether they hold letters outside ACGT) on the device (run_dp_binned) and only the per-bin counts come back
//     to the host.
// Also accumulates AlnBlock_t::aln_len / score per block and applies nothing else: the identity filter and
// the final block order are O(#blocks) host logic (see gsa_impl_fill at the bottom).
#include "dpx.cuh"
#include "fm.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <algorithm>
#include <stdlib.h>

// GSA_NO_PACK=1 sends every DP problem through the one-warp-per-problem kernels (A/B measurements)
static int dpx_use_pack()
{
	static const int v = getenv("GSA_NO_PACK") ? 0 : 1;
	return v;
}

// fragment type codes
// (values are part of the record format: gsa_frag::reserved, gather.cu)
enum { FT_SEED = 0, FT_DEL = 1, FT_INS = 2, FT_COPY = 3, FT_DP = 4 };

// ---- classification ----------------------------------------------------------------------------------
// One thread per fragment: type, mismatch count for equal-length fragments (CheckFragPairMismatch,
// src/ProcessCandidateAlignment.cpp:49-61), upper bound of its row length, DP cell count.
__global__ void k_frag_classify(gsa_frag *frag, int64_t nfr, const unsigned char *seq, const uint32_t *qinv, DevIndex ix, uint8_t *type, int32_t *mism,
                                int64_t *row_len, int64_t *flag_len, uint8_t *is_dp, uint8_t *dp_cls, int use_pack)
{
	int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nfr) return;
	gsa_frag f = frag[t];
	uint8_t ty = FT_SEED, cls = 0; int64_t rl = 0, fl = 0; int mm = 0;
	if (!f.bSeed) {
		if (f.qLen == 0) { ty = FT_DEL; rl = f.rLen; }
		else if (f.rLen == 0) { ty = FT_INS; rl = f.qLen; }
		else {
			ty = FT_DP;
			if (f.qLen == f.rLen) {
				for (int k = 0; k < f.qLen && mm <= 5; k++) {
					int b = gsa_nt4(seq[f.qPos + k]);
					if (b != 4 && b != gsa_pk_base(ix.txt, (uint64_t)(f.rPos + k))) mm++;
				}
				if (mm <= 5) ty = FT_COPY;
			}
			if (ty == FT_COPY) rl = f.qLen;
			else {
				rl = (int64_t)f.qLen + f.rLen;
				bool other = false; // any non-ACGT query base in the fragment (the 2-bit reference text holds none)
				for (uint32_t p = (uint32_t)f.qPos, e = p + (uint32_t)f.qLen; p < e && !other; p += 32) other = (gsa_bit_window(qinv, p) >> (32 - min(32u, e - p))) != 0;
				cls = (uint8_t)dpx_class(f.rLen, f.qLen, other, use_pack != 0);
				fl = dpx_flag_bytes(f.rLen, f.qLen, cls);
			}
		}
	}
	type[t] = ty; mism[t] = mm; row_len[t] = rl; flag_len[t] = fl; is_dp[t] = ty == FT_DP; dp_cls[t] = cls;
	if (ty != FT_SEED) frag[t].reserved = ty;   // gsa_frag::reserved: how the fragment's rows are produced (its row slot follows from it)
}

// rows of the non-DP fragment types + per-block sums (src/ProcessCandidateAlignment.cpp:303-331)
__global__ void k_frag_simple(gsa_frag *frag, int64_t nfr, const int32_t *fblk, const uint8_t *type, const int32_t *mism, const int64_t *row_off,
                              const unsigned char *seq, DevIndex ix, char *aln1, char *aln2, unsigned int *bsum)
{
	int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	unsigned int alen = 0, sc = 0;
	int b = -1 - (int)(threadIdx.x & 31); // lanes with nothing to add form singleton groups
	if (t < nfr) {
		gsa_frag f = frag[t];
		int ty = type[t];
		int64_t o = row_off[t];
		if (ty != FT_DP) b = fblk[t];     // DP fragments are accounted by the DP kernels
		if (ty == FT_SEED) { alen = (unsigned)f.qLen; sc = (unsigned)f.qLen; }
		else if (ty == FT_DEL) {
			for (int k = 0; k < f.rLen; k++) { aln1[o + k] = gsa_text_char(ix, f.rPos + k); aln2[o + k] = '-'; }
			alen = (unsigned)f.rLen; frag[t].aln_off = o; frag[t].aln_len = f.rLen;
		} else if (ty == FT_INS) {
			for (int k = 0; k < f.qLen; k++) { aln1[o + k] = '-'; aln2[o + k] = (char)seq[f.qPos + k]; }
			alen = (unsigned)f.qLen; frag[t].aln_off = o; frag[t].aln_len = f.qLen;
		} else if (ty == FT_COPY) {
			for (int k = 0; k < f.qLen; k++) { aln1[o + k] = gsa_text_char(ix, f.rPos + k); aln2[o + k] = (char)seq[f.qPos + k]; }
			alen = (unsigned)f.qLen; sc = (unsigned)(f.qLen - mism[t]); frag[t].aln_off = o; frag[t].aln_len = f.qLen;
		}
	}
	// per-block sums: fragments are in block order, so a warp usually feeds one block (peers summed in registers), and so does
	// the whole CTA: the warps' sums for the CTA's first block meet in shared memory and leave as ONE pair of global atomics.
	// (A collinear contig is one block: every warp of the grid adding to the same two words serialises in the L2.)
	__shared__ int s_blk;
	__shared__ unsigned int s_len, s_sc;
	if (threadIdx.x == 0) { s_blk = -1; s_len = 0; s_sc = 0; }
	__syncthreads();
	if (threadIdx.x == 0 && t < nfr) s_blk = fblk[t];
	__syncthreads();
	unsigned peers = __match_any_sync(0xffffffffu, b);
	bool leader;
	unsigned long long both = gsa_peer_sum(peers, ((unsigned long long)alen << 32) | sc, leader); // a warp adds < 2^32 to either half
	if (leader && b >= 0) {
		if (b == s_blk) { atomicAdd(&s_len, (unsigned)(both >> 32)); atomicAdd(&s_sc, (unsigned)both); }
		else { atomicAdd(bsum + 2 * b, (unsigned)(both >> 32)); atomicAdd(bsum + 2 * b + 1, (unsigned)both); }
	}
	__syncthreads();
	if (threadIdx.x == 0 && s_blk >= 0 && (s_len | s_sc)) { atomicAdd(bsum + 2 * s_blk, s_len); atomicAdd(bsum + 2 * s_blk + 1, s_sc); }
}

__global__ void k_dp_problems(const int32_t *dp_idx, int64_t ndp, const gsa_frag *frag, const int64_t *row_off, const int64_t *flag_off,
                              const uint8_t *dp_cls, const unsigned char *seq, DpProblem *prob)
{
	int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= ndp) return;
	int32_t t = dp_idx[k];
	gsa_frag f = frag[t];
	DpProblem p; p.ref_chars = nullptr; p.qry_chars = (const char *)seq + f.qPos; p.rpos = f.rPos; p.flag_off = flag_off[t]; p.out_off = row_off[t];
	p.m = f.rLen; p.n = f.qLen; p.frag = t; p.cls = dp_cls[t];
	prob[k] = p;
}

struct Ws3 { // scratch slots of the context, handed out in order
	gsa_ctx *ctx; int next = 0; int rc = GSA_OK;
	explicit Ws3(gsa_ctx *c) : ctx(c) {}
	template <typename T> T *get(int64_t n)
	{
		if (next >= 64) { rc = gsa_fail(ctx, GSA_ERR_NOMEM, "fill: out of scratch slots"); return nullptr; }
		DevBuf &b = ctx->d_tmp[next++];
		int r = gsa_ensure(ctx, b, (size_t)(n > 0 ? n : 1) * sizeof(T) + 64);
		if (r != GSA_OK) { rc = r; return nullptr; }
		return (T *)b.p;
	}
};

// ---- device-side binning ------------------------------------------------------------------------------------------------
// Problems are ordered by (bin, size descending) with one radix sort; only the per-bin counts come back to the host.
// Bins = the classes of dpx.cuh (size class x {ACGT only, other letters}), one launch each; last bin: too long.
#define DP_BIN_TOOLONG DPX_NCLS
#define DP_NBINS (DP_BIN_TOOLONG + 1)
struct DpStats { unsigned int count[DP_NBINS]; int max_m[DPX_NCLS], max_n[DPX_NCLS]; unsigned long long cells; };

__global__ void k_dp_keys(const DpProblem *prob, int n, uint32_t *key, uint32_t *idx, DpStats *st)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	int bin = -1 - (int)(threadIdx.x & 31), m = 0, q = 0;
	if (i < n) {
		m = prob[i].m; q = prob[i].n;
		int cls = prob[i].cls, big = max(m, q);
		bin = (big > DP_MAX_DIM || cls < 0 || cls >= DPX_NCLS) ? DP_BIN_TOOLONG : cls;
		// inside a bin: largest first; the pack classes run warp-uniform loops over the problems that share a warp, so they are
		// ordered by rows, then columns
		const bool pack = bin < DP_BIN_TOOLONG && (bin % DPX_NSIZE) >= DPX_CLS_P1;
		uint32_t sz = pack ? (uint32_t)((min(q, 127) << 4) | min(15, m >> 4)) : (uint32_t)min(2047, (m + q) >> 3);
		key[i] = ((uint32_t)bin << 11) | (2047u - sz);
		idx[i] = (uint32_t)i;
	}
	// per-bin statistics: one atomic per distinct bin of the warp (the counters are few and every thread hits them)
	unsigned peers = __match_any_sync(0xffffffffu, bin);
	bool leader;
	unsigned long long cells = gsa_peer_sum(peers, (unsigned long long)m * (unsigned long long)q, leader);
	int cnt = gsa_peer_sum(peers, 1, leader);
	int mm = m, mq = q;
	for (unsigned rest = peers; rest; rest &= rest - 1) { int src = __ffs(rest) - 1; mm = max(mm, __shfl_sync(peers, m, src)); mq = max(mq, __shfl_sync(peers, q, src)); }
	if (leader && bin >= 0) {
		atomicAdd(&st->count[bin], (unsigned)cnt);
		atomicAdd(&st->cells, cells);
		if (bin < DP_BIN_TOOLONG) { atomicMax(&st->max_m[bin], mm); atomicMax(&st->max_n[bin], mq); }
	}
}

__global__ void k_dp_permute(const DpProblem *in, const uint32_t *idx, DpProblem *out, int n)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = in[idx[i]];
}

// d_prob: the problems in any order (device); d_sorted: scratch of the same size.  Records ctx->ev[10]/[11] around the DP
// launches alone when timed.
static int run_dp_binned(gsa_ctx *ctx, Ws3 &ws, const DpProblem *d_prob, DpProblem *d_sorted, int ndp, uint8_t *flags, char *a1, char *a2, int32_t *out_len,
                         int64_t *out_start, gsa_frag *frag, const int32_t *fblk, unsigned int *bsum, cudaEvent_t e0, cudaEvent_t e1)
{
	uint32_t *key_in = ws.get<uint32_t>(ndp), *key_out = ws.get<uint32_t>(ndp), *idx_in = ws.get<uint32_t>(ndp), *idx_out = ws.get<uint32_t>(ndp);
	if (ws.rc) return ws.rc;
	GSA_TRY(gsa_ensure(ctx, ctx->d_counter, 1024));
	DpStats *d_st = (DpStats *)((char *)ctx->d_counter.p + 512);
	CUDA_TRY(ctx, cudaMemsetAsync(d_st, 0, sizeof(DpStats), ctx->stream));
	k_dp_keys<<<gsa_grid(ndp, 256), 256, 0, ctx->stream>>>(d_prob, ndp, key_in, idx_in, d_st);
	KERNEL_CHECK(ctx);
	size_t bytes = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, bytes, key_in, key_out, idx_in, idx_out, ndp, 0, 16, ctx->stream);
	GSA_TRY(gsa_ensure(ctx, ctx->d_cub, bytes));
	CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->d_cub.p, bytes, key_in, key_out, idx_in, idx_out, ndp, 0, 16, ctx->stream));
	ctx->tm.launches += 3;
	k_dp_permute<<<gsa_grid(ndp, 256), 256, 0, ctx->stream>>>(d_prob, idx_out, d_sorted, ndp);
	KERNEL_CHECK(ctx);
	DpStats *st = (DpStats *)((char *)ctx->h_small.p + 256);
	GSA_TRY(gsa_small_d2h(ctx, st, d_st, sizeof(DpStats)));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->tm.dp_cells = (int64_t)st->cells;
	if (st->count[DP_BIN_TOOLONG]) return gsa_fail(ctx, GSA_ERR_LIMIT, "DP fragment longer than %d", DP_MAX_DIM);
	if (e0) CUDA_TRY(ctx, cudaEventRecord(e0, ctx->stream));
	// one launch per non-empty class; the classes are independent, so they are spread over the context's side streams and
	// run next to each other (the few big problems of G8 / G16 have long critical paths, the small classes fill the SMs)
	size_t off[DP_NBINS + 1]; off[0] = 0;
	for (int b = 0; b < DP_NBINS; b++) off[b + 1] = off[b] + st->count[b];
	const int order[] = {DPX_CLS_G16, DPX_CLS_G8, DPX_CLS_G4, DPX_CLS_S2, DPX_CLS_S1, DPX_CLS_P16, DPX_CLS_P8, DPX_CLS_P4, DPX_CLS_P2, DPX_CLS_P1};
	int nlaunch = 0;
	for (int c = 0; c < DPX_NCLS; c++) nlaunch += st->count[c] > 0;
	const bool side = nlaunch > 1;
	if (side) { CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream)); for (int i = 0; i < GSA_NSIDE; i++) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->side[i], ctx->ev_fork, 0)); }
	int turn = 0;
	for (int size : order)
		for (int hasn = 0; hasn < 2; hasn++) {
			const int cls = size + hasn * DPX_NSIZE;
			if (st->count[cls] == 0) continue;
			cudaStream_t st_cls = !side ? ctx->stream : (turn % (GSA_NSIDE + 1) == GSA_NSIDE ? ctx->stream : ctx->side[turn % (GSA_NSIDE + 1)]);
			turn++;
			GSA_TRY(gsa_dpx_launch(ctx, st_cls, cls, std::max(1, st->max_m[cls]), std::max(1, st->max_n[cls]), d_sorted + off[cls], (int)st->count[cls], flags, a1, a2, out_len, out_start, frag, fblk, bsum));
		}
	if (side) for (int i = 0; i < GSA_NSIDE; i++) { CUDA_TRY(ctx, cudaEventRecord(ctx->ev_side[i], ctx->side[i])); CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_side[i], 0)); }
	if (e1) CUDA_TRY(ctx, cudaEventRecord(e1, ctx->stream));
	return GSA_OK;
}

static int scan_ex64(gsa_ctx *ctx, const int64_t *in, int64_t *out, int64_t n)
{
	size_t bytes = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, ctx->stream);
	GSA_TRY(gsa_ensure(ctx, ctx->d_cub, bytes));
	CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ctx->d_cub.p, bytes, in, out, (int)n, ctx->stream));
	ctx->tm.launches++;
	return GSA_OK;
}

int gsa_impl_fill(gsa_ctx *ctx, gsa_alignment *out)
{
	memset(out, 0, sizeof(*out));
	ctx->out_blocks.clear();
	int nblk = (int)ctx->final_blocks.size();
	int64_t nfr = ctx->n_frags;
	ctx->tm.n_frags = nfr;
	if (nblk == 0 || nfr == 0) return GSA_OK;
	Ws3 ws(ctx);
	gsa_frag *frag = (gsa_frag *)ctx->d_frag.p; const int32_t *fblk = (const int32_t *)ctx->d_fblk.p;
	const unsigned char *seq = (const unsigned char *)ctx->d_seq.p;
	uint8_t *type = ws.get<uint8_t>(nfr), *is_dp = ws.get<uint8_t>(nfr), *dp_cls = ws.get<uint8_t>(nfr);
	int32_t *mism = ws.get<int32_t>(nfr), *dp_idx = ws.get<int32_t>(nfr);
	int64_t *row_len = ws.get<int64_t>(nfr + 1), *flag_len = ws.get<int64_t>(nfr + 1), *row_off = ws.get<int64_t>(nfr + 1), *flag_off = ws.get<int64_t>(nfr + 1);
	if (ws.rc) return ws.rc;
	GSA_TRY(gsa_ensure(ctx, ctx->d_bsum, (size_t)nblk * 8));
	unsigned int *bsum = (unsigned int *)ctx->d_bsum.p;
	int32_t *d_ndp = (int32_t *)ctx->d_counter.p + 32;
	CUDA_TRY(ctx, cudaMemsetAsync(bsum, 0, (size_t)nblk * 8, ctx->stream));
	k_frag_classify<<<gsa_grid(nfr, 128), 128, 0, ctx->stream>>>(frag, nfr, seq, (const uint32_t *)ctx->d_qinv.p, ctx->ix, type, mism, row_len, flag_len, is_dp, dp_cls, dpx_use_pack());
	KERNEL_CHECK(ctx);
	CUDA_TRY(ctx, cudaMemsetAsync(row_len + nfr, 0, 8, ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(flag_len + nfr, 0, 8, ctx->stream));
	GSA_TRY(scan_ex64(ctx, row_len, row_off, nfr + 1));
	GSA_TRY(scan_ex64(ctx, flag_len, flag_off, nfr + 1));
	{
		size_t bytes = 0;
		thrust::counting_iterator<int32_t> it(0);
		cub::DeviceSelect::Flagged(nullptr, bytes, it, is_dp, dp_idx, d_ndp, (int)nfr, ctx->stream);
		GSA_TRY(gsa_ensure(ctx, ctx->d_cub, bytes));
		CUDA_TRY(ctx, cub::DeviceSelect::Flagged(ctx->d_cub.p, bytes, it, is_dp, dp_idx, d_ndp, (int)nfr, ctx->stream));
		ctx->tm.launches += 2;
	}
	int64_t *hs = (int64_t *)ctx->h_small.p;
	GSA_TRY(gsa_small_d2h(ctx, hs, row_off + nfr, 8));
	GSA_TRY(gsa_small_d2h(ctx, hs + 1, flag_off + nfr, 8));
	GSA_TRY(gsa_small_d2h(ctx, hs + 2, d_ndp, 4));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	int64_t row_bytes = hs[0], flag_bytes = hs[1], ndp = *(int32_t *)(hs + 2);
	ctx->aln_bytes = row_bytes;
	GSA_TRY(gsa_ensure(ctx, ctx->d_aln1, (size_t)row_bytes + 16));
	GSA_TRY(gsa_ensure(ctx, ctx->d_aln2, (size_t)row_bytes + 16));
	char *a1 = (char *)ctx->d_aln1.p, *a2 = (char *)ctx->d_aln2.p;
	k_frag_simple<<<gsa_grid(nfr, 256), 256, 0, ctx->stream>>>(frag, nfr, fblk, type, mism, row_off, seq, ctx->ix, a1, a2, bsum);
	KERNEL_CHECK(ctx);
	ctx->tm.n_dp = ndp; ctx->tm.dp_cells = 0;
	if (ndp > 0) {
		DpProblem *d_prob = ws.get<DpProblem>(ndp);
		uint8_t *flags = ws.get<uint8_t>(flag_bytes + 256);
		if (ws.rc) return ws.rc;
		k_dp_problems<<<gsa_grid(ndp, 128), 128, 0, ctx->stream>>>(dp_idx, ndp, frag, row_off, flag_off, dp_cls, seq, d_prob);
		KERNEL_CHECK(ctx);
		DpProblem *d_sorted = ws.get<DpProblem>(ndp);
		if (ws.rc) return ws.rc;
		GSA_TRY(run_dp_binned(ctx, ws, d_prob, d_sorted, (int)ndp, flags, a1, a2, nullptr, nullptr, frag, fblk, bsum, ctx->ev[10], ctx->ev[11]));
		ctx->dp_timed = true;
	}
	// ---- results to pinned host memory (skipped when the consumer reads them on the device, gsa_set_host_results) ---------
	GSA_TRY(gsa_ensure_host(ctx, ctx->h_blocks, (size_t)nblk * 8));
	if (ctx->host_results) {
		GSA_TRY(gsa_ensure_host(ctx, ctx->h_frag, (size_t)nfr * sizeof(gsa_frag)));
		GSA_TRY(gsa_ensure_host(ctx, ctx->h_aln1, (size_t)row_bytes + 16));
		GSA_TRY(gsa_ensure_host(ctx, ctx->h_aln2, (size_t)row_bytes + 16));
		GSA_TRY(gsa_bulk_copy(ctx, ctx->h_frag.p, frag, (size_t)nfr * sizeof(gsa_frag), cudaMemcpyDeviceToHost, ctx->stream));
		if (row_bytes) {
			GSA_TRY(gsa_bulk_copy(ctx, ctx->h_aln1.p, a1, (size_t)row_bytes, cudaMemcpyDeviceToHost, ctx->stream));
			GSA_TRY(gsa_bulk_copy(ctx, ctx->h_aln2.p, a2, (size_t)row_bytes, cudaMemcpyDeviceToHost, ctx->stream));
		}
	}
	GSA_TRY(gsa_small_d2h(ctx, ctx->h_blocks.p, bsum, (size_t)nblk * 8));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	// ---- identity filter + final order (src/GSAlign.cpp:529-540): O(#blocks), same std::sort as the reference
	const unsigned int *hb = (const unsigned int *)ctx->h_blocks.p;
	std::vector<BlockHdr> vec = ctx->final_blocks;
	for (int k = 0; k < nblk; k++) {
		vec[k].aln_len = (int32_t)hb[2 * k]; vec[k].score = (int32_t)hb[2 * k + 1];
		if ((int)(100 * (1.0 * vec[k].score / vec[k].aln_len)) < ctx->prm.min_idy) vec[k].score = 0;
	}
	gsa_host_remove_bad(vec);
	ctx->out_blocks.resize(vec.size());
	for (size_t k = 0; k < vec.size(); k++) {
		gsa_block &b = ctx->out_blocks[k];
		b.score = vec[k].score; b.aln_len = vec[k].aln_len; b.bDup = vec[k].bDup; b.n_frags = vec[k].n_frags; b.frag_beg = vec[k].frag_beg;
	}
	out->n_blocks = (int32_t)ctx->out_blocks.size(); out->blocks = ctx->out_blocks.data();
	out->n_frags = nfr; out->aln_bytes = row_bytes;
	if (ctx->host_results) { out->frags = (const gsa_frag *)ctx->h_frag.p; out->aln1 = (const char *)ctx->h_aln1.p; out->aln2 = (const char *)ctx->h_aln2.p; }
	return GSA_OK;
}

// batch mode: rows sit wherever the kernels left them inside each pair's slot; move them to the start of the slot
__global__ void k_batch_left_align(const DpProblem *prob, int n, const int32_t *len, const int64_t *start, const char *t1, const char *t2, char *o1, char *o2)
{
	int pi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if (pi >= n) return;
	int64_t dst = prob[pi].out_off, src = start[pi];
	for (int k = lane; k < len[pi]; k += 32) { o1[dst + k] = t1[src + k]; o2[dst + k] = t2[src + k]; }
}

// ---- stand-alone DP batch (dump hook / DP stress bench) -----------------------------------------------------
int gsa_impl_dp_batch(gsa_ctx *ctx, int32_t n_pairs, const char *ref, const int64_t *ref_off, const char *qry,
                      const int64_t *qry_off, char *out1, char *out2, int32_t *out_len, int32_t *out_identical, float *kernel_ms)
{
	if (kernel_ms) *kernel_ms = 0;
	if (n_pairs == 0) return GSA_OK;
	if (!ref || !qry || !ref_off || !qry_off || !out1 || !out2 || !out_len) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_dp_batch: null argument");
	Ws3 ws(ctx);
	int64_t rb = ref_off[n_pairs], qb = qry_off[n_pairs];
	char *d_ref = ws.get<char>(rb + 1), *d_qry = ws.get<char>(qb + 1), *d_o1 = ws.get<char>(rb + qb + 1), *d_o2 = ws.get<char>(rb + qb + 1);
	char *d_t1 = ws.get<char>(rb + qb + 1), *d_t2 = ws.get<char>(rb + qb + 1);
	int32_t *d_len = ws.get<int32_t>(n_pairs);
	int64_t *d_start = ws.get<int64_t>(n_pairs);
	// with out_identical every pair is its own "block": the kernels' per-block sums then are the per-pair column counts
	gsa_frag *d_frag = out_identical ? ws.get<gsa_frag>(n_pairs) : nullptr;
	int32_t *d_fblk = out_identical ? ws.get<int32_t>(n_pairs) : nullptr;
	unsigned int *d_bsum = out_identical ? ws.get<unsigned int>(2 * (int64_t)n_pairs) : nullptr;
	DpProblem *d_prob = ws.get<DpProblem>(n_pairs), *d_sorted = ws.get<DpProblem>(n_pairs);
	if (ws.rc) return ws.rc;
	std::vector<DpProblem> hp((size_t)n_pairs);
	int64_t fbytes = 0;
	for (int i = 0; i < n_pairs; i++) {
		DpProblem &p = hp[i];
		p.m = (int32_t)(ref_off[i + 1] - ref_off[i]); p.n = (int32_t)(qry_off[i + 1] - qry_off[i]);
		if (p.m <= 0 || p.n <= 0) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_dp_batch: empty fragment in pair %d", i);
		p.ref_chars = d_ref + ref_off[i]; p.qry_chars = d_qry + qry_off[i]; p.rpos = 0; p.flag_off = fbytes; p.out_off = ref_off[i] + qry_off[i];
		if (std::max(p.m, p.n) > DP_MAX_DIM) return gsa_fail(ctx, GSA_ERR_LIMIT, "gsa_dp_batch: fragment longer than %d in pair %d", DP_MAX_DIM, i);
		bool other = false; // any letter outside ACGT/acgt: the HASN variant of the kernels
		for (int64_t k = ref_off[i]; k < ref_off[i + 1] && !other; k++) { char c = ref[k] & 0xDF; other = !(c == 'A' || c == 'C' || c == 'G' || c == 'T'); }
		for (int64_t k = qry_off[i]; k < qry_off[i + 1] && !other; k++) { char c = qry[k] & 0xDF; other = !(c == 'A' || c == 'C' || c == 'G' || c == 'T'); }
		p.frag = i; p.cls = dpx_class(p.m, p.n, other, dpx_use_pack() != 0);
		fbytes += dpx_flag_bytes(p.m, p.n, p.cls);
	}
	uint8_t *flags = ws.get<uint8_t>(fbytes + 256);
	if (ws.rc) return ws.rc;
	CUDA_TRY(ctx, cudaMemcpyAsync(d_ref, ref, (size_t)rb, cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(ctx, cudaMemcpyAsync(d_qry, qry, (size_t)qb, cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(ctx, cudaMemcpyAsync(d_prob, hp.data(), hp.size() * sizeof(DpProblem), cudaMemcpyHostToDevice, ctx->stream));
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	if (out_identical) {
		std::vector<int32_t> ident((size_t)n_pairs);
		for (int i = 0; i < n_pairs; i++) ident[(size_t)i] = i;
		CUDA_TRY(ctx, cudaMemcpyAsync(d_fblk, ident.data(), (size_t)n_pairs * 4, cudaMemcpyHostToDevice, ctx->stream));
		CUDA_TRY(ctx, cudaMemsetAsync(d_bsum, 0, (size_t)n_pairs * 8, ctx->stream));
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); // ident is a local
	}
	int rc = run_dp_binned(ctx, ws, d_prob, d_sorted, n_pairs, flags, d_t1, d_t2, d_len, d_start, d_frag, d_fblk, d_bsum, e0, e1);
	if (rc == GSA_OK) {
		k_batch_left_align<<<gsa_grid(n_pairs, 8), 256, 0, ctx->stream>>>(d_prob, n_pairs, d_len, d_start, d_t1, d_t2, d_o1, d_o2);
		KERNEL_CHECK(ctx);
		CUDA_TRY(ctx, cudaMemcpyAsync(out1, d_o1, (size_t)(rb + qb), cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(ctx, cudaMemcpyAsync(out2, d_o2, (size_t)(rb + qb), cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(ctx, cudaMemcpyAsync(out_len, d_len, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
		std::vector<unsigned int> hb;
		if (out_identical) { hb.resize(2 * (size_t)n_pairs); CUDA_TRY(ctx, cudaMemcpyAsync(hb.data(), d_bsum, hb.size() * 4, cudaMemcpyDeviceToHost, ctx->stream)); }
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		for (int i = 0; out_identical && i < n_pairs; i++) out_identical[i] = (int32_t)hb[2 * (size_t)i + 1];
		if (kernel_ms) cudaEventElapsedTime(kernel_ms, e0, e1);
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	return rc;
}
