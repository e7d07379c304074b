// cluster.cu -- K2: seed clustering / chaining on the device.
//
// Replaces, seed-level and data-parallel, what the reference does with per-group serial scans:
//   SeedGrouping                       src/GSAlign.cpp:126-143
//   SeedGroupAnalysis                  src/GSAlign.cpp:305-375  (+ RemoveOutlierSeeds :260-296, RefinePDFmap :245-258,
//                                      FindNeighboringPosDiffAvg :178-206, RemoveRedundantSeeds :208-225, AddAlnBlock :29-49)
//   RemoveOverlaps                     src/ProcessCandidateAlignment.cpp:189-231
//   CheckGapsBetweenSeeds              src/ProcessCandidateAlignment.cpp:120-139 (+ CalGapSimilarity, src/KmerAnalysis.cpp:32-121)
//   CheckAlnBlockSpanMultipleRefChrs   src/ProcessCandidateAlignment.cpp:81-99
//   IdentifyNormalPairs                src/ProcessCandidateAlignment.cpp:241-265
// following the exact restatement of SURVEY.md appendix B.  Everything is expressed as radix sorts
// (diagonal,qpos) / (group,qpos), prefix scans, stream compactions and elementwise kernels over ALL groups
// at once, so the one giant main-diagonal group of a collinear contig costs the same as many small ones:
//   * the greedy outlier windows (data-dependent resets) become "next window start" pointers computed
//     by binary search per seed and resolved by pointer jumping;
//   * the per-window PosDiff histograms become one global (window,bin) hash table with atomic counts;
//   * overlap trimming, gap / contig-span break points and normal-pair insertion are adjacent-pair maps.
// Only O(#blocks) headers go to the host (block_logic.cpp) for the reference's float/std::sort logic.
#include "fm.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <algorithm>

// ------------------------------------------------------------------------------------------------
// scratch management
// ------------------------------------------------------------------------------------------------
struct Ws {
	gsa_ctx *ctx; int next = 0; int rc = GSA_OK;
	explicit Ws(gsa_ctx *c) : ctx(c) {}
	template <typename T> T *get(int64_t n)
	{
		if (next >= 64) { rc = gsa_fail(ctx, GSA_ERR_NOMEM, "cluster: out of scratch slots"); return nullptr; }
		DevBuf &b = ctx->d_tmp[next++];
		int r = gsa_ensure(ctx, b, (size_t)(n > 0 ? n : 1) * sizeof(T) + 64);
		if (r != GSA_OK) { rc = r; return nullptr; }
		return (T *)b.p;
	}
};

template <typename T>
static int scan_inclusive(gsa_ctx *ctx, const T *in, T *out, int64_t n)
{
	if (n <= 0) return GSA_OK;
	size_t bytes = 0;
	cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, (int)n, ctx->stream);
	GSA_TRY(gsa_ensure(ctx, ctx->d_cub, bytes));
	CUDA_TRY(ctx, cub::DeviceScan::InclusiveSum(ctx->d_cub.p, bytes, in, out, (int)n, ctx->stream));
	ctx->tm.launches += 1;
	return GSA_OK;
}

template <typename T>
static int scan_exclusive(gsa_ctx *ctx, const T *in, T *out, int64_t n)
{
	if (n <= 0) return GSA_OK;
	size_t bytes = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, ctx->stream);
	GSA_TRY(gsa_ensure(ctx, ctx->d_cub, bytes));
	CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ctx->d_cub.p, bytes, in, out, (int)n, ctx->stream));
	ctx->tm.launches += 1;
	return GSA_OK;
}

// indices i in [0,n) with flags[i] != 0 -> out (ascending); the count is left at d_count (device)
static int select_indices(gsa_ctx *ctx, const uint8_t *flags, int32_t *out, int32_t *d_count, int64_t n)
{
	if (n <= 0) { CUDA_TRY(ctx, cudaMemsetAsync(d_count, 0, 4, ctx->stream)); return GSA_OK; }
	thrust::counting_iterator<int32_t> it(0);
	size_t bytes = 0;
	cub::DeviceSelect::Flagged(nullptr, bytes, it, flags, out, d_count, (int)n, ctx->stream);
	GSA_TRY(gsa_ensure(ctx, ctx->d_cub, bytes));
	CUDA_TRY(ctx, cub::DeviceSelect::Flagged(ctx->d_cub.p, bytes, it, flags, out, d_count, (int)n, ctx->stream));
	ctx->tm.launches += 2;
	return GSA_OK;
}

static int read_count(gsa_ctx *ctx, const int32_t *d_count, int64_t *out)
{
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_small.p, d_count, 4, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	*out = *(int32_t *)ctx->h_small.p;
	return GSA_OK;
}

#define LAUNCH(kernel, n, ...)                                                                     \
	do {                                                                                           \
		if ((n) > 0) { kernel<<<gsa_grid((n), 256), 256, 0, ctx->stream>>>(__VA_ARGS__); KERNEL_CHECK(ctx); } \
	} while (0)

// ------------------------------------------------------------------------------------------------
// kernels: grouping
// ------------------------------------------------------------------------------------------------
__global__ void k_group_flags(const int32_t *q, const int64_t *r, int32_t *flag, int64_t n, int max_indel)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	flag[i] = i == 0 || ((r[i] - q[i]) - (r[i - 1] - q[i - 1])) > max_indel; // SeedGrouping, src/GSAlign.cpp:133
}

__global__ void k_group_score(const int32_t *gid1, const int32_t *len, unsigned long long *score, int64_t n)
{ // FindSeedGroupScore, src/GSAlign.cpp:298-303
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	bool in = i < n;
	int g = in ? gid1[i] - 1 : -1, v = in ? len[i] : 0;
	// warp-aggregate when the whole warp sits in one group (the common case: one giant diagonal group)
	int g0 = __shfl_sync(0xffffffffu, g, 0);
	if (__all_sync(0xffffffffu, g == g0)) {
		for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
		if ((threadIdx.x & 31) == 0 && g0 >= 0) atomicAdd(score + g0, (unsigned long long)(long long)v);
	} else if (in) atomicAdd(score + g, (unsigned long long)(long long)v);
}

__global__ void k_group_keep(const int32_t *gid1, const unsigned long long *score, uint8_t *keep, int64_t n, int min_score)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	keep[i] = (long long)score[gid1[i] - 1] >= (long long)min_score; // FindSeedGroupScore < MinAlnBlockScore -> skip, src/GSAlign.cpp:387
}

// sort key of the per-group order CompByQueryPos (src/ProcessCandidateAlignment.cpp:9-13): (group, qPos); ties on
// qPos keep their (PosDiff,qPos) input order under the stable radix sort, which for equal qPos is rPos order
__global__ void k_group_keys(const int32_t *idx, const int32_t *gid1, const int32_t *q, uint64_t *key, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int32_t s = idx[i];
	key[i] = ((uint64_t)(uint32_t)(gid1[s] - 1) << 32) | (uint32_t)q[s];
}

__global__ void k_gather_seeds(const int32_t *idx, const int32_t *q, const int64_t *r, const int32_t *l, const int32_t *g,
                               int32_t *oq, int64_t *orr, int32_t *ol, int32_t *og, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int32_t s = idx[i];
	oq[i] = q[s]; orr[i] = r[s]; ol[i] = l[s]; og[i] = g[s];
}

__global__ void k_seg_flags(const int32_t *g, uint8_t *flag, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	flag[i] = i == 0 || g[i] != g[i - 1];
}

__global__ void k_fill_i32(int32_t *a, int32_t v, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) a[i] = v;
}

// dense segment id per element from the ascending list of segment starts
__global__ void k_seg_ids(const int32_t *starts, const int32_t *d_nseg, int32_t *dg, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int lo = 0, hi = *d_nseg; // last start <= i
	while (lo < hi) { int m = (lo + hi) >> 1; if (starts[m] <= (int32_t)i) lo = m + 1; else hi = m; }
	dg[i] = lo - 1;
}

// ------------------------------------------------------------------------------------------------
// kernels: SeedGroupAnalysis
// ------------------------------------------------------------------------------------------------
struct GroupView {
	const int32_t *q; const int64_t *r; const int32_t *l; const int32_t *dg; const int32_t *gstart; // gstart[ng] = n
	int64_t n;
};

__device__ __forceinline__ int64_t pd_of(const GroupView &v, int64_t i) { return v.r[i] - v.q[i]; }

__global__ void k_uniq(GroupView v, int32_t *uq)
{ // UniqueArr, src/GSAlign.cpp:316-325
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= v.n) return;
	int gs = v.gstart[v.dg[i]], ge = v.gstart[v.dg[i] + 1];
	bool same_prev = i > gs && v.q[i - 1] == v.q[i], same_next = i + 1 < ge && v.q[i + 1] == v.q[i];
	uq[i] = !(same_prev || same_next);
}

__global__ void k_cand(GroupView v, const int32_t *uq, uint8_t *cand)
{ // positions where a window may close: unique, and PosDiff differs from the previous element (src/GSAlign.cpp:328-331)
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= v.n) return;
	int gs = v.gstart[v.dg[i]];
	cand[i] = uq[i] && i > gs && pd_of(v, i) != pd_of(v, i - 1);
}

// next window start after a window that starts at i (src/GSAlign.cpp:326-337):
// the first candidate j > i with (#unique in the window so far) >= 30 and q[j] - q[i] > 3000
__global__ void k_next(GroupView v, const int32_t *uq, const int32_t *U, const uint8_t *cand, const int32_t *C, const int32_t *d_nC, int32_t *nxt)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i > v.n) return;
	if (i == v.n) { nxt[i] = (int32_t)v.n; return; }
	int gs = v.gstart[v.dg[i]], ge = v.gstart[v.dg[i] + 1];
	if (i != gs && !cand[i]) { nxt[i] = (int32_t)v.n; return; }
	int baseU = i == gs ? U[i] - uq[i] : U[i]; // the first window counts its own first element, later ones restart at 0
	int lo = (int)i + 1, hi = ge;
	while (lo < hi) { int m = (lo + hi) >> 1; if (U[m] - baseU >= 30) hi = m; else lo = m + 1; }
	int jA = lo;
	lo = (int)i + 1; hi = ge;
	int qi = v.q[i];
	while (lo < hi) { int m = (lo + hi) >> 1; if (v.q[m] - qi > 3000) hi = m; else lo = m + 1; }
	int j0 = max(jA, lo);
	int res = (int)v.n;
	if (j0 < ge) {
		int nC = *d_nC; lo = 0; hi = nC;
		while (lo < hi) { int m = (lo + hi) >> 1; if (C[m] < j0) lo = m + 1; else hi = m; }
		if (lo < nC && C[lo] < ge) res = C[lo];
	}
	nxt[i] = res;
}

__global__ void k_reach_init(GroupView v, int32_t *reach)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i > v.n) return;
	reach[i] = i < v.n && (int64_t)v.gstart[v.dg[i]] == i;
}

// pointer jumping: reach is monotone and updated in place, the jump table is ping-ponged
__global__ void k_jump(int32_t *reach, const int32_t *nin, int32_t *nout, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i > n) return;
	int32_t j = nin[i];
	if (reach[i] && j < n) reach[j] = 1;
	nout[i] = nin[j];
}

#define HASH_EMPTY 0xFFFFFFFFFFFFFFFFull
__device__ __forceinline__ uint32_t hash64(uint64_t k)
{
	k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
	return (uint32_t)k;
}

// per-window PosDiff>>4 histogram (PDFmap, src/GSAlign.cpp:264-270) as a global (window,bin) hash table
__global__ void k_hist_insert(GroupView v, const int32_t *uq, const int32_t *wid1, unsigned long long *keys, int32_t *cnt, int32_t *slot_of, uint32_t hmask)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const bool act = i < v.n && uq[i];
	// inactive lanes get distinct keys no (window,bin) can take (window ids are < 2^31)
	unsigned long long key = 0xFFFFFFFF00000000ull | (threadIdx.x & 31);
	if (act) key = ((unsigned long long)(uint32_t)(wid1[i] - 1) << 32) | (uint32_t)(int32_t)(pd_of(v, i) >> 4);
	// neighbouring seeds mostly share the bin: one probe + one add per distinct key of the warp
	unsigned peers = __match_any_sync(0xffffffffu, key);
	bool leader;
	int total = gsa_peer_sum(peers, 1, leader);
	uint32_t s = 0;
	if (leader && act) {
		s = hash64(key) & hmask;
		for (;;) {
			unsigned long long old = atomicCAS(keys + s, HASH_EMPTY, key);
			if (old == HASH_EMPTY || old == key) break;
			s = (s + 1) & hmask;
		}
		atomicAdd(cnt + s, total);
	}
	s = __shfl_sync(0xffffffffu, s, __ffs(peers) - 1);
	if (act) slot_of[i] = (int32_t)s;
}

// mode per window = the smallest bin with the maximal count (RefinePDFmap, src/GSAlign.cpp:250-251)
__global__ void k_win_best(const unsigned long long *keys, const int32_t *cnt, unsigned long long *best, int64_t hsize)
{
	int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= hsize) return;
	unsigned long long key = keys[s];
	if (key == HASH_EMPTY) return;
	int32_t bin = (int32_t)(uint32_t)key;
	unsigned long long packed = ((unsigned long long)(uint32_t)cnt[s] << 32) | (uint32_t)(0x7FFFFFFFll - (long long)bin);
	atomicMax(best + (key >> 32), packed);
}

__device__ __forceinline__ int32_t mode_of(unsigned long long packed) { return (int32_t)(0x7FFFFFFFll - (long long)(uint32_t)packed); }

// sum / count of PosDiff over unique seeds whose bin survives |bin - mode| < 3 (src/GSAlign.cpp:254-282)
__global__ void k_win_sum(GroupView v, const int32_t *uq, const int32_t *wid1, const unsigned long long *best, unsigned long long *sum, int32_t *cntk)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	int w = -1 - (int)(threadIdx.x & 31);
	int64_t pd = 0;
	if (i < v.n && uq[i]) {
		int wi = wid1[i] - 1;
		pd = pd_of(v, i);
		int32_t bin = (int32_t)(pd >> 4), mode = mode_of(best[wi]);
		long long d = (long long)bin - mode;
		if (d < 0) d = -d;
		if (d < 3) w = wi;
	}
	unsigned peers = __match_any_sync(0xffffffffu, w);
	bool leader;
	unsigned long long tot = gsa_peer_sum(peers, (unsigned long long)(w >= 0 ? pd : 0), leader);
	int c = gsa_peer_sum(peers, w >= 0 ? 1 : 0, leader);
	if (leader && w >= 0) { atomicAdd(sum + w, tot); atomicAdd(cntk + w, c); }
}

__global__ void k_outlier_kill(GroupView v, const int32_t *uq, const int32_t *wid1, const unsigned long long *best, const unsigned long long *sum,
                               const int32_t *cntk, const int32_t *cnt, const int32_t *slot_of, uint8_t *alive, int64_t genome, int max_indel)
{ // src/GSAlign.cpp:282-294 with Check_PD_Frequency (:145-153), Min_PD_Freq = 3
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= v.n) return;
	uint8_t a = 1;
	if (uq[i]) {
		int w = wid1[i] - 1;
		int64_t pd = pd_of(v, i);
		int32_t bin = (int32_t)(pd >> 4), mode = mode_of(best[w]);
		long long d = (long long)bin - mode;
		if (d < 0) d = -d;
		int own = d < 3 ? cnt[slot_of[i]] : 0;
		int64_t avg = cntk[w] > 0 ? (int64_t)sum[w] / cntk[w] : genome;
		int64_t diff = avg - pd;
		if (diff < 0) diff = -diff;
		if (diff > max_indel && own < 3) a = 0;
	}
	alive[i] = a;
}

__global__ void k_live_unique(const int32_t *uq, const uint8_t *alive, uint8_t *lu, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) lu[i] = uq[i] && alive[i];
}

// multi-hit runs (same qPos): keep the hit nearest to the mean PosDiff of <= 5 + 5 neighbouring live unique seeds
// (src/GSAlign.cpp:341-350 with FindNeighboringPosDiffAvg :178-206 and RemoveRedundantSeeds :208-225)
__global__ void k_runs(GroupView v, const int32_t *LU, const int32_t *d_nLU, uint8_t *alive, int64_t genome, int max_indel)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= v.n) return;
	int gs = v.gstart[v.dg[i]], ge = v.gstart[v.dg[i] + 1];
	int qi = v.q[i];
	if (i > gs && v.q[i - 1] == qi) return;         // not a run start
	if (!(i + 1 < ge && v.q[i + 1] == qi)) return;  // unique
	int j = (int)i + 1;
	while (j < ge && v.q[j] == qi) j++;
	int nL = *d_nLU, lo = 0, hi = nL;
	while (lo < hi) { int m = (lo + hi) >> 1; if (LU[m] < (int)i) lo = m + 1; else hi = m; }
	int64_t sum = 0; int cnt = 0;
	for (int p = lo - 1, k = 0; p >= 0 && k < 5 && LU[p] >= gs; p--, k++) { sum += pd_of(v, LU[p]); cnt++; }
	lo = 0; hi = nL;
	while (lo < hi) { int m = (lo + hi) >> 1; if (LU[m] < j) lo = m + 1; else hi = m; }
	for (int p = lo, k = 0; p < nL && k < 5 && LU[p] < ge; p++, k++) { sum += pd_of(v, LU[p]); cnt++; }
	int64_t avg = cnt > 0 ? sum / cnt : pd_of(v, i);
	int keep = -1; int64_t min_diff = genome;
	for (int k = (int)i; k < j; k++) {
		int64_t d = pd_of(v, k) - avg;
		if (d < 0) d = -d;
		if (d < max_indel && d < min_diff) { min_diff = d; keep = k; }
	}
	for (int k = (int)i; k < j; k++) if (k != keep) alive[k] = 0;
}

// noise: interior seed whose PosDiff is > 5 away from both live neighbours of its group (src/GSAlign.cpp:355-362)
__global__ void k_noise(const int32_t *q, const int64_t *r, const int32_t *g, uint8_t *alive, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint8_t a = 1;
	if (i > 0 && i + 1 < n && g[i - 1] == g[i] && g[i + 1] == g[i]) {
		int64_t p = r[i] - q[i], a1 = p - (r[i - 1] - q[i - 1]), a2 = p - (r[i + 1] - q[i + 1]);
		if (a1 < 0) a1 = -a1;
		if (a2 < 0) a2 = -a2;
		if (a1 > 5 && a2 > 5) a = 0;
	}
	alive[i] = a;
}

// block cuts inside a group (src/GSAlign.cpp:364-374)
__global__ void k_cut(const int32_t *q, const int64_t *r, const int32_t *l, const int32_t *g, uint8_t *flag, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	bool cut = i == 0 || g[i] != g[i - 1];
	if (!cut) {
		int64_t d = (r[i - 1] - q[i - 1]) - (r[i] - q[i]);
		if (d < 0) d = -d;
		cut = q[i] - q[i - 1] - l[i - 1] > GSA_MAX_SEED_GAP || d > 100;
	}
	flag[i] = cut;
}

// out[i] = l[i] for i < n, out[n] = 0 (launch n + 1 threads) so that an exclusive scan over n + 1 yields S[n] = total
__global__ void k_len64(const int32_t *l, int64_t *out, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = l[i];
	else if (i == n) out[i] = 0;
}

__global__ void k_widen(const uint8_t *in, int32_t *out, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = in[i];
}

// AddAlnBlock acceptance (src/GSAlign.cpp:29-49); S = exclusive prefix sum of len, S[n] = total
__global__ void k_block_eval(const int32_t *bstart, int64_t nb, int64_t n, const int32_t *q, const int32_t *l, const int64_t *S,
                             int32_t *score, uint8_t *accept, int min_score, int min_len)
{
	int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nb) return;
	int64_t beg = bstart[b], end = b + 1 < nb ? bstart[b + 1] : n;
	int32_t sc = (int32_t)(S[end] - S[beg]);
	int32_t region = (q[end - 1] + l[end - 1]) - q[beg];
	score[b] = sc;
	accept[b] = !(sc < min_score || region < min_len || (sc < 1000 && sc < region * 0.05));
}

__global__ void k_seed_accept(const int32_t *bid, const uint8_t *accept, uint8_t *keep, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) keep[i] = accept[bid[i]];
}

__global__ void k_gather_block_seeds(const int32_t *idx, const int32_t *q, const int64_t *r, const int32_t *l, const int32_t *bid,
                                     const int32_t *newid1, int32_t *oq, int64_t *orr, int32_t *ol, int32_t *ob, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int32_t s = idx[i];
	oq[i] = q[s]; orr[i] = r[s]; ol[i] = l[s]; ob[i] = newid1[bid[s]] - 1;
}

// ------------------------------------------------------------------------------------------------
// kernels: RemoveOverlaps, gap / span break points, pieces
// ------------------------------------------------------------------------------------------------
__global__ void k_overlap_pass(const int32_t *q, const int64_t *r, int32_t *l, const int32_t *b, uint8_t *alive, int32_t *kills, int64_t n)
{ // one pass of RemoveOverlaps (src/ProcessCandidateAlignment.cpp:197-227): element i only reads q/r of i+1
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint8_t a = 1;
	if (i + 1 < n && b[i + 1] == b[i]) {
		if (r[i + 1] <= r[i]) a = 0;
		else {
			int32_t len = l[i], ov = (int32_t)(r[i] + len - r[i + 1]);
			if (ov > 0) { len -= ov; if (len <= 0) a = 0; }
			if (a && (ov = q[i] + len - q[i + 1]) > 0) { len -= ov; if (len <= 0) a = 0; }
			l[i] = len;
		}
	}
	alive[i] = a;
	if (!a) atomicAdd(kills, 1); // rare
}

__device__ __forceinline__ int64_t contig_end_of(const ContigEnd *ce, int nce, int64_t rpos)
{
	int lo = 0, hi = nce;
	while (lo < hi) { int m = (lo + hi) >> 1; if (ce[m].end < rpos) lo = m + 1; else hi = m; }
	return ce[lo < nce ? lo : nce - 1].end;
}

// bit0: certain gap break, bit1: contig-span break, bit2: gap needs the similarity test
__global__ void k_gap_flags(const int32_t *q, const int64_t *r, const int32_t *l, const int32_t *b, const ContigEnd *ce, int nce, uint8_t *flag, uint8_t *need, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint8_t f = 0;
	if (i > 0 && b[i] == b[i - 1]) {
		int32_t qg = q[i] - q[i - 1] - l[i - 1], rg = (int32_t)(r[i] - r[i - 1] - l[i - 1]);
		if (qg > 300 || rg > 300) f |= (qg > GSA_MAX_SEED_GAP || rg > GSA_MAX_SEED_GAP) ? 1 : 4;
		if (contig_end_of(ce, nce, r[i]) != contig_end_of(ce, nce, r[i - 1])) f |= 2;
	}
	flag[i] = f; need[i] = (f >> 2) & 1;
}

// CalGapSimilarity (src/KmerAnalysis.cpp:78-121) for the gap in front of seed cand[blockIdx.x]; one warp per gap.
// hist: ids of CreateKmerVecFromReadSeq are < 2048 (rolling ((id & 0xFF) << 2) + nt with nt in 0..4)
__global__ void __launch_bounds__(32) k_gap_similarity(const int32_t *cand, const int32_t *q, const int64_t *r, const int32_t *l,
                                                       const unsigned char *seq, DevIndex ix, uint8_t *flag)
{
	__shared__ unsigned short h1[2048], h2[2048];
	int i = cand[blockIdx.x], lane = threadIdx.x;
	int q1 = q[i - 1] + l[i - 1], q2 = q[i];
	int64_t r1 = r[i - 1] + l[i - 1], r2 = r[i];
	int q_len = q2 - q1, r_len = (int)(r2 - r1);
	bool similar = false;
	if (r1 - q1 == r2 - q2) { // same diagonal: linear identity (ref text never holds N)
		int idy = 0;
		for (int k = lane; k < q_len; k += 32) {
			int a = gsa_pk_base(ix.txt, (uint64_t)(r1 + k)), c = gsa_nt4(seq[q1 + k]);
			idy += (a == c || c == 4);
		}
		for (int o = 16; o > 0; o >>= 1) idy += __shfl_xor_sync(0xffffffffu, idy, o);
		similar = idy >= q_len * 0.5;
	}
	if (!similar && q_len <= GSA_MAX_SEED_GAP && r_len <= GSA_MAX_SEED_GAP) {
		for (int k = lane; k < 2048; k += 32) { h1[k] = 0; h2[k] = 0; }
		__syncwarp();
		if (lane == 0) { // query k-mers, quirks kept: only the byte 'N' restarts, stale head after a restart
			const unsigned char *s = seq + q1;
			uint32_t wid = 0, count = 0, head = 0, tail = 0, len = (uint32_t)q_len;
			while (count < 5 && tail < len) { if (s[tail++] != 'N') count++; else count = 0; }
			if (count == 5) {
				for (uint32_t k = head; k < head + 5; k++) wid = (wid << 2) + (uint32_t)gsa_nt4(s[k]);
				h1[wid]++;
				for (head += 1; tail < len; head++, tail++) {
					if (s[tail] != 'N') { wid = ((wid & 0xFF) << 2) + (uint32_t)gsa_nt4(s[tail]); h1[wid]++; }
					else {
						count = 0; tail++;
						while (count < 5 && tail < len) { if (s[tail++] != 'N') count++; else count = 0; }
						if (count != 5) break;
						wid = 0;
						for (uint32_t k = head; k < head + 5; k++) wid = (wid << 2) + (uint32_t)gsa_nt4(s[k]);
						h1[wid]++;
					}
				}
			}
		} else if (lane == 1 && r_len >= 5) { // reference k-mers (no N in the text)
			uint32_t wid = 0;
			for (int k = 0; k < 5; k++) wid = (wid << 2) + (uint32_t)gsa_pk_base(ix.txt, (uint64_t)(r1 + k));
			h2[wid]++;
			for (int k = 5; k < r_len; k++) { wid = ((wid & 0xFF) << 2) + (uint32_t)gsa_pk_base(ix.txt, (uint64_t)(r1 + k)); h2[wid]++; }
		}
		__syncwarp();
		int common = 0;
		for (int k = lane; k < 2048; k += 32) common += min((int)h1[k], (int)h2[k]);
		for (int o = 16; o > 0; o >>= 1) common += __shfl_xor_sync(0xffffffffu, common, o);
		similar = common > (q_len + r_len) * 0.1;
	}
	if (lane == 0 && !similar) flag[i] |= 1;
}

__global__ void k_piece_flags(const int32_t *b, const uint8_t *flag, uint8_t mask, uint8_t *out, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = i == 0 || b[i] != b[i - 1] || (flag[i] & mask);
}

__global__ void k_piece_table(const int32_t *starts, int64_t np, int64_t n, const int32_t *q, const int64_t *r, const int32_t *l, const int64_t *S, Piece *out)
{
	int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= np) return;
	int64_t beg = starts[p], end = p + 1 < np ? starts[p + 1] : n;
	Piece x; x.beg = beg; x.end = end; x.sumlen = S[end] - S[beg]; x.rf = r[beg]; x.rl = r[end - 1];
	x.qf = q[beg]; x.ql = q[end - 1]; x.lenl = l[end - 1]; x.pad = 0;
	out[p] = x;
}

// ------------------------------------------------------------------------------------------------
// kernels: IdentifyNormalPairs -> fragment list
// ------------------------------------------------------------------------------------------------
struct NpBlock { int64_t src_beg, dst_beg; int32_t n, pad; };

// element t of the concatenated kept blocks -> (block k, source seed s)
__device__ __forceinline__ void np_locate(const NpBlock *nb, int nblk, int64_t t, int &k, int64_t &s)
{
	int lo = 0, hi = nblk;
	while (lo < hi) { int m = (lo + hi) >> 1; if (nb[m].dst_beg <= t) lo = m + 1; else hi = m; }
	k = lo - 1; s = nb[k].src_beg + (t - nb[k].dst_beg);
}

__global__ void k_np_count(const NpBlock *nb, int nblk, int64_t total, const int32_t *q, const int64_t *r, const int32_t *l, int32_t *cnt)
{
	int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= total) return;
	int k; int64_t s; np_locate(nb, nblk, t, k, s);
	int c = 1;
	if (t + 1 < nb[k].dst_beg + nb[k].n) { // not the last seed of its block
		int32_t qg = q[s + 1] - (q[s] + l[s]); int64_t rg = r[s + 1] - (r[s] + l[s]);
		if (qg > 0 || (int32_t)rg > 0) c = 2;
	}
	cnt[t] = c;
}

__global__ void k_np_write(const NpBlock *nb, int nblk, int64_t total, const int32_t *q, const int64_t *r, const int32_t *l, const int32_t *off,
                           gsa_frag *frag, int32_t *fblk, int64_t *blk_frag_beg)
{
	int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= total) return;
	int k; int64_t s; np_locate(nb, nblk, t, k, s);
	int64_t o = off[t];
	if (t == nb[k].dst_beg) blk_frag_beg[k] = o;
	gsa_frag f; f.rPos = r[s]; f.qPos = q[s]; f.qLen = l[s]; f.rLen = l[s]; f.bSeed = 1; f.aln_off = 0; f.aln_len = l[s]; f.reserved = 0;
	frag[o] = f; fblk[o] = k;
	if (t + 1 < nb[k].dst_beg + nb[k].n) {
		int32_t qg = q[s + 1] - (q[s] + l[s]); int32_t rg = (int32_t)(r[s + 1] - (r[s] + l[s]));
		if (qg < 0) qg = 0;
		if (rg < 0) rg = 0;
		if (qg > 0 || rg > 0) {
			gsa_frag g; g.rPos = r[s] + l[s]; g.qPos = q[s] + l[s]; g.qLen = qg; g.rLen = rg; g.bSeed = 0; g.aln_off = 0; g.aln_len = 0; g.reserved = 0;
			frag[o + 1] = g; fblk[o + 1] = k;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
static BlockHdr hdr_from_piece(const Piece &p, int32_t score)
{
	BlockHdr b; b.score = score; b.bDup = 0; b.beg = p.beg; b.end = p.end; b.qf = p.qf; b.ql = p.ql; b.lenl = p.lenl; b.rf = p.rf; b.rl = p.rl;
	b.frag_beg = 0; b.n_frags = 0; b.aln_len = 0;
	return b;
}

static int fetch_pieces(gsa_ctx *ctx, Ws &ws, const int32_t *cb, const uint8_t *gflag, uint8_t mask, int64_t n, const int32_t *q, const int64_t *r, const int32_t *l,
                        const int64_t *S, uint8_t *pflag, int32_t *pstart, int32_t *d_cnt, std::vector<Piece> &out)
{
	LAUNCH(k_piece_flags, n, cb, gflag, mask, pflag, n);
	GSA_TRY(select_indices(ctx, pflag, pstart, d_cnt, n));
	int64_t np = 0;
	GSA_TRY(read_count(ctx, d_cnt, &np));
	out.resize((size_t)np);
	if (np == 0) return GSA_OK;
	Piece *d_p = ws.get<Piece>(np);
	if (!d_p) return ws.rc;
	LAUNCH(k_piece_table, np, pstart, np, n, q, r, l, S, d_p);
	GSA_TRY(gsa_ensure_host(ctx, ctx->h_stage, (size_t)np * sizeof(Piece)));
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_stage.p, d_p, (size_t)np * sizeof(Piece), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	memcpy(out.data(), ctx->h_stage.p, (size_t)np * sizeof(Piece));
	return GSA_OK;
}

int gsa_impl_cluster(gsa_ctx *ctx)
{
	Ws ws(ctx);
	const gsa_params &P = ctx->prm;
	for (auto &v : ctx->blocks_stage) v.clear();
	ctx->final_blocks.clear(); ctx->n_cseeds = 0; ctx->n_frags = 0; ctx->n_s0 = 0;
	int64_t n = ctx->n_seeds;
	if (n >= 0x7FFFFFF0ll) return gsa_fail(ctx, GSA_ERR_LIMIT, "gsa_cluster: more than 2^31 seeds in one contig");
	if (n == 0) return GSA_OK;
	const int32_t *sq = (const int32_t *)ctx->d_sq.p; const int64_t *sr = (const int64_t *)ctx->d_sr.p; const int32_t *sl = (const int32_t *)ctx->d_sl.p;
	int32_t *d_cnt = (int32_t *)ctx->d_counter.p + 8; // a few device counters

	// ---- 1. diagonal groups over the (PosDiff,qPos)-sorted seeds; drop groups below MinAlnBlockScore -------------------
	int32_t *gflag = ws.get<int32_t>(n), *gid1 = ws.get<int32_t>(n);
	uint8_t *keep = ws.get<uint8_t>(n);
	int32_t *kidx = ws.get<int32_t>(n);
	if (ws.rc) return ws.rc;
	LAUNCH(k_group_flags, n, sq, sr, gflag, n, P.max_indel);
	GSA_TRY(scan_inclusive(ctx, gflag, gid1, n));
	int64_t ngroups = 0;
	GSA_TRY(read_count(ctx, gid1 + (n - 1), &ngroups));
	unsigned long long *gscore = ws.get<unsigned long long>(ngroups);
	if (ws.rc) return ws.rc;
	CUDA_TRY(ctx, cudaMemsetAsync(gscore, 0, (size_t)ngroups * 8, ctx->stream));
	LAUNCH(k_group_score, n, gid1, sl, gscore, n);
	LAUNCH(k_group_keep, n, gid1, gscore, keep, n, P.min_block_score);
	GSA_TRY(select_indices(ctx, keep, kidx, d_cnt, n));
	int64_t n2 = 0;
	GSA_TRY(read_count(ctx, d_cnt, &n2));
	if (n2 == 0) return GSA_OK;

	// ---- 2. per-group order (qPos, rPos): stable radix sort on (group, qPos) ---------------------------------------------
	uint64_t *key_in = ws.get<uint64_t>(n2), *key_out = ws.get<uint64_t>(n2);
	int32_t *val_out = ws.get<int32_t>(n2);
	int32_t *q2 = ws.get<int32_t>(n2 + 1), *l2 = ws.get<int32_t>(n2 + 1), *g2 = ws.get<int32_t>(n2 + 1);
	int64_t *r2 = ws.get<int64_t>(n2 + 1);
	if (ws.rc) return ws.rc;
	LAUNCH(k_group_keys, n2, kidx, gid1, sq, key_in, n2);
	{
		int gbits = 1; while ((1ll << gbits) < ngroups + 1) gbits++;
		size_t bytes = 0;
		cub::DeviceRadixSort::SortPairs(nullptr, bytes, key_in, key_out, kidx, val_out, (int)n2, 0, 32 + gbits, ctx->stream);
		GSA_TRY(gsa_ensure(ctx, ctx->d_cub, bytes));
		CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->d_cub.p, bytes, key_in, key_out, kidx, val_out, (int)n2, 0, 32 + gbits, ctx->stream));
		ctx->tm.launches += 4;
	}
	LAUNCH(k_gather_seeds, n2, val_out, sq, sr, sl, gid1, q2, r2, l2, g2, n2);

	// dense group ids + group start table
	uint8_t *f8 = ws.get<uint8_t>(n2 + 1);
	int32_t *gstart = ws.get<int32_t>(n2 + 2), *dg = ws.get<int32_t>(n2 + 1);
	if (ws.rc) return ws.rc;
	LAUNCH(k_seg_flags, n2, g2, f8, n2);
	GSA_TRY(select_indices(ctx, f8, gstart, d_cnt + 1, n2));
	int64_t ng2 = 0;
	GSA_TRY(read_count(ctx, d_cnt + 1, &ng2));
	LAUNCH(k_fill_i32, 1, gstart + ng2, (int32_t)n2, 1);
	LAUNCH(k_seg_ids, n2, gstart, d_cnt + 1, dg, n2);
	GroupView gv; gv.q = q2; gv.r = r2; gv.l = l2; gv.dg = dg; gv.gstart = gstart; gv.n = n2;

	// ---- 3. outlier windows ----------------------------------------------------------------------------------------------
	int32_t *uq = ws.get<int32_t>(n2 + 1), *U = ws.get<int32_t>(n2 + 1);
	uint8_t *cand = ws.get<uint8_t>(n2 + 1);
	int32_t *C = ws.get<int32_t>(n2 + 1);
	int32_t *nxtA = ws.get<int32_t>(n2 + 1), *nxtB = ws.get<int32_t>(n2 + 1), *reach = ws.get<int32_t>(n2 + 1), *wid1 = ws.get<int32_t>(n2 + 1);
	if (ws.rc) return ws.rc;
	LAUNCH(k_uniq, n2, gv, uq);
	GSA_TRY(scan_inclusive(ctx, uq, U, n2));
	LAUNCH(k_cand, n2, gv, uq, cand);
	GSA_TRY(select_indices(ctx, cand, C, d_cnt + 2, n2));
	LAUNCH(k_next, n2 + 1, gv, uq, U, cand, C, d_cnt + 2, nxtA);
	LAUNCH(k_reach_init, n2 + 1, gv, reach);
	{
		int rounds = 1; while ((1ll << rounds) < n2 / 30 + 2) rounds++;
		for (int k = 0; k <= rounds; k++) { LAUNCH(k_jump, n2 + 1, reach, nxtA, nxtB, n2); std::swap(nxtA, nxtB); }
	}
	GSA_TRY(scan_inclusive(ctx, reach, wid1, n2));

	// per-window histogram of PosDiff>>4 over unique seeds -> mode, average, outlier kill
	int64_t hsize = 1024; while (hsize < 2 * n2) hsize <<= 1;
	unsigned long long *hkeys = ws.get<unsigned long long>(hsize);
	int32_t *hcnt = ws.get<int32_t>(hsize), *slot_of = ws.get<int32_t>(n2);
	unsigned long long *wbest = ws.get<unsigned long long>(n2), *wsum = ws.get<unsigned long long>(n2);
	int32_t *wcnt = ws.get<int32_t>(n2);
	uint8_t *alive = ws.get<uint8_t>(n2 + 1);
	if (ws.rc) return ws.rc;
	CUDA_TRY(ctx, cudaMemsetAsync(hkeys, 0xFF, (size_t)hsize * 8, ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(hcnt, 0, (size_t)hsize * 4, ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(wbest, 0, (size_t)n2 * 8, ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(wsum, 0, (size_t)n2 * 8, ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(wcnt, 0, (size_t)n2 * 4, ctx->stream));
	LAUNCH(k_hist_insert, n2, gv, uq, wid1, hkeys, hcnt, slot_of, (uint32_t)(hsize - 1));
	LAUNCH(k_win_best, hsize, hkeys, hcnt, wbest, hsize);
	LAUNCH(k_win_sum, n2, gv, uq, wid1, wbest, wsum, wcnt);
	LAUNCH(k_outlier_kill, n2, gv, uq, wid1, wbest, wsum, wcnt, hcnt, slot_of, alive, ctx->N, P.max_indel);

	// ---- 4. multi-hit runs ------------------------------------------------------------------------------------------------
	int32_t *LU = ws.get<int32_t>(n2 + 1);
	if (ws.rc) return ws.rc;
	LAUNCH(k_live_unique, n2, uq, alive, f8, n2);
	GSA_TRY(select_indices(ctx, f8, LU, d_cnt + 3, n2));
	LAUNCH(k_runs, n2, gv, LU, d_cnt + 3, alive, ctx->N, P.max_indel);

	// ---- 5. compact, noise filter, compact ------------------------------------------------------------------------------------
	int32_t *idx3 = ws.get<int32_t>(n2 + 1);
	if (ws.rc) return ws.rc;
	GSA_TRY(select_indices(ctx, alive, idx3, d_cnt + 4, n2));
	int64_t n3 = 0;
	GSA_TRY(read_count(ctx, d_cnt + 4, &n3));
	if (n3 == 0) return GSA_OK;
	int32_t *q3 = ws.get<int32_t>(n3 + 1), *l3 = ws.get<int32_t>(n3 + 1), *g3 = ws.get<int32_t>(n3 + 1);
	int64_t *r3 = ws.get<int64_t>(n3 + 1);
	uint8_t *alive3 = ws.get<uint8_t>(n3 + 1);
	int32_t *idx4 = ws.get<int32_t>(n3 + 1);
	if (ws.rc) return ws.rc;
	LAUNCH(k_gather_seeds, n3, idx3, q2, r2, l2, dg, q3, r3, l3, g3, n3);
	LAUNCH(k_noise, n3, q3, r3, g3, alive3, n3);
	GSA_TRY(select_indices(ctx, alive3, idx4, d_cnt + 5, n3));
	int64_t n4 = 0;
	GSA_TRY(read_count(ctx, d_cnt + 5, &n4));
	if (n4 == 0) return GSA_OK;
	// reuse the n2-sized arrays for the n4 generation (n4 <= n3 <= n2)
	int32_t *q4 = q2, *l4 = l2, *g4 = g2; int64_t *r4 = r2;
	LAUNCH(k_gather_seeds, n4, idx4, q3, r3, l3, g3, q4, r4, l4, g4, n4);

	// ---- 6. cut groups into blocks, AddAlnBlock acceptance ------------------------------------------------------------------------
	int32_t *bid = uq, *bstart = C; // reuse n2-sized scratch
	int64_t *len64 = ws.get<int64_t>(n4 + 2), *S = ws.get<int64_t>(n4 + 2);
	if (ws.rc) return ws.rc;
	LAUNCH(k_cut, n4, q4, r4, l4, g4, f8, n4);
	GSA_TRY(select_indices(ctx, f8, bstart, d_cnt + 6, n4));
	int64_t nb0 = 0;
	GSA_TRY(read_count(ctx, d_cnt + 6, &nb0));
	LAUNCH(k_seg_ids, n4, bstart, d_cnt + 6, bid, n4);
	LAUNCH(k_len64, n4 + 1, l4, len64, n4);
	GSA_TRY(scan_exclusive(ctx, len64, S, n4 + 1));
	int32_t *bscore = ws.get<int32_t>(nb0 + 1), *acc32 = ws.get<int32_t>(nb0 + 1), *newid1 = ws.get<int32_t>(nb0 + 1);
	uint8_t *accept = ws.get<uint8_t>(nb0 + 1);
	if (ws.rc) return ws.rc;
	LAUNCH(k_block_eval, nb0, bstart, nb0, n4, q4, l4, S, bscore, accept, P.min_block_score, P.min_aln_len);
	LAUNCH(k_widen, nb0, accept, acc32, nb0);
	GSA_TRY(scan_inclusive(ctx, acc32, newid1, nb0));
	LAUNCH(k_seed_accept, n4, bid, accept, f8, n4);
	int32_t *idx5 = idx3;
	GSA_TRY(select_indices(ctx, f8, idx5, d_cnt + 7, n4));
	int64_t n5 = 0;
	GSA_TRY(read_count(ctx, d_cnt + 7, &n5));
	if (n5 == 0) return GSA_OK;
	// the working set of the remaining phases lives in the context (K3 and the dump hooks read it)
	GSA_TRY(gsa_ensure(ctx, ctx->d_cq, (size_t)(n5 + 1) * 4)); GSA_TRY(gsa_ensure(ctx, ctx->d_cr, (size_t)(n5 + 1) * 8));
	GSA_TRY(gsa_ensure(ctx, ctx->d_cl, (size_t)(n5 + 1) * 4)); GSA_TRY(gsa_ensure(ctx, ctx->d_cb, (size_t)(n5 + 1) * 4));
	int32_t *cq = (int32_t *)ctx->d_cq.p, *cl = (int32_t *)ctx->d_cl.p, *cb = (int32_t *)ctx->d_cb.p; int64_t *cr = (int64_t *)ctx->d_cr.p;
	LAUNCH(k_gather_block_seeds, n5, idx5, q4, r4, l4, bid, newid1, cq, cr, cl, cb, n5);

	// level-0 piece table = the candidate blocks in the reference's -t 1 push order (group order, then qPos order)
	uint8_t *gapf = ws.get<uint8_t>(n5 + 1), *need = ws.get<uint8_t>(n5 + 1), *pflag = ws.get<uint8_t>(n5 + 1);
	int32_t *pstart = ws.get<int32_t>(n5 + 2);
	int64_t *S5 = ws.get<int64_t>(n5 + 2);
	if (ws.rc) return ws.rc;
	LAUNCH(k_len64, n5 + 1, cl, len64, n5);
	GSA_TRY(scan_exclusive(ctx, len64, S5, n5 + 1));
	CUDA_TRY(ctx, cudaMemsetAsync(gapf, 0, (size_t)n5, ctx->stream));
	std::vector<Piece> pc0, pc1, pc2;
	GSA_TRY(fetch_pieces(ctx, ws, cb, gapf, 0, n5, cq, cr, cl, S5, pflag, pstart, d_cnt + 8, pc0));
	std::vector<BlockHdr> vec;
	vec.reserve(pc0.size());
	for (const Piece &p : pc0) vec.push_back(hdr_from_piece(p, (int32_t)p.sumlen)); // score = sum of seed lengths (AddAlnBlock :36)
	if (ctx->keep_dumps) {
		ctx->blocks_stage[0] = vec; ctx->n_s0 = n5;
		GSA_TRY(gsa_ensure(ctx, ctx->d_s0q, (size_t)n5 * 4)); GSA_TRY(gsa_ensure(ctx, ctx->d_s0r, (size_t)n5 * 8)); GSA_TRY(gsa_ensure(ctx, ctx->d_s0l, (size_t)n5 * 4));
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_s0q.p, cq, (size_t)n5 * 4, cudaMemcpyDeviceToDevice, ctx->stream));
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_s0r.p, cr, (size_t)n5 * 8, cudaMemcpyDeviceToDevice, ctx->stream));
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_s0l.p, cl, (size_t)n5 * 4, cudaMemcpyDeviceToDevice, ctx->stream));
	}

	// ---- 7. RemoveOverlaps: elementwise passes + compaction until nothing dies ------------------------------------------------------
	int64_t n6 = n5;
	{
		int32_t *tq = q3, *tl = l3, *tb = g3; int64_t *tr = r3; // n3-sized scratch (n5 <= n3)
		uint8_t *al = alive;
		for (int pass = 0; pass < 1000; pass++) {
			CUDA_TRY(ctx, cudaMemsetAsync(d_cnt + 9, 0, 4, ctx->stream));
			LAUNCH(k_overlap_pass, n6, cq, cr, cl, cb, al, d_cnt + 9, n6);
			int64_t kills = 0;
			GSA_TRY(read_count(ctx, d_cnt + 9, &kills));
			if (kills == 0) break;
			GSA_TRY(select_indices(ctx, al, idx5, d_cnt + 10, n6));
			int64_t m = n6 - kills;
			LAUNCH(k_gather_seeds, m, idx5, cq, cr, cl, cb, tq, tr, tl, tb, m);
			CUDA_TRY(ctx, cudaMemcpyAsync(cq, tq, (size_t)m * 4, cudaMemcpyDeviceToDevice, ctx->stream));
			CUDA_TRY(ctx, cudaMemcpyAsync(cr, tr, (size_t)m * 8, cudaMemcpyDeviceToDevice, ctx->stream));
			CUDA_TRY(ctx, cudaMemcpyAsync(cl, tl, (size_t)m * 4, cudaMemcpyDeviceToDevice, ctx->stream));
			CUDA_TRY(ctx, cudaMemcpyAsync(cb, tb, (size_t)m * 4, cudaMemcpyDeviceToDevice, ctx->stream));
			n6 = m;
		}
	}
	ctx->n_cseeds = n6;
	LAUNCH(k_len64, n6 + 1, cl, len64, n6);
	GSA_TRY(scan_exclusive(ctx, len64, S5, n6 + 1));
	CUDA_TRY(ctx, cudaMemsetAsync(gapf, 0, (size_t)n6, ctx->stream));
	// blocks keep their push order and their pre-overlap score; only the ranges move
	GSA_TRY(fetch_pieces(ctx, ws, cb, gapf, 0, n6, cq, cr, cl, S5, pflag, pstart, d_cnt + 8, pc0));
	if (pc0.size() != vec.size()) return gsa_fail(ctx, GSA_ERR_CUDA, "gsa_cluster: block count changed in RemoveOverlaps (%zu -> %zu)", vec.size(), pc0.size());
	for (size_t i = 0; i < vec.size(); i++) { int32_t sc = vec[i].score; vec[i] = hdr_from_piece(pc0[i], sc); }
	if (ctx->keep_dumps) ctx->blocks_stage[1] = vec;

	// ---- 8. gap and contig-span break points -> piece tables -> host split logic ---------------------------------------------------------
	LAUNCH(k_gap_flags, n6, cq, cr, cl, cb, (const ContigEnd *)ctx->d_cend.p, (int)ctx->cend.size(), gapf, need, n6);
	GSA_TRY(select_indices(ctx, need, idx5, d_cnt + 11, n6));
	int64_t ncand = 0;
	GSA_TRY(read_count(ctx, d_cnt + 11, &ncand));
	if (ncand > 0) {
		k_gap_similarity<<<(unsigned)ncand, 32, 0, ctx->stream>>>(idx5, cq, cr, cl, (const unsigned char *)ctx->d_seq.p, ctx->ix, gapf);
		KERNEL_CHECK(ctx);
	}
	GSA_TRY(fetch_pieces(ctx, ws, cb, gapf, 1, n6, cq, cr, cl, S5, pflag, pstart, d_cnt + 8, pc1));
	GSA_TRY(fetch_pieces(ctx, ws, cb, gapf, 3, n6, cq, cr, cl, S5, pflag, pstart, d_cnt + 8, pc2));
	gsa_host_split(ctx, vec, pc1, pc2);
	if (ctx->keep_dumps) ctx->blocks_stage[2] = vec;

	// ---- 9. block-level dedup on the host (float ratios + std::sort ties, O(#blocks)) ------------------------------------------------------
	gsa_host_dedup(ctx, vec);

	// ---- 10. IdentifyNormalPairs for the surviving blocks -> fragment list --------------------------------------------------------------------
	int nblk = (int)vec.size();
	ctx->final_blocks = vec;
	if (nblk == 0) return GSA_OK;
	std::vector<NpBlock> npb((size_t)nblk);
	int64_t total = 0;
	for (int k = 0; k < nblk; k++) { npb[k].src_beg = vec[k].beg; npb[k].dst_beg = total; npb[k].n = (int32_t)(vec[k].end - vec[k].beg); npb[k].pad = 0; total += npb[k].n; }
	NpBlock *d_npb = ws.get<NpBlock>(nblk);
	int32_t *npcnt = ws.get<int32_t>(total + 1), *npoff = ws.get<int32_t>(total + 1);
	int64_t *d_fbeg = ws.get<int64_t>(nblk + 1);
	if (ws.rc) return ws.rc;
	CUDA_TRY(ctx, cudaMemcpyAsync(d_npb, npb.data(), (size_t)nblk * sizeof(NpBlock), cudaMemcpyHostToDevice, ctx->stream));
	LAUNCH(k_np_count, total, d_npb, nblk, total, cq, cr, cl, npcnt);
	CUDA_TRY(ctx, cudaMemsetAsync(npcnt + total, 0, 4, ctx->stream));
	GSA_TRY(scan_exclusive(ctx, npcnt, npoff, total + 1));
	int64_t nfr = 0;
	GSA_TRY(read_count(ctx, npoff + total, &nfr));
	GSA_TRY(gsa_ensure(ctx, ctx->d_frag, (size_t)(nfr + 1) * sizeof(gsa_frag)));
	GSA_TRY(gsa_ensure(ctx, ctx->d_fblk, (size_t)(nfr + 1) * 4));
	LAUNCH(k_np_write, total, d_npb, nblk, total, cq, cr, cl, npoff, (gsa_frag *)ctx->d_frag.p, (int32_t *)ctx->d_fblk.p, d_fbeg);
	GSA_TRY(gsa_ensure_host(ctx, ctx->h_stage, (size_t)nblk * 8)); // O(#blocks): a fixed-size buffer overflows on highly fragmented contigs
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_stage.p, d_fbeg, (size_t)nblk * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	const int64_t *fb = (const int64_t *)ctx->h_stage.p;
	for (int k = 0; k < nblk; k++) {
		ctx->final_blocks[k].frag_beg = fb[k];
		ctx->final_blocks[k].n_frags = (int32_t)((k + 1 < nblk ? fb[k + 1] : nfr) - fb[k]);
	}
	ctx->n_frags = nfr;
	return GSA_OK;
}

// ------------------------------------------------------------------------------------------------
// dump hook
// ------------------------------------------------------------------------------------------------
extern "C" int gsa_set_dump(gsa_ctx *ctx, int enable)
{
	if (!ctx) return GSA_ERR_ARG;
	ctx->keep_dumps = enable != 0;
	return GSA_OK;
}

extern "C" int64_t gsa_dump_blocks(gsa_ctx *ctx, int32_t stage, int64_t *out)
{
	if (!ctx || stage < 0 || stage > 3) return GSA_ERR_ARG;
	if (!ctx->have_cluster) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_dump_blocks: call gsa_cluster first");
	if (stage < 3 && !ctx->keep_dumps) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_dump_blocks: enable with gsa_set_dump before gsa_cluster");
	if (cudaSetDevice(ctx->device) != cudaSuccess) return GSA_ERR_CUDA;
	const std::vector<BlockHdr> &vec = stage == 3 ? ctx->final_blocks : ctx->blocks_stage[stage];
	int64_t words = 1;
	for (const BlockHdr &b : vec) words += 4 + 5 * (stage == 3 ? (int64_t)b.n_frags : b.end - b.beg);
	if (!out) return words;
	int64_t w = 0;
	out[w++] = (int64_t)vec.size();
	if (stage == 3) {
		std::vector<gsa_frag> fr((size_t)ctx->n_frags);
		if (ctx->n_frags && cudaMemcpy(fr.data(), ctx->d_frag.p, fr.size() * sizeof(gsa_frag), cudaMemcpyDeviceToHost) != cudaSuccess) return GSA_ERR_CUDA;
		for (const BlockHdr &b : vec) {
			out[w++] = b.score; out[w++] = 0; out[w++] = b.bDup; out[w++] = b.n_frags;
			for (int64_t t = b.frag_beg; t < b.frag_beg + b.n_frags; t++) { out[w++] = fr[t].bSeed; out[w++] = fr[t].qPos; out[w++] = fr[t].rPos; out[w++] = fr[t].qLen; out[w++] = fr[t].rLen; }
		}
		return w;
	}
	int64_t ns = stage == 0 ? ctx->n_s0 : ctx->n_cseeds;
	std::vector<int32_t> q((size_t)ns), l((size_t)ns); std::vector<int64_t> r((size_t)ns);
	const void *dq = stage == 0 ? ctx->d_s0q.p : ctx->d_cq.p, *dr = stage == 0 ? ctx->d_s0r.p : ctx->d_cr.p, *dl = stage == 0 ? ctx->d_s0l.p : ctx->d_cl.p;
	if (ns && (cudaMemcpy(q.data(), dq, (size_t)ns * 4, cudaMemcpyDeviceToHost) != cudaSuccess || cudaMemcpy(r.data(), dr, (size_t)ns * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
	           cudaMemcpy(l.data(), dl, (size_t)ns * 4, cudaMemcpyDeviceToHost) != cudaSuccess)) return GSA_ERR_CUDA;
	for (const BlockHdr &b : vec) {
		out[w++] = b.score; out[w++] = 0; out[w++] = 0; out[w++] = b.end - b.beg;
		for (int64_t t = b.beg; t < b.end; t++) { out[w++] = 1; out[w++] = q[t]; out[w++] = r[t]; out[w++] = l[t]; out[w++] = l[t]; }
	}
	return w;
}
