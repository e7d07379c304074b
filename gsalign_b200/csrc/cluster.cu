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
// (diagonal,qpos) / (group,qpos) and single-pass chained scans whose input and output are functors (scan.cuh: flag or
// count per seed -> prefix -> scatter / segment id / hash insert in ONE launch) over ALL groups at once, so the one giant
// main-diagonal group of a collinear contig costs the same as many small ones.  Element counts stay in device memory and
// the O(#blocks) logic runs in a kernel too, so the host waits ONCE per contig (final block list + fragment count):
//   * the greedy outlier windows (data-dependent resets) become "next window start" pointers computed
//     by binary search per seed and resolved by pointer jumping over the candidate list (one cooperative launch);
//   * the per-window PosDiff histograms become one global (window,bin) hash table with atomic counts;
//   * overlap trimming, gap / contig-span break points and normal-pair insertion are adjacent-pair maps.
// The reference's float / std::sort logic over the O(#blocks) headers is block_logic.cuh (k_block_logic); contigs the kernel
// declines take its host form, block_logic.cpp (two waits).
#include "fm.cuh"
#include "scan.cuh"
#include "block_logic.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>

// ------------------------------------------------------------------------------------------------
// scratch, device counters, chain launches
// ------------------------------------------------------------------------------------------------
// Every element count of the phase lives in device memory (dc[]): the kernels of one contig are queued back to back and
// sized by a host-side upper bound (the seed count), so the host only waits once, for the final block list and the fragment
// count (twice on the host path of the block logic: piece tables first).
enum { DC_N0 = 0, DC_NGROUPS, DC_N2, DC_NG2, DC_NC, DC_NLU, DC_N3, DC_N4, DC_NB0, DC_NB1, DC_N5, DC_N6A, DC_N6, DC_KILLS, DC_NCAND,
       DC_NP0, DC_NP1, DC_NP2, DC_NPL0, DC_NFR, DC_NPTOT, DC_COUNT };
#define K2_CHAINS 24

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

struct Chains { // the chain states of one gsa_cluster call: [tickets | totals | status slices], zeroed by one memset
	unsigned int *ticket; unsigned long long *total, *status; int64_t tiles; int used = 0;
	ChainState next() { ChainState c; c.ticket = ticket + used; c.total = total + used; c.status = status + (size_t)used * tiles; used++; return c; }
};

template <typename F>
static int run_chain(gsa_ctx *ctx, Chains &ch, const F &f, const int32_t *d_n, int64_t bound)
{
	if (ch.used >= K2_CHAINS) return gsa_fail(ctx, GSA_ERR_NOMEM, "cluster: out of chain states");
	if (bound + 1 > (ch.tiles - 1) * CH_TILE) return gsa_fail(ctx, GSA_ERR_ARG, "cluster: chain bound above the allocated tiles");
	k_chain<F><<<(unsigned)chain_tiles(bound), CH_THREADS, 0, ctx->stream>>>(f, d_n, ch.next());
	KERNEL_CHECK(ctx);
	return GSA_OK;
}

#define LAUNCH(kernel, n, ...)                                                                     \
	do {                                                                                           \
		if ((n) > 0) { kernel<<<gsa_grid((n), 256), 256, 0, ctx->stream>>>(__VA_ARGS__); KERNEL_CHECK(ctx); } \
	} while (0)

struct NoFinish { __device__ void finish(unsigned long long) const {} };

__global__ void k_k2_init(int32_t *dc, int32_t n0)
{
	if (threadIdx.x < DC_COUNT) dc[threadIdx.x] = threadIdx.x == DC_N0 ? n0 : 0;
}

// peers of the calling lane among the lanes with the same key; every lane of the warp must call (inactive lanes pass a key
// no active lane uses)
__device__ __forceinline__ unsigned peers_of(int key) { return __match_any_sync(0xffffffffu, key); }
#define DEAD_KEY (-1 - (int)(threadIdx.x & 31))

// ------------------------------------------------------------------------------------------------
// 1. diagonal groups (SeedGrouping, src/GSAlign.cpp:126-143) and their scores (FindSeedGroupScore, :298-303)
// ------------------------------------------------------------------------------------------------
struct FGroup {
	const int32_t *q; const int64_t *r; const int32_t *l; int32_t *gid1; unsigned long long *gscore; int32_t *dc; int max_indel;
	struct Item { int32_t len; uint8_t flag; };
	__device__ Item load(int64_t i) const
	{
		Item it; it.len = l[i];
		it.flag = i == 0 || ((r[i] - q[i]) - (r[i - 1] - q[i - 1])) > max_indel; // src/GSAlign.cpp:133
		return it;
	}
	__device__ unsigned long long value(const Item &it) const { return it.flag; }
	__device__ void emit(int64_t i, unsigned long long excl, const Item &it, bool valid) const
	{
		int g = valid ? (int)excl + it.flag : 0; // 1-based group id
		if (valid) gid1[i] = g;
		unsigned peers = peers_of(valid ? g : DEAD_KEY);
		bool leader;
		long long tot = gsa_peer_sum(peers, (long long)(valid ? it.len : 0), leader);
		if (leader && valid) atomicAdd(gscore + (g - 1), (unsigned long long)tot);
	}
	__device__ void finish(unsigned long long total) const { dc[DC_NGROUPS] = (int32_t)total; }
};

// groups below MinAlnBlockScore are skipped (src/GSAlign.cpp:387); the survivors get the sort key of the per-group order
// CompByQueryPos (src/ProcessCandidateAlignment.cpp:9-13): (group, qPos); ties on qPos keep their (PosDiff,qPos) input
// order under the stable radix sort, which for equal qPos is rPos order
struct FKeep {
	const int32_t *q, *gid1; const unsigned long long *gscore; uint64_t *key; int32_t *val; int32_t *dc; int min_score, qbits;
	struct Item { int32_t q, g; uint8_t keep; };
	__device__ Item load(int64_t i) const
	{
		Item it; it.q = q[i]; it.g = gid1[i] - 1;
		it.keep = (long long)gscore[it.g] >= (long long)min_score;
		return it;
	}
	__device__ unsigned long long value(const Item &it) const { return it.keep; }
	__device__ void emit(int64_t i, unsigned long long excl, const Item &it, bool valid) const
	{
		if (valid && it.keep) { key[excl] = ((uint64_t)(uint32_t)it.g << qbits) | (uint32_t)it.q; val[excl] = (int32_t)i; }
	}
	__device__ void finish(unsigned long long total) const { dc[DC_N2] = (int32_t)total; }
};

// gather in the sorted order + dense group ids + group start table
struct FSeg {
	const int32_t *val, *gid1, *q; const int64_t *r; const int32_t *l;
	int32_t *oq; int64_t *orr; int32_t *ol, *dg, *gstart, *dc;
	struct Item { int32_t q, l; int64_t r; uint8_t flag; };
	__device__ Item load(int64_t i) const
	{
		int32_t s = val[i];
		Item it; it.q = q[s]; it.r = r[s]; it.l = l[s];
		it.flag = i == 0 || gid1[s] != gid1[val[i - 1]];
		return it;
	}
	__device__ unsigned long long value(const Item &it) const { return it.flag; }
	__device__ void emit(int64_t i, unsigned long long excl, const Item &it, bool valid) const
	{
		if (!valid) return;
		oq[i] = it.q; orr[i] = it.r; ol[i] = it.l; dg[i] = (int32_t)excl + it.flag - 1;
		if (it.flag) gstart[excl] = (int32_t)i;
	}
	__device__ void finish(unsigned long long total) const { dc[DC_NG2] = (int32_t)total; gstart[total] = dc[DC_N2]; }
};

// ------------------------------------------------------------------------------------------------
// 2. SeedGroupAnalysis
// ------------------------------------------------------------------------------------------------
struct GroupView {
	const int32_t *q; const int64_t *r; const int32_t *l; const int32_t *dg; const int32_t *gstart; // gstart[ng] = n
	const int32_t *dn;                                                                              // element count (device)
};

__device__ __forceinline__ int64_t pd_of(const GroupView &v, int64_t i) { return v.r[i] - v.q[i]; }

// UniqueArr (src/GSAlign.cpp:316-325) and the positions where a window may close: unique, and PosDiff differs from the
// previous element (:328-331); U = inclusive count of unique seeds, C = list of the candidates
struct FUniq {
	GroupView v; int32_t *uq, *U; uint8_t *cand; int32_t *C, *cpos, *dc;   // cpos[i] = index of candidate i in C
	struct Item { uint8_t uq, cand; };
	__device__ Item load(int64_t i) const
	{
		const int64_t n = *v.dn;
		int gs = v.gstart[v.dg[i]], ge = v.gstart[v.dg[i] + 1];
		bool same_prev = i > gs && v.q[i - 1] == v.q[i], same_next = i + 1 < ge && i + 1 < n && v.q[i + 1] == v.q[i];
		Item it; it.uq = !(same_prev || same_next);
		it.cand = it.uq && i > gs && pd_of(v, i) != pd_of(v, i - 1);
		return it;
	}
	__device__ unsigned long long value(const Item &it) const { return CH_PACK2(it.uq, it.cand); }
	__device__ void emit(int64_t i, unsigned long long excl, const Item &it, bool valid) const
	{
		if (!valid) return;
		uq[i] = it.uq; U[i] = (int32_t)CH_LO(excl) + it.uq; cand[i] = it.cand;
		if (it.cand) { C[CH_HI(excl)] = (int32_t)i; cpos[i] = (int32_t)CH_HI(excl); }
	}
	__device__ void finish(unsigned long long total) const { dc[DC_NC] = (int32_t)CH_HI(total); }
};

// next window start after a window that starts at i (src/GSAlign.cpp:326-337): the first candidate j > i with
// (#unique in the window so far) >= 30 and q[j] - q[i] > 3000.  Only group starts and candidates can start a window, and a
// window's successor is always a candidate: the chain of window starts lives on the candidate list C.  So the successor is
// kept as an INDEX INTO C (nC = none): jump[c] for candidate c, and creach[] -- "candidate c starts a window" -- is seeded with
// the successors of the group starts.  reach[] (per seed) starts as the group starts.
__global__ void k_next(GroupView v, const int32_t *uq, const int32_t *U, const uint8_t *cand, const int32_t *C, const int32_t *cpos, const int32_t *d_nC,
                       int32_t *jump, int32_t *creach, int32_t *reach)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t n = *v.dn;
	if (i >= n) return;
	int gs = v.gstart[v.dg[i]], ge = v.gstart[v.dg[i] + 1];
	reach[i] = i == gs;
	if (i != gs && !cand[i]) return;
	const int nC = *d_nC;
	int baseU = i == gs ? U[i] - uq[i] : U[i]; // the first window counts its own first element, later ones restart at 0
	int lo = (int)i + 1, hi = ge;
	while (lo < hi) { int m = (lo + hi) >> 1; if (U[m] - baseU >= 30) hi = m; else lo = m + 1; }
	int jA = lo;
	lo = (int)i + 1; hi = ge;
	int qi = v.q[i];
	while (lo < hi) { int m = (lo + hi) >> 1; if (v.q[m] - qi > 3000) hi = m; else lo = m + 1; }
	int j0 = max(jA, lo);
	int res = nC;
	if (j0 < ge) {
		lo = 0; hi = nC;
		while (lo < hi) { int m = (lo + hi) >> 1; if (C[m] < j0) lo = m + 1; else hi = m; }
		if (lo < nC && C[lo] < ge) res = lo;
	}
	if (i == gs) { if (res < nC) creach[res] = 1; }
	else jump[cpos[i]] = res;
}

// Pointer jumping over the candidate list in one persistent launch: every round doubles the reach of jump[] (ping-ponged
// between two tables) and marks the candidates it lands on; creach is monotone and updated in place.  The rounds are
// separated by a grid-wide barrier (a counter in global memory; the launch is cooperative, so all CTAs are resident).
// At the end the marks go back to the seeds: reach[C[c]] = 1.
__global__ void __launch_bounds__(512) k_jump_all(int32_t *creach, int32_t *ja, int32_t *jb, const int32_t *C, const int32_t *d_nC, int32_t *reach, int rounds, unsigned int *bar)
{
	const int nC = *d_nC;
	const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
	for (int r = 0; r < rounds; r++) {
		for (int c = t0; c < nC; c += stride) {
			const int j = ja[c];
			int jj = nC;
			if (j < nC) { if (creach[c]) creach[j] = 1; jj = ja[j]; }
			jb[c] = jj;
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			__threadfence();
			atomicAdd(bar, 1u);
			const unsigned int want = (unsigned int)(r + 1) * gridDim.x;
			while (*(volatile unsigned int *)bar < want) { }
			__threadfence();
		}
		__syncthreads();
		int32_t *t = ja; ja = jb; jb = t;
	}
	for (int c = t0; c < nC; c += stride) if (creach[c]) reach[C[c]] = 1;
}

#define HASH_EMPTY 0xFFFFFFFFFFFFFFFFull
__device__ __forceinline__ uint32_t hash64(uint64_t k)
{
	k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
	return (uint32_t)k;
}

// window ids (inclusive count of window starts) and, in the same pass, the per-window PosDiff>>4 histogram (PDFmap,
// src/GSAlign.cpp:264-270) as a global (window,bin) hash table
struct FWin : NoFinish {
	GroupView v; const int32_t *uq, *reach; int32_t *wid1; unsigned long long *keys; int32_t *cnt, *slot_of; uint32_t hmask;
	struct Item { uint8_t r; };
	__device__ Item load(int64_t i) const { Item it; it.r = reach[i] != 0; return it; }
	__device__ unsigned long long value(const Item &it) const { return it.r; }
	__device__ void emit(int64_t i, unsigned long long excl, const Item &it, bool valid) const
	{
		int w1 = valid ? (int)excl + it.r : 0;
		if (valid) wid1[i] = w1;
		const bool act = valid && uq[i];
		// inactive lanes get distinct keys no (window,bin) can take (window ids are < 2^31)
		unsigned long long key = 0xFFFFFFFF00000000ull | (threadIdx.x & 31);
		if (act) key = ((unsigned long long)(uint32_t)(w1 - 1) << 32) | (uint32_t)(int32_t)(pd_of(v, i) >> 4);
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
};

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
	if (i < *v.dn && uq[i]) {
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

// outlier kill (src/GSAlign.cpp:282-294 with Check_PD_Frequency :145-153, Min_PD_Freq = 3) and the list of live unique seeds
struct FOutlier {
	GroupView v; const int32_t *uq, *wid1; const unsigned long long *best, *sum; const int32_t *cntk, *cnt, *slot_of;
	uint8_t *alive; int32_t *LU, *dc; int64_t genome; int max_indel;
	struct Item { uint8_t a, lu; };
	__device__ Item load(int64_t i) const
	{
		Item it; it.a = 1; it.lu = 0;
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
			if (diff > max_indel && own < 3) it.a = 0;
			it.lu = it.a;
		}
		return it;
	}
	__device__ unsigned long long value(const Item &it) const { return it.lu; }
	__device__ void emit(int64_t i, unsigned long long excl, const Item &it, bool valid) const
	{
		if (!valid) return;
		alive[i] = it.a;
		if (it.lu) LU[excl] = (int32_t)i;
	}
	__device__ void finish(unsigned long long total) const { dc[DC_NLU] = (int32_t)total; }
};

// multi-hit runs (same qPos): keep the hit nearest to the mean PosDiff of <= 5 + 5 neighbouring live unique seeds
// (src/GSAlign.cpp:341-350 with FindNeighboringPosDiffAvg :178-206 and RemoveRedundantSeeds :208-225)
__global__ void k_runs(GroupView v, const int32_t *LU, const int32_t *d_nLU, uint8_t *alive, int64_t genome, int max_indel)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= *v.dn) return;
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

// generic compaction of (q, r, l, tag) by a byte flag
struct FCompact {
	const uint8_t *alive; const int32_t *q; const int64_t *r; const int32_t *l, *g; int32_t *oq; int64_t *orr; int32_t *ol, *og, *dc; int slot;
	struct Item { uint8_t a; };
	__device__ Item load(int64_t i) const { Item it; it.a = alive[i]; return it; }
	__device__ unsigned long long value(const Item &it) const { return it.a; }
	__device__ void emit(int64_t i, unsigned long long excl, const Item &it, bool valid) const
	{
		if (valid && it.a) { oq[excl] = q[i]; orr[excl] = r[i]; ol[excl] = l[i]; og[excl] = g[i]; }
	}
	__device__ void finish(unsigned long long total) const { dc[slot] = (int32_t)total; }
};

// noise: interior seed whose PosDiff is > 5 away from both live neighbours of its group (src/GSAlign.cpp:355-362); compacts
struct FNoise {
	const int32_t *q; const int64_t *r; const int32_t *l, *g; const int32_t *dn; int32_t *oq; int64_t *orr; int32_t *ol, *og, *dc;
	struct Item { int32_t q, l, g; int64_t r; uint8_t a; };
	__device__ Item load(int64_t i) const
	{
		const int64_t n = *dn;
		Item it; it.q = q[i]; it.r = r[i]; it.l = l[i]; it.g = g[i]; it.a = 1;
		if (i > 0 && i + 1 < n && g[i - 1] == it.g && g[i + 1] == it.g) {
			int64_t p = it.r - it.q, a1 = p - (r[i - 1] - q[i - 1]), a2 = p - (r[i + 1] - q[i + 1]);
			if (a1 < 0) a1 = -a1;
			if (a2 < 0) a2 = -a2;
			if (a1 > 5 && a2 > 5) it.a = 0;
		}
		return it;
	}
	__device__ unsigned long long value(const Item &it) const { return it.a; }
	__device__ void emit(int64_t i, unsigned long long excl, const Item &it, bool valid) const
	{
		if (valid && it.a) { oq[excl] = it.q; orr[excl] = it.r; ol[excl] = it.l; og[excl] = it.g; }
	}
	__device__ void finish(unsigned long long total) const { dc[DC_N4] = (int32_t)total; }
};

// block cuts inside a group (src/GSAlign.cpp:364-374): block id per seed, block start table, block scores (sum of lengths)
struct FCut {
	const int32_t *q; const int64_t *r; const int32_t *l, *g; int32_t *bid, *bstart, *bscore, *dc;
	struct Item { int32_t len; uint8_t cut; };
	__device__ Item load(int64_t i) const
	{
		Item it; it.len = l[i];
		bool cut = i == 0 || g[i] != g[i - 1];
		if (!cut) {
			int64_t d = (r[i - 1] - q[i - 1]) - (r[i] - q[i]);
			if (d < 0) d = -d;
			cut = q[i] - q[i - 1] - l[i - 1] > GSA_MAX_SEED_GAP || d > 100;
		}
		it.cut = cut;
		return it;
	}
	__device__ unsigned long long value(const Item &it) const { return it.cut; }
	__device__ void emit(int64_t i, unsigned long long excl, const Item &it, bool valid) const
	{
		int b = valid ? (int)excl + it.cut - 1 : 0;
		if (valid) { bid[i] = b; if (it.cut) bstart[excl] = (int32_t)i; }
		unsigned peers = peers_of(valid ? b : DEAD_KEY);
		bool leader;
		int tot = gsa_peer_sum(peers, valid ? it.len : 0, leader);
		if (leader && valid) atomicAdd(bscore + b, tot);
	}
	__device__ void finish(unsigned long long total) const { dc[DC_NB0] = (int32_t)total; bstart[total] = dc[DC_N4]; }
};

// AddAlnBlock acceptance (src/GSAlign.cpp:29-49): new (dense) ids of the accepted blocks and their scores
struct FBlockEval {
	const int32_t *bstart, *q, *l, *bscore; uint8_t *accept; int32_t *newid1, *kept_score, *dc; int min_score, min_len;
	struct Item { int32_t sc; uint8_t a; };
	__device__ Item load(int64_t b) const
	{
		int64_t beg = bstart[b], end = bstart[b + 1];
		Item it; it.sc = bscore[b];
		int32_t region = (q[end - 1] + l[end - 1]) - q[beg];
		it.a = !(it.sc < min_score || region < min_len || (it.sc < 1000 && it.sc < region * 0.05));
		return it;
	}
	__device__ unsigned long long value(const Item &it) const { return it.a; }
	__device__ void emit(int64_t b, unsigned long long excl, const Item &it, bool valid) const
	{
		if (!valid) return;
		accept[b] = it.a; newid1[b] = (int32_t)excl + it.a;
		if (it.a) kept_score[excl] = it.sc;
	}
	__device__ void finish(unsigned long long total) const { dc[DC_NB1] = (int32_t)total; }
};

struct FSeedAccept {
	const int32_t *bid; const uint8_t *accept; const int32_t *newid1, *q; const int64_t *r; const int32_t *l; int32_t *oq; int64_t *orr; int32_t *ol, *ob, *dc;
	struct Item { int32_t b; uint8_t a; };
	__device__ Item load(int64_t i) const { Item it; it.b = bid[i]; it.a = accept[it.b]; return it; }
	__device__ unsigned long long value(const Item &it) const { return it.a; }
	__device__ void emit(int64_t i, unsigned long long excl, const Item &it, bool valid) const
	{
		if (valid && it.a) { oq[excl] = q[i]; orr[excl] = r[i]; ol[excl] = l[i]; ob[excl] = newid1[it.b] - 1; }
	}
	__device__ void finish(unsigned long long total) const { dc[DC_N5] = (int32_t)total; }
};

// ------------------------------------------------------------------------------------------------
// 3. RemoveOverlaps, gap / span break points, pieces
// ------------------------------------------------------------------------------------------------
// one pass of RemoveOverlaps (src/ProcessCandidateAlignment.cpp:197-227): element i only reads the untouched q/r of i+1, so a
// pass is elementwise; the survivors are compacted into the other buffer
struct FOverlap {
	const int32_t *q; const int64_t *r; const int32_t *l, *b; const int32_t *dn; int32_t *oq; int64_t *orr; int32_t *ol, *ob, *dc; int slot_n;
	struct Item { int32_t q, len, b; int64_t r; uint8_t a; };
	__device__ Item load(int64_t i) const
	{
		const int64_t n = *dn;
		Item it; it.q = q[i]; it.r = r[i]; it.len = l[i]; it.b = b[i]; it.a = 1;
		if (i + 1 < n && b[i + 1] == it.b) {
			if (r[i + 1] <= it.r) it.a = 0;
			else {
				int32_t len = it.len, ov = (int32_t)(it.r + len - r[i + 1]);
				if (ov > 0) { len -= ov; if (len <= 0) it.a = 0; }
				if (it.a && (ov = it.q + len - q[i + 1]) > 0) { len -= ov; if (len <= 0) it.a = 0; }
				it.len = len;
			}
		}
		return it;
	}
	__device__ unsigned long long value(const Item &it) const { return it.a; }
	__device__ void emit(int64_t i, unsigned long long excl, const Item &it, bool valid) const
	{
		if (valid && it.a) { oq[excl] = it.q; orr[excl] = it.r; ol[excl] = it.len; ob[excl] = it.b; }
	}
	__device__ void finish(unsigned long long total) const { dc[DC_KILLS] = *dn - (int32_t)total; dc[slot_n] = (int32_t)total; }
};

__device__ __forceinline__ int64_t contig_end_of(const ContigEnd *ce, int nce, int64_t rpos)
{
	int lo = 0, hi = nce;
	while (lo < hi) { int m = (lo + hi) >> 1; if (ce[m].end < rpos) lo = m + 1; else hi = m; }
	return ce[lo < nce ? lo : nce - 1].end;
}

// break-point flags per seed -- bit0: certain gap break, bit1: contig-span break, bit2: gap needs the similarity test -- and
// the list of the gaps to test
struct FGapFlags {
	const int32_t *q; const int64_t *r; const int32_t *l, *b; const ContigEnd *ce; int nce; uint8_t *flag; int32_t *cand, *dc;
	struct Item { uint8_t f; };
	__device__ Item load(int64_t i) const
	{
		Item it; it.f = 0;
		if (i > 0 && b[i] == b[i - 1]) {
			int32_t qg = q[i] - q[i - 1] - l[i - 1], rg = (int32_t)(r[i] - r[i - 1] - l[i - 1]);
			if (qg > 300 || rg > 300) it.f |= (qg > GSA_MAX_SEED_GAP || rg > GSA_MAX_SEED_GAP) ? 1 : 4;
			if (contig_end_of(ce, nce, r[i]) != contig_end_of(ce, nce, r[i - 1])) it.f |= 2;
		}
		return it;
	}
	__device__ unsigned long long value(const Item &it) const { return (it.f >> 2) & 1; }
	__device__ void emit(int64_t i, unsigned long long excl, const Item &it, bool valid) const
	{
		if (!valid) return;
		flag[i] = it.f;
		if (it.f & 4) cand[excl] = (int32_t)i;
	}
	__device__ void finish(unsigned long long total) const { dc[DC_NCAND] = (int32_t)total; }
};

// CalGapSimilarity (src/KmerAnalysis.cpp:78-121) for the gap in front of seed cand[blockIdx.x]; one warp per gap.
// hist: ids of CreateKmerVecFromReadSeq are < 2048 (rolling ((id & 0xFF) << 2) + nt with nt in 0..4).
// An id only depends on the five bases it ends on (each step keeps 8 bits of the previous id and shifts them by 2), so the
// lanes take the positions of a gap side by side and count into shared-memory histograms with atomics.  The reference's
// quirks on the query side (only the byte 'N' restarts the window, the window head is stale after a restart, any other
// non-ACGT letter enters the id as 4 and carries) only show when the gap holds a letter outside ACGT: those gaps -- rare --
// are walked by one lane exactly like the reference does.
__global__ void __launch_bounds__(32) k_gap_similarity(const int32_t *cand, const int32_t *d_ncand, const int32_t *q, const int64_t *r, const int32_t *l,
                                                       const unsigned char *seq, DevIndex ix, uint8_t *flag)
{
	__shared__ unsigned int h1[2048], h2[2048];
	const int lane = threadIdx.x, ncand = *d_ncand;
	for (int c = blockIdx.x; c < ncand; c += gridDim.x) { // a fixed grid of warps walks the list: its length only exists on the device
	int i = cand[c];
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
		const unsigned char *s = seq + q1;
		bool other = false;
		for (int k = lane; k < q_len; k += 32) other |= gsa_nt4(s[k]) == 4;
		other = __any_sync(0xffffffffu, other);
		__syncwarp();
		// reference k-mers (no N in the text): the 5-mer starting at k is the top 10 bits of the 16-base window there
		for (int k = lane; k + 5 <= r_len; k += 32) atomicAdd(&h2[gsa_pk_window(ix.txt, (uint64_t)(r1 + k)) >> 22], 1u);
		if (!other) { // query k-mers, ACGT only: plain 5-mers
			for (int k = lane; k + 5 <= q_len; k += 32) {
				uint32_t wid = 0;
#pragma unroll
				for (int j = 0; j < 5; j++) wid = (wid << 2) + (uint32_t)gsa_nt4(s[k + j]);
				atomicAdd(&h1[wid], 1u);
			}
		} else if (lane == 0) { // quirks kept: only the byte 'N' restarts, stale head after a restart
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
		}
		__syncwarp();
		int common = 0;
		for (int k = lane; k < 2048; k += 32) common += (int)min(h1[k], h2[k]);
		for (int o = 16; o > 0; o >>= 1) common += __shfl_xor_sync(0xffffffffu, common, o);
		similar = common > (q_len + r_len) * 0.1;
	}
	if (lane == 0 && !similar) flag[i] |= 1;
	__syncwarp();
	}
}


// pieces: maximal runs of seeds between break points (block boundaries and the flags selected by mask); start table + sums
struct FPiece {
	const int32_t *b, *l; const uint8_t *flag; uint8_t mask; const int32_t *dn; int32_t *pstart, *psum, *dc; int slot;
	struct Item { int32_t len; uint8_t f; };
	__device__ Item load(int64_t i) const
	{
		Item it; it.len = l[i];
		it.f = i == 0 || b[i] != b[i - 1] || (mask && (flag[i] & mask));
		return it;
	}
	__device__ unsigned long long value(const Item &it) const { return it.f; }
	__device__ void emit(int64_t i, unsigned long long excl, const Item &it, bool valid) const
	{
		int p = valid ? (int)excl + it.f - 1 : 0;
		if (valid && it.f) pstart[excl] = (int32_t)i;
		unsigned peers = peers_of(valid ? p : DEAD_KEY);
		bool leader;
		int tot = gsa_peer_sum(peers, valid ? it.len : 0, leader);
		if (leader && valid) atomicAdd(psum + p, tot);
	}
	__device__ void finish(unsigned long long total) const { dc[slot] = (int32_t)total; pstart[total] = *dn; }
};

__global__ void k_piece_table(const int32_t *starts, const int32_t *psum, const int32_t *d_np, const int32_t *q, const int64_t *r, const int32_t *l, Piece *out)
{
	int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= *d_np) return;
	int64_t beg = starts[p], end = starts[p + 1];
	Piece x; x.beg = beg; x.end = end; x.sumlen = psum[p]; x.rf = r[beg]; x.rl = r[end - 1];
	x.qf = q[beg]; x.ql = q[end - 1]; x.lenl = l[end - 1]; x.pad = 0;
	out[p] = x;
}

// ------------------------------------------------------------------------------------------------
// 4. IdentifyNormalPairs -> fragment list (src/ProcessCandidateAlignment.cpp:241-265)
// ------------------------------------------------------------------------------------------------
struct NpBlock { int64_t src_beg, dst_beg; int32_t n, pad; };

// element t of the concatenated kept blocks -> (block k, source seed s)
__device__ __forceinline__ void np_locate(const NpBlock *nb, int nblk, int64_t t, int &k, int64_t &s)
{
	int lo = 0, hi = nblk;
	while (lo < hi) { int m = (lo + hi) >> 1; if (nb[m].dst_beg <= t) lo = m + 1; else hi = m; }
	k = lo - 1; s = nb[k].src_beg + (t - nb[k].dst_beg);
}

struct FNormalPairs {
	const NpBlock *nb; int nblk; const int32_t *d_nblk;   // the block count: on the host, or (d_nblk != null) where the device block logic left it
	const int32_t *q; const int64_t *r; const int32_t *l; gsa_frag *frag; int32_t *fblk; int64_t *blk_frag_beg; int32_t *dc;
	struct Item { int32_t k, q, l, qg, rg; int64_t r; uint8_t first; };
	__device__ Item load(int64_t t) const
	{
		Item it; int64_t s;
		np_locate(nb, d_nblk ? *d_nblk : nblk, t, it.k, s);
		it.q = q[s]; it.r = r[s]; it.l = l[s]; it.qg = it.rg = 0; it.first = t == nb[it.k].dst_beg;
		if (t + 1 < nb[it.k].dst_beg + nb[it.k].n) { // not the last seed of its block
			it.qg = q[s + 1] - (it.q + it.l); it.rg = (int32_t)(r[s + 1] - (it.r + it.l));
		}
		return it;
	}
	__device__ unsigned long long value(const Item &it) const { return (it.qg > 0 || it.rg > 0) ? 2 : 1; }
	__device__ void emit(int64_t t, unsigned long long excl, const Item &it, bool valid) const
	{
		if (!valid) return;
		const int64_t o = (int64_t)excl;
		if (it.first) blk_frag_beg[it.k] = o;
		gsa_frag f; f.rPos = it.r; f.qPos = it.q; f.qLen = it.l; f.rLen = it.l; f.bSeed = 1; f.aln_off = 0; f.aln_len = it.l; f.reserved = 0;
		frag[o] = f; fblk[o] = it.k;
		if (it.qg > 0 || it.rg > 0) {
			gsa_frag g; g.rPos = it.r + it.l; g.qPos = it.q + it.l; g.qLen = max(it.qg, 0); g.rLen = max(it.rg, 0); g.bSeed = 0; g.aln_off = 0; g.aln_len = 0; g.reserved = 0;
			frag[o + 1] = g; fblk[o + 1] = it.k;
		}
	}
	__device__ void finish(unsigned long long total) const { dc[DC_NFR] = (int32_t)total; }
};

// ------------------------------------------------------------------------------------------------
// 5. the block logic on the device (SURVEY rows A8 tail + N4): block headers from the piece tables, the two split phases,
// EstChromosomeSimilarity + RemoveRedundantAlnBlocks, and the table IdentifyNormalPairs works from -- block_logic.cuh, the
// same statements as the host path (block_logic.cpp) with libstdc++'s introsort restated (stdsort.cuh).  One thread: the
// work is O(#blocks) with a handful of blocks per contig and the order of every step is part of the result.  It exists so
// that the phase needs no host round trip between the piece tables and the fragment list: everything after RemoveOverlaps
// is queued blindly and the host looks at the outcome once, at the end.  A contig with more blocks than the kernel's list
// holds, or whose RemoveOverlaps needed another round, takes the host path instead.
// ------------------------------------------------------------------------------------------------
#define BLK_DEV_CAP 1024
enum { BR_NFINAL = 0, BR_STATUS, BR_HAZARD, BR_N1, BR_N2, BR_COUNT = 8 };   // res[]: status 0 ok, 1 list too small, 2 inconsistent counts
struct BlkLogicArgs {
	const Piece *pt0, *pt1, *pt2; const int32_t *kept_score; int32_t *dc;
	BlockHdr *vec; NpBlock *npb; BlkParams P;
	BlockHdr *stage1, *stage2;      // snapshots for the dump hook (may be null)
	int32_t *res;
	int cap;                        // blocks the list may hold (<= BLK_DEV_CAP; tests lower it to reach the declining path)
};

__global__ void k_block_logic(BlkLogicArgs A)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	int32_t *res = A.res;
	for (int i = 0; i < BR_COUNT; i++) res[i] = 0;
	A.dc[DC_N0] = 0;
	const int np0 = A.dc[DC_NP0], np1 = A.dc[DC_NP1], np2 = A.dc[DC_NP2];
	if (np0 != A.dc[DC_NB1]) { res[BR_STATUS] = 2; return; }   // RemoveOverlaps must not change the number of blocks
	if (np0 > A.cap) { res[BR_STATUS] = 1; return; }
	for (int i = 0; i < np0; i++) A.vec[i] = blk_from_piece(A.pt0[i], A.kept_score[i]); // blocks keep their pre-overlap score
	res[BR_N1] = np0;
	if (A.stage1) for (int i = 0; i < np0; i++) A.stage1[i] = A.vec[i];
	int hz = 0;
	int n = blk_split(A.P, A.vec, np0, A.cap, A.pt1, np1, &hz);      // CheckAlnBlockLargeGaps + RemoveBadAlnBlocks
	if (n >= 0) n = blk_split(A.P, A.vec, n, A.cap, A.pt2, np2, &hz); // CheckAlnBlockSpanMultiSeqs + RemoveBadAlnBlocks
	if (n < 0) { res[BR_STATUS] = 1; return; }
	res[BR_N2] = n; res[BR_HAZARD] = hz;
	if (A.stage2) for (int i = 0; i < n; i++) A.stage2[i] = A.vec[i];
	n = blk_dedup(A.P, A.vec, n);
	long long total = 0;
	for (int k = 0; k < n; k++) {
		NpBlock b; b.src_beg = A.vec[k].beg; b.dst_beg = total; b.n = (int32_t)(A.vec[k].end - A.vec[k].beg); b.pad = 0;
		A.npb[k] = b; total += b.n;
	}
	if (total >= 0x3FFFFFF0ll) { res[BR_STATUS] = 2; return; }
	res[BR_NFINAL] = n;
	A.dc[DC_N0] = (int32_t)total;   // element count of the IdentifyNormalPairs chain
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

#define PIECE_FIRST 4096   // piece-table entries copied to the host before their count is known

struct PieceTable { Piece *d = nullptr; int slot = 0; };

// queues the piece table of (q, r, l, b) under `mask`: one chain + one table kernel; nothing comes to the host yet
static int queue_pieces(gsa_ctx *ctx, Ws &ws, Chains &ch, int32_t *dc, const int32_t *dn, int64_t bound, const int32_t *q, const int64_t *r, const int32_t *l,
                        const int32_t *b, const uint8_t *gflag, uint8_t mask, int slot, PieceTable &pt)
{
	int32_t *pstart = ws.get<int32_t>(bound + 2), *psum = ws.get<int32_t>(bound + 2);
	pt.d = ws.get<Piece>(bound + 1); pt.slot = slot;
	if (ws.rc) return ws.rc;
	CUDA_TRY(ctx, cudaMemsetAsync(psum, 0, (size_t)(bound + 2) * 4, ctx->stream));
	FPiece f; f.b = b; f.l = l; f.flag = gflag; f.mask = mask; f.dn = dn; f.pstart = pstart; f.psum = psum; f.dc = dc; f.slot = slot;
	GSA_TRY(run_chain(ctx, ch, f, dn, bound));
	LAUNCH(k_piece_table, bound, pstart, psum, dc + slot, q, r, l, pt.d);
	return GSA_OK;
}

// after a synchronisation that brought dc[] to the host: the table itself (the first PIECE_FIRST entries were prefetched)
static int fetch_piece_table(gsa_ctx *ctx, const PieceTable &pt, int64_t np, const Piece *prefetched, std::vector<Piece> &out)
{
	out.resize((size_t)np);
	if (np == 0) return GSA_OK;
	if (np <= PIECE_FIRST) { memcpy(out.data(), prefetched, (size_t)np * sizeof(Piece)); return GSA_OK; }
	GSA_TRY(gsa_ensure_host(ctx, ctx->h_stage, (size_t)np * sizeof(Piece)));
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_stage.p, pt.d, (size_t)np * sizeof(Piece), cudaMemcpyDeviceToHost, ctx->stream));
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
	const int64_t n = ctx->n_seeds;
	if (n >= 0x7FFFFFF0ll) return gsa_fail(ctx, GSA_ERR_LIMIT, "gsa_cluster: more than 2^31 seeds in one contig");
	if (n == 0) return GSA_OK;
	const int32_t *sq = (const int32_t *)ctx->d_sq.p; const int64_t *sr = (const int64_t *)ctx->d_sr.p; const int32_t *sl = (const int32_t *)ctx->d_sl.p;
	GSA_TRY(gsa_ensure(ctx, ctx->d_counter, 4096));
	int32_t *dc = (int32_t *)ctx->d_counter.p + 512; // byte 2048 onwards: the K2 counters
	Chains ch;
	ch.tiles = chain_tiles(n + 2);
	{
		const size_t words = (size_t)K2_CHAINS * 2 + (size_t)K2_CHAINS * ch.tiles; // tickets (u32, padded to u64) + totals + status
		GSA_TRY(gsa_ensure(ctx, ctx->d_chain, words * 8));
		CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_chain.p, 0, words * 8, ctx->stream));
		ch.ticket = (unsigned int *)ctx->d_chain.p;
		ch.total = (unsigned long long *)ctx->d_chain.p + K2_CHAINS;
		ch.status = ch.total + K2_CHAINS;
	}
	k_k2_init<<<1, 64, 0, ctx->stream>>>(dc, (int32_t)n);
	KERNEL_CHECK(ctx);

	// ---- 1. diagonal groups over the (PosDiff,qPos)-sorted seeds; drop groups below MinAlnBlockScore; per-group order --------
	int32_t *gid1 = ws.get<int32_t>(n + 2);
	unsigned long long *gscore = ws.get<unsigned long long>(n + 2);
	uint64_t *key_in = ws.get<uint64_t>(n + 2), *key_out = ws.get<uint64_t>(n + 2);
	int32_t *val_in = ws.get<int32_t>(n + 2), *val_out = ws.get<int32_t>(n + 2);
	if (ws.rc) return ws.rc;
	CUDA_TRY(ctx, cudaMemsetAsync(gscore, 0, (size_t)(n + 2) * 8, ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(key_in, 0xFF, (size_t)(n + 2) * 8, ctx->stream)); // dropped seeds sort behind every kept one
	{ FGroup f; f.q = sq; f.r = sr; f.l = sl; f.gid1 = gid1; f.gscore = gscore; f.dc = dc; f.max_indel = P.max_indel; GSA_TRY(run_chain(ctx, ch, f, dc + DC_N0, n)); }
	int qbits = 1, gbits = 1;
	while ((1ull << qbits) < (uint64_t)ctx->qlen) qbits++;
	while ((1ll << gbits) < n + 1) gbits++;
	{ FKeep f; f.q = sq; f.gid1 = gid1; f.gscore = gscore; f.key = key_in; f.val = val_in; f.dc = dc; f.min_score = P.min_block_score; f.qbits = qbits; GSA_TRY(run_chain(ctx, ch, f, dc + DC_N0, n)); }
	{
		size_t bytes = 0;
		cub::DeviceRadixSort::SortPairs(nullptr, bytes, key_in, key_out, val_in, val_out, n, 0, qbits + gbits, ctx->stream);
		GSA_TRY(gsa_ensure(ctx, ctx->d_cub, bytes));
		CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->d_cub.p, bytes, key_in, key_out, val_in, val_out, n, 0, qbits + gbits, ctx->stream));
		ctx->tm.launches += 2 + (qbits + gbits + 7) / 8;
	}
	int32_t *q2 = ws.get<int32_t>(n + 2), *l2 = ws.get<int32_t>(n + 2), *dg = ws.get<int32_t>(n + 2), *gstart = ws.get<int32_t>(n + 3);
	int64_t *r2 = ws.get<int64_t>(n + 2);
	if (ws.rc) return ws.rc;
	{ FSeg f; f.val = val_out; f.gid1 = gid1; f.q = sq; f.r = sr; f.l = sl; f.oq = q2; f.orr = r2; f.ol = l2; f.dg = dg; f.gstart = gstart; f.dc = dc; GSA_TRY(run_chain(ctx, ch, f, dc + DC_N2, n)); }
	GroupView gv; gv.q = q2; gv.r = r2; gv.l = l2; gv.dg = dg; gv.gstart = gstart; gv.dn = dc + DC_N2;

	// ---- 2. outlier windows --------------------------------------------------------------------------------------------------------
	int32_t *uq = ws.get<int32_t>(n + 2), *U = ws.get<int32_t>(n + 2), *C = ws.get<int32_t>(n + 2);
	uint8_t *cand = ws.get<uint8_t>(n + 2);
	int32_t *nxtA = ws.get<int32_t>(n + 2), *nxtB = ws.get<int32_t>(n + 2), *reach = ws.get<int32_t>(n + 2), *wid1 = ws.get<int32_t>(n + 2);
	if (ws.rc) return ws.rc;
	int32_t *cpos = ws.get<int32_t>(n + 2), *creach = ws.get<int32_t>(n + 2);
	if (ws.rc) return ws.rc;
	{ FUniq f; f.v = gv; f.uq = uq; f.U = U; f.cand = cand; f.C = C; f.cpos = cpos; f.dc = dc; GSA_TRY(run_chain(ctx, ch, f, dc + DC_N2, n)); }
	CUDA_TRY(ctx, cudaMemsetAsync(creach, 0, (size_t)(n + 2) * 4, ctx->stream));
	LAUNCH(k_next, n, gv, uq, U, cand, C, cpos, dc + DC_NC, nxtA, creach, reach);
	{
		// a window holds at least 30 unique seeds: a chain has at most n / 30 + 1 links
		int rounds = 1; while ((1ll << rounds) < n / 30 + 2) rounds++;
		rounds++;
		unsigned int *bar = (unsigned int *)ctx->d_counter.p + 96;
		CUDA_TRY(ctx, cudaMemsetAsync(bar, 0, 4, ctx->stream));
		static int jump_ctas = 0;   // CTAs of k_jump_all that are resident at once (the grid barrier needs them all)
		if (jump_ctas == 0) {
			int per_sm = 0, sms = 0;
			CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_jump_all, 512, 0));
			CUDA_TRY(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
			jump_ctas = std::max(1, std::min(per_sm, 1) * sms);
		}
		const int32_t *Cc = C, *dnc = dc + DC_NC;
		void *args[] = {&creach, &nxtA, &nxtB, &Cc, &dnc, &reach, &rounds, &bar};
		CUDA_TRY(ctx, cudaLaunchCooperativeKernel((const void *)k_jump_all, dim3((unsigned)jump_ctas), dim3(512), args, 0, ctx->stream));
		KERNEL_CHECK(ctx);
	}
	// per-window histogram of PosDiff>>4 over unique seeds -> mode, average, outlier kill
	int64_t hsize = 1024; while (hsize < 2 * n) hsize <<= 1;
	unsigned long long *hkeys = ws.get<unsigned long long>(hsize);
	int32_t *hcnt = ws.get<int32_t>(hsize), *slot_of = ws.get<int32_t>(n + 2);
	unsigned long long *wbest = ws.get<unsigned long long>(n + 2), *wsum = ws.get<unsigned long long>(n + 2);
	int32_t *wcnt = ws.get<int32_t>(n + 2), *LU = ws.get<int32_t>(n + 2);
	uint8_t *alive = ws.get<uint8_t>(n + 2);
	if (ws.rc) return ws.rc;
	CUDA_TRY(ctx, cudaMemsetAsync(hkeys, 0xFF, (size_t)hsize * 8, ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(hcnt, 0, (size_t)hsize * 4, ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(wbest, 0, (size_t)(n + 2) * 8, ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(wsum, 0, (size_t)(n + 2) * 8, ctx->stream));
	CUDA_TRY(ctx, cudaMemsetAsync(wcnt, 0, (size_t)(n + 2) * 4, ctx->stream));
	{ FWin f; f.v = gv; f.uq = uq; f.reach = reach; f.wid1 = wid1; f.keys = hkeys; f.cnt = hcnt; f.slot_of = slot_of; f.hmask = (uint32_t)(hsize - 1); GSA_TRY(run_chain(ctx, ch, f, dc + DC_N2, n)); }
	LAUNCH(k_win_best, hsize, hkeys, hcnt, wbest, hsize);
	LAUNCH(k_win_sum, n, gv, uq, wid1, wbest, wsum, wcnt);
	{ FOutlier f; f.v = gv; f.uq = uq; f.wid1 = wid1; f.best = wbest; f.sum = wsum; f.cntk = wcnt; f.cnt = hcnt; f.slot_of = slot_of; f.alive = alive; f.LU = LU; f.dc = dc;
	  f.genome = ctx->N; f.max_indel = P.max_indel; GSA_TRY(run_chain(ctx, ch, f, dc + DC_N2, n)); }
	// multi-hit runs
	LAUNCH(k_runs, n, gv, LU, dc + DC_NLU, alive, ctx->N, P.max_indel);

	// ---- 3. compact, noise filter + compact, cut groups into blocks, AddAlnBlock acceptance -----------------------------------------
	int32_t *q3 = ws.get<int32_t>(n + 2), *l3 = ws.get<int32_t>(n + 2), *g3 = ws.get<int32_t>(n + 2);
	int64_t *r3 = ws.get<int64_t>(n + 2);
	if (ws.rc) return ws.rc;
	{ FCompact f; f.alive = alive; f.q = q2; f.r = r2; f.l = l2; f.g = dg; f.oq = q3; f.orr = r3; f.ol = l3; f.og = g3; f.dc = dc; f.slot = DC_N3; GSA_TRY(run_chain(ctx, ch, f, dc + DC_N2, n)); }
	// the n2 generation is dead from here on: its arrays carry the n4 generation
	int32_t *q4 = q2, *l4 = l2, *g4 = uq; int64_t *r4 = r2;
	{ FNoise f; f.q = q3; f.r = r3; f.l = l3; f.g = g3; f.dn = dc + DC_N3; f.oq = q4; f.orr = r4; f.ol = l4; f.og = g4; f.dc = dc; GSA_TRY(run_chain(ctx, ch, f, dc + DC_N3, n)); }
	int32_t *bid = U, *bstart = C, *bscore = wcnt, *newid1 = nxtA, *kept_score = nxtB; uint8_t *accept = cand;
	CUDA_TRY(ctx, cudaMemsetAsync(bscore, 0, (size_t)(n + 2) * 4, ctx->stream));
	{ FCut f; f.q = q4; f.r = r4; f.l = l4; f.g = g4; f.bid = bid; f.bstart = bstart; f.bscore = bscore; f.dc = dc; GSA_TRY(run_chain(ctx, ch, f, dc + DC_N4, n)); }
	{ FBlockEval f; f.bstart = bstart; f.q = q4; f.l = l4; f.bscore = bscore; f.accept = accept; f.newid1 = newid1; f.kept_score = kept_score; f.dc = dc;
	  f.min_score = P.min_block_score; f.min_len = P.min_aln_len; GSA_TRY(run_chain(ctx, ch, f, dc + DC_NB0, n)); }
	// the working set of the remaining phases lives in the context (K3 and the dump hooks read it)
	GSA_TRY(gsa_ensure(ctx, ctx->d_cq, (size_t)(n + 2) * 4)); GSA_TRY(gsa_ensure(ctx, ctx->d_cr, (size_t)(n + 2) * 8));
	GSA_TRY(gsa_ensure(ctx, ctx->d_cl, (size_t)(n + 2) * 4)); GSA_TRY(gsa_ensure(ctx, ctx->d_cb, (size_t)(n + 2) * 4));
	int32_t *cq = (int32_t *)ctx->d_cq.p, *cl = (int32_t *)ctx->d_cl.p, *cb = (int32_t *)ctx->d_cb.p; int64_t *cr = (int64_t *)ctx->d_cr.p;
	{ FSeedAccept f; f.bid = bid; f.accept = accept; f.newid1 = newid1; f.q = q4; f.r = r4; f.l = l4; f.oq = cq; f.orr = cr; f.ol = cl; f.ob = cb; f.dc = dc; GSA_TRY(run_chain(ctx, ch, f, dc + DC_N4, n)); }

	int32_t *hc = (int32_t *)ctx->h_small.p;                       // dc[] on the host after a synchronisation
	Piece *h_first = (Piece *)((char *)ctx->h_small.p + 4096);       // 3 x PIECE_FIRST prefetched piece-table entries
	static_assert(4096 + 3 * PIECE_FIRST * sizeof(Piece) <= (1u << 20), "h_small too small for the prefetched piece tables");
	std::vector<Piece> pc0, pc1, pc2;
	std::vector<BlockHdr> vec;
	std::vector<int32_t> scores;
	if (ctx->keep_dumps) { // stage 0 = the candidate blocks in the reference's -t 1 push order (group order, then qPos order), before RemoveOverlaps
		PieceTable pl0;
		GSA_TRY(queue_pieces(ctx, ws, ch, dc, dc + DC_N5, n, cq, cr, cl, cb, nullptr, 0, DC_NPL0, pl0));
		GSA_TRY(gsa_small_d2h(ctx, hc, dc, DC_COUNT * 4));
		GSA_TRY(gsa_small_d2h_counted(ctx, h_first, pl0.d, sizeof(Piece), dc + DC_NPL0, (int)std::min<int64_t>(PIECE_FIRST, n + 1)));
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		const int64_t n5 = hc[DC_N5];
		GSA_TRY(fetch_piece_table(ctx, pl0, hc[DC_NPL0], h_first, pc0));
		for (const Piece &p : pc0) ctx->blocks_stage[0].push_back(hdr_from_piece(p, (int32_t)p.sumlen)); // score = sum of seed lengths (AddAlnBlock :36)
		ctx->n_s0 = n5;
		GSA_TRY(gsa_ensure(ctx, ctx->d_s0q, (size_t)(n5 + 1) * 4)); GSA_TRY(gsa_ensure(ctx, ctx->d_s0r, (size_t)(n5 + 1) * 8)); GSA_TRY(gsa_ensure(ctx, ctx->d_s0l, (size_t)(n5 + 1) * 4));
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_s0q.p, cq, (size_t)n5 * 4, cudaMemcpyDeviceToDevice, ctx->stream));
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_s0r.p, cr, (size_t)n5 * 8, cudaMemcpyDeviceToDevice, ctx->stream));
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_s0l.p, cl, (size_t)n5 * 4, cudaMemcpyDeviceToDevice, ctx->stream));
	}

	// ---- 4. RemoveOverlaps: two passes queued blindly (survivors ping-pong between the context arrays and scratch); the rare
	// contig that still loses seeds in the second pass continues pass pair by pass pair below -----------------------------------------
	int32_t *tq = q3, *tl = l3, *tb = g3; int64_t *tr = r3;
	uint8_t *gapf = alive;
	int32_t *gcand = reach;
	PieceTable pt0, pt1, pt2;
	// GSA_BLOCK_LOGIC=host: the O(#blocks) logic on the host (two waits per contig) instead of in a kernel (one)
	const char *blenv = getenv("GSA_BLOCK_LOGIC");
	const bool dev_logic = !(blenv && strcmp(blenv, "host") == 0);
	for (int round = 0;; round++) {
		{ FOverlap f; f.q = cq; f.r = cr; f.l = cl; f.b = cb; f.dn = dc + (round == 0 ? DC_N5 : DC_N6); f.oq = tq; f.orr = tr; f.ol = tl; f.ob = tb; f.dc = dc; f.slot_n = DC_N6A;
		  if (round > 0) { // every earlier chain has run (the host waited): all chain states are free again
			  CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_chain.p, 0, ((size_t)K2_CHAINS * 2 + (size_t)K2_CHAINS * ch.tiles) * 8, ctx->stream));
			  ch.used = 0;
		  }
		  GSA_TRY(run_chain(ctx, ch, f, f.dn, n)); }
		{ FOverlap f; f.q = tq; f.r = tr; f.l = tl; f.b = tb; f.dn = dc + DC_N6A; f.oq = cq; f.orr = cr; f.ol = cl; f.ob = cb; f.dc = dc; f.slot_n = DC_N6; GSA_TRY(run_chain(ctx, ch, f, f.dn, n)); }
		// ---- 5. gap and contig-span break points, piece tables ----------------------------------------------------------------------
		Ws ws2 = ws; // the piece scratch of a repeated round reuses the same slots
		{ FGapFlags f; f.q = cq; f.r = cr; f.l = cl; f.b = cb; f.ce = (const ContigEnd *)ctx->d_cend.p; f.nce = (int)ctx->cend.size(); f.flag = gapf; f.cand = gcand; f.dc = dc;
		  GSA_TRY(run_chain(ctx, ch, f, dc + DC_N6, n)); }
		k_gap_similarity<<<1184, 32, 0, ctx->stream>>>(gcand, dc + DC_NCAND, cq, cr, cl, (const unsigned char *)ctx->d_seq.p, ctx->ix, gapf);
		KERNEL_CHECK(ctx);
		GSA_TRY(queue_pieces(ctx, ws2, ch, dc, dc + DC_N6, n, cq, cr, cl, cb, gapf, 0, DC_NP0, pt0)); // blocks keep their push order; only the ranges moved
		GSA_TRY(queue_pieces(ctx, ws2, ch, dc, dc + DC_N6, n, cq, cr, cl, cb, gapf, 1, DC_NP1, pt1));
		GSA_TRY(queue_pieces(ctx, ws2, ch, dc, dc + DC_N6, n, cq, cr, cl, cb, gapf, 3, DC_NP2, pt2));
		if (round == 0 && dev_logic) {
			// ---- 6-7 on the device: block logic and IdentifyNormalPairs queued blindly, ONE wait for the whole phase ------------------
			const int nctg = (int)ctx->contig_len.size();
			BlockHdr *d_vec = ws2.get<BlockHdr>(BLK_DEV_CAP), *d_st1 = nullptr, *d_st2 = nullptr;
			NpBlock *d_npb = ws2.get<NpBlock>(BLK_DEV_CAP);
			int64_t *d_fbeg = ws2.get<int64_t>(BLK_DEV_CAP + 1), *d_chr = ws2.get<int64_t>(nctg + 1);
			int32_t *d_res = ws2.get<int32_t>(BR_COUNT);
			if (ctx->keep_dumps) { d_st1 = ws2.get<BlockHdr>(BLK_DEV_CAP); d_st2 = ws2.get<BlockHdr>(BLK_DEV_CAP); }
			if (ws2.rc) return ws2.rc;
			GSA_TRY(gsa_ensure(ctx, ctx->d_frag, (size_t)(2 * n + 2) * sizeof(gsa_frag)));
			GSA_TRY(gsa_ensure(ctx, ctx->d_fblk, (size_t)(2 * n + 2) * 4));
			BlkLogicArgs A;
			static const int dev_cap = [] { const char *e = getenv("GSA_BLOCK_DEV_CAP"); int v = e ? atoi(e) : BLK_DEV_CAP; return v < 1 ? 1 : v > BLK_DEV_CAP ? BLK_DEV_CAP : v; }();
			A.cap = dev_cap;
			A.pt0 = pt0.d; A.pt1 = pt1.d; A.pt2 = pt2.d; A.kept_score = kept_score; A.dc = dc; A.vec = d_vec; A.npb = d_npb; A.stage1 = d_st1; A.stage2 = d_st2; A.res = d_res;
			A.P.ce = (const ContigEnd *)ctx->d_cend.p; A.P.nce = (int)ctx->cend.size(); A.P.genome = ctx->N; A.P.min_aln_len = P.min_aln_len;
			A.P.min_block_score = P.min_block_score; A.P.one_on_one = P.one_on_one; A.P.chr_score = d_chr; A.P.n_contigs = nctg;
			k_block_logic<<<1, 32, 0, ctx->stream>>>(A);
			KERNEL_CHECK(ctx);
			{ FNormalPairs f; f.nb = d_npb; f.nblk = 0; f.d_nblk = d_res + BR_NFINAL; f.q = cq; f.r = cr; f.l = cl; f.frag = (gsa_frag *)ctx->d_frag.p; f.fblk = (int32_t *)ctx->d_fblk.p;
			  f.blk_frag_beg = d_fbeg; f.dc = dc; GSA_TRY(run_chain(ctx, ch, f, dc + DC_N0, n)); }
			// what the host needs of it all: counters, outcome, the final block list, where each block's fragments start
			GSA_TRY(gsa_ensure_host(ctx, ctx->h_stage, 3 * (size_t)BLK_DEV_CAP * sizeof(BlockHdr) + (size_t)(BLK_DEV_CAP + 1) * 8 + 256));
			int32_t *h_res = (int32_t *)ctx->h_stage.p;
			BlockHdr *h_vec = (BlockHdr *)((char *)ctx->h_stage.p + 64), *h_st1 = h_vec + BLK_DEV_CAP, *h_st2 = h_st1 + BLK_DEV_CAP;
			int64_t *h_fb = (int64_t *)(h_st2 + BLK_DEV_CAP);
			GSA_TRY(gsa_small_d2h(ctx, hc, dc, DC_COUNT * 4));
			GSA_TRY(gsa_small_d2h(ctx, h_res, d_res, BR_COUNT * 4));
			GSA_TRY(gsa_small_d2h_counted(ctx, h_vec, d_vec, sizeof(BlockHdr), d_res + BR_NFINAL, BLK_DEV_CAP));
			GSA_TRY(gsa_small_d2h_counted(ctx, h_fb, d_fbeg, 8, d_res + BR_NFINAL, BLK_DEV_CAP));
			if (ctx->keep_dumps) {
				GSA_TRY(gsa_small_d2h_counted(ctx, h_st1, d_st1, sizeof(BlockHdr), d_res + BR_N1, BLK_DEV_CAP));
				GSA_TRY(gsa_small_d2h_counted(ctx, h_st2, d_st2, sizeof(BlockHdr), d_res + BR_N2, BLK_DEV_CAP));
			}
			CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); // ---- the one wait of the phase
			if (hc[DC_KILLS] == 0 && h_res[BR_STATUS] == 0) {
				ctx->n_cseeds = hc[DC_N6];
				ctx->split_hazard = h_res[BR_HAZARD];
				if (ctx->keep_dumps) { ctx->blocks_stage[1].assign(h_st1, h_st1 + h_res[BR_N1]); ctx->blocks_stage[2].assign(h_st2, h_st2 + h_res[BR_N2]); }
				const int nblk = h_res[BR_NFINAL];
				const int64_t nfr = hc[DC_NFR];
				ctx->final_blocks.assign(h_vec, h_vec + nblk);
				for (int k = 0; k < nblk; k++) {
					ctx->final_blocks[k].frag_beg = h_fb[k];
					ctx->final_blocks[k].n_frags = (int32_t)((k + 1 < nblk ? h_fb[k + 1] : nfr) - h_fb[k]);
				}
				ctx->n_frags = nblk ? nfr : 0;
				return GSA_OK;
			}
			// otherwise: the host path below, from the piece tables that are still where they were
		}
		const int64_t first = std::min<int64_t>(PIECE_FIRST, n + 1);
		GSA_TRY(gsa_small_d2h(ctx, hc, dc, DC_COUNT * 4));
		GSA_TRY(gsa_small_d2h_counted(ctx, h_first, pt0.d, sizeof(Piece), dc + DC_NP0, (int)first));
		GSA_TRY(gsa_small_d2h_counted(ctx, h_first + PIECE_FIRST, pt1.d, sizeof(Piece), dc + DC_NP1, (int)first));
		GSA_TRY(gsa_small_d2h_counted(ctx, h_first + 2 * PIECE_FIRST, pt2.d, sizeof(Piece), dc + DC_NP2, (int)first));
		if (round == 0) {
			GSA_TRY(gsa_ensure_host(ctx, ctx->h_stage, (size_t)first * 4));
			GSA_TRY(gsa_small_d2h_counted(ctx, ctx->h_stage.p, kept_score, 4, dc + DC_NB1, (int)first));
		}
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); // ---- first wait of the phase
		if (round == 0) {
			const int64_t nb1 = hc[DC_NB1];
			scores.resize((size_t)nb1);
			if (nb1 <= first) memcpy(scores.data(), ctx->h_stage.p, (size_t)nb1 * 4);
			else CUDA_TRY(ctx, cudaMemcpy(scores.data(), kept_score, (size_t)nb1 * 4, cudaMemcpyDeviceToHost));
		}
		if (hc[DC_KILLS] == 0 || round > 500) break;
	}
	const int64_t n6 = hc[DC_N6];
	ctx->n_cseeds = n6;
	if (hc[DC_N5] == 0 || n6 == 0) return GSA_OK;
	GSA_TRY(fetch_piece_table(ctx, pt0, hc[DC_NP0], h_first, pc0));
	if (pc0.size() != scores.size()) return gsa_fail(ctx, GSA_ERR_CUDA, "gsa_cluster: block count changed in RemoveOverlaps (%zu -> %zu)", scores.size(), pc0.size());
	vec.reserve(pc0.size());
	for (size_t i = 0; i < pc0.size(); i++) vec.push_back(hdr_from_piece(pc0[i], scores[i])); // blocks keep their pre-overlap score
	if (ctx->keep_dumps) ctx->blocks_stage[1] = vec;
	GSA_TRY(fetch_piece_table(ctx, pt1, hc[DC_NP1], h_first + PIECE_FIRST, pc1));
	GSA_TRY(fetch_piece_table(ctx, pt2, hc[DC_NP2], h_first + 2 * PIECE_FIRST, pc2));
	gsa_host_split(ctx, vec, pc1, pc2);
	if (ctx->keep_dumps) ctx->blocks_stage[2] = vec;

	// ---- 6. block-level dedup on the host (float ratios + std::sort ties, O(#blocks)) ------------------------------------------------------
	gsa_host_dedup(ctx, vec);

	// ---- 7. IdentifyNormalPairs for the surviving blocks -> fragment list --------------------------------------------------------------------
	int nblk = (int)vec.size();
	ctx->final_blocks = vec;
	if (nblk == 0) return GSA_OK;
	std::vector<NpBlock> npb((size_t)nblk);
	int64_t total = 0;
	for (int k = 0; k < nblk; k++) { npb[k].src_beg = vec[k].beg; npb[k].dst_beg = total; npb[k].n = (int32_t)(vec[k].end - vec[k].beg); npb[k].pad = 0; total += npb[k].n; }
	if (total >= 0x3FFFFFF0ll) return gsa_fail(ctx, GSA_ERR_LIMIT, "gsa_cluster: more than 2^30 seeds in the kept blocks of one contig");
	NpBlock *d_npb = ws.get<NpBlock>(nblk);
	int64_t *d_fbeg = ws.get<int64_t>(nblk + 1);
	if (ws.rc) return ws.rc;
	GSA_TRY(gsa_ensure(ctx, ctx->d_frag, (size_t)(2 * total + 2) * sizeof(gsa_frag)));
	GSA_TRY(gsa_ensure(ctx, ctx->d_fblk, (size_t)(2 * total + 2) * 4));
	GSA_TRY(gsa_ensure_host(ctx, ctx->h_stage, (size_t)nblk * (sizeof(NpBlock) + 8) + 64));
	memcpy(ctx->h_stage.p, npb.data(), (size_t)nblk * sizeof(NpBlock));
	GSA_TRY(gsa_small_h2d(ctx, d_npb, ctx->h_stage.p, (size_t)nblk * sizeof(NpBlock)));
	k_k2_init<<<1, 64, 0, ctx->stream>>>(dc, (int32_t)total); // dc[DC_N0] = element count of the last chain
	KERNEL_CHECK(ctx);
	{ FNormalPairs f; f.nb = d_npb; f.nblk = nblk; f.d_nblk = nullptr; f.q = cq; f.r = cr; f.l = cl; f.frag = (gsa_frag *)ctx->d_frag.p; f.fblk = (int32_t *)ctx->d_fblk.p; f.blk_frag_beg = d_fbeg; f.dc = dc;
	  GSA_TRY(run_chain(ctx, ch, f, dc + DC_N0, total)); }
	int64_t *h_fb = (int64_t *)((char *)ctx->h_stage.p + (size_t)nblk * sizeof(NpBlock) + 8);
	h_fb = (int64_t *)(((uintptr_t)h_fb + 7) & ~(uintptr_t)7);
	GSA_TRY(gsa_small_d2h(ctx, hc, dc, DC_COUNT * 4));
	GSA_TRY(gsa_small_d2h(ctx, h_fb, d_fbeg, (size_t)nblk * 8));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); // ---- second wait of the phase
	const int64_t nfr = hc[DC_NFR];
	for (int k = 0; k < nblk; k++) {
		ctx->final_blocks[k].frag_beg = h_fb[k];
		ctx->final_blocks[k].n_frags = (int32_t)((k + 1 < nblk ? h_fb[k + 1] : nfr) - h_fb[k]);
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
