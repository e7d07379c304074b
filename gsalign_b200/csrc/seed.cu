// seed.cu -- K1: maximal-exact-match seeding of one query contig against the FM-index in HBM.
//
// Replaces IdentifyLocalMEM (reference src/GSAlign.cpp:51-107) and BWT_Search with its rank/locate
// helpers (src/bwt_search.cpp:45-185).  Semantics reproduced exactly:
//   * the contig is cut into 10 000-bp chunks; inside a chunk searches are serially dependent
//     (next start = start + len + 1 on a hit, +5 in sensitive mode, +1 on a miss); chunks are independent
//   * one search = the longest prefix of seq[start, chunk_end) made of ACGT that occurs anywhere in
//     T = F . revcomp(F); it yields seeds iff len >= MinSeedLength and it occurs <= 100 times
// How it is computed differs from the reference (which is legal because the result is a canonical
// function of T): the reference spends one dependent 64-byte rank-block read per matched base plus a
// ~31-step LF walk per occurrence.  Here a search is
//   1. one k-mer prefix-table read (k <= MinSeedLength, so a k-mer that is absent, cut by the chunk end
//      or containing a non-ACGT base is a guaranteed miss and costs no index access at all),
//   2. backward-search steps on 32-byte rank blocks only while the interval holds more than one row,
//   3. one read of the full suffix array, and
//   4. a streaming comparison of the 2-bit query against the 2-bit text (16 bases per word).
// One thread walks one chunk; all chunks of the contig are in flight together.
#include "fm.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>

// ---- K0: 2-bit packing of the query + invalid-base bitmap -----------------------------------------
// one thread per 32 bases: reads 32 chars (two 16-byte loads), writes two packed words + one bitmap word
__global__ void k_pack_query(const unsigned char *seq, uint32_t qlen, uint32_t *qpk, uint32_t *qinv, uint32_t nwords32)
{
	uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= nwords32) return;
	uint32_t base = w << 5, p0 = 0, p1 = 0, inv = 0;
	if (base + 32 <= qlen) {
		const uint4 *src = (const uint4 *)(seq + base);
		uint4 a = __ldg(src), b = __ldg(src + 1);
		uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
		for (int i = 0; i < 32; i++) {
			int c = gsa_nt4((unsigned char)(v[i >> 2] >> ((i & 3) << 3)));
			uint32_t s = (uint32_t)(c & 3) << ((15 - (i & 15)) << 1);
			if (i < 16) p0 |= s; else p1 |= s;
			inv |= (uint32_t)(c > 3) << (31 - i);
		}
	} else {
		for (int i = 0; i < 32; i++) {
			int c = base + i < qlen ? gsa_nt4(seq[base + i]) : 4;
			uint32_t s = (uint32_t)(c & 3) << ((15 - (i & 15)) << 1);
			if (i < 16) p0 |= s; else p1 |= s;
			inv |= (uint32_t)(c > 3) << (31 - i);
		}
	}
	qpk[2 * w] = p0; qpk[2 * w + 1] = p1; qinv[w] = inv;
}

int gsa_impl_pack_query(gsa_ctx *ctx)
{
	uint32_t nw = (ctx->qlen >> 5) + 2; // padded: bases past qlen read as invalid
	GSA_TRY(gsa_ensure(ctx, ctx->d_qpk, (size_t)nw * 8 + 16));
	GSA_TRY(gsa_ensure(ctx, ctx->d_qinv, (size_t)nw * 4 + 16));
	k_pack_query<<<gsa_grid(nw, 256), 256, 0, ctx->stream>>>((const unsigned char *)ctx->d_seq.p, ctx->qlen, (uint32_t *)ctx->d_qpk.p, (uint32_t *)ctx->d_qinv.p, nw);
	KERNEL_CHECK(ctx);
	return GSA_OK;
}

// ---- K1 ---------------------------------------------------------------------------------------------
struct SeedOut {
	int32_t *q; int64_t *r; int32_t *len;
	unsigned long long *count;   // seeds produced (keeps counting past capacity)
	unsigned long long capacity;
};

struct SeedArgs {
	const uint32_t *qpk, *qinv;
	uint32_t qlen, nchunks;
	int min_seed_len, sensitive;
};

__device__ __forceinline__ void emit_seed(const SeedOut &o, unsigned long long slot, int32_t q, int64_t r, int32_t len)
{
	if (slot < o.capacity) { o.q[slot] = q; o.r[slot] = r; o.len[slot] = len; }
}

// One search (BWT_Search semantics) from `start` inside a chunk ending at `stop`; the first K bases are known to be
// ACGT and inside the chunk.  Returns the match length; lo/size = row interval of revcomp(match); rpos = start of the
// match in T when it is unique.
__device__ __forceinline__ int seed_search(const DevIndex &ix, const uint32_t *__restrict__ qpk, const uint32_t *__restrict__ qinv,
                                           uint32_t start, uint32_t stop, int K, uint32_t &lo, uint32_t &size, uint32_t &rpos)
{
	// 1. prefix table
	uint32_t code = gsa_pk_window(qpk, start) >> (32 - 2 * K);
	uint2 iv = __ldg(ix.ktab + code);
	lo = iv.x; size = iv.y;
	if (size == 0) return 0;
	uint32_t pos = start + K;
	// 2. backward search while the interval holds several rows
	while (size > 1 && pos < stop) {
		if ((__ldg(qinv + (pos >> 5)) >> (~pos & 31)) & 1) break;
		int c = 3 - gsa_pk_base(qpk, pos);
		uint32_t o1, o2;
		gsa_occ2(ix, c, lo - 1, lo + size - 1, o1, o2);
		if (o2 == o1) break;
		lo = ix.L2[c] + o1 + 1; size = o2 - o1; pos++;
	}
	// 3/4. unique: locate once, then compare the 2-bit query against the 2-bit text
	if (size == 1) {
		uint32_t m = pos - start;
		uint32_t p = ix.n - __ldg(ix.sa + lo) - m;  // start of the match in T
		rpos = p;
		uint32_t tpos = p + m;
		while (pos < stop && tpos < ix.n) {
			uint32_t x = gsa_pk_window(qpk, pos) ^ gsa_pk_window(ix.txt, tpos);
			uint32_t iw = gsa_bit_window(qinv, pos);
			int ext = min(min(__clz(x) >> 1, __clz(iw)), 16);   // first mismatch / first non-ACGT
			uint32_t lim = min(stop - pos, ix.n - tpos);
			if ((uint32_t)ext >= lim) { pos += lim; break; }
			pos += ext; tpos += ext;
			if (ext < 16) break;
		}
	}
	return (int)(pos - start);
}

#define SEED_SUB 313            // ceil(10000 / 32): one lane per sub-chunk
#define SEED_VIS_WORDS 10       // 313 bits
#define SEED_WARPS 4

__device__ __forceinline__ void vis_mark(uint32_t *vis, uint32_t a, uint32_t b)
{ // set bits [a, b) (LSB-first inside a word)
	for (uint32_t w = a >> 5; w <= ((b - 1) >> 5) && a < b; w++) {
		uint32_t lo = max(a, w << 5) & 31, hi = min(b, (w + 1) << 5) - (w << 5); // bits [lo, hi) of word w
		vis[w] |= (hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1)) & ~((1u << lo) - 1);
	}
}

// Walks the chunk's search chain from `start` until it leaves [.., limit); returns the first chain start >= limit.
//   MODE 0: speculative pass -- records every visited start of [base, limit) in vis
//   MODE 1: repair pass      -- stops as soon as it lands on a start the speculative pass visited (returns UINT_MAX then)
//   MODE 2: emitting pass    -- writes the seeds
// Misses advance by one position, so a run of guaranteed misses (k-mer cut by a non-ACGT base or by the chunk end)
// is a run of consecutive visited starts.
template <int MODE>
__device__ __forceinline__ uint32_t seed_walk(const DevIndex &ix, const SeedArgs &A, uint32_t start, uint32_t base, uint32_t limit, uint32_t stop,
                                              uint32_t *vis, const SeedOut &out)
{
	const int K = ix.ktab_k;
	while (start < limit) {
		if (MODE == 1 && ((vis[(start - base) >> 5] >> ((start - base) & 31)) & 1)) return 0xFFFFFFFFu;
		if (start + K > stop) { // fewer than K (<= MinSeedLength) bases left in the chunk: misses all the way
			if (MODE == 0) vis_mark(vis, start - base, limit - base);
			return limit;
		}
		int bad = __clz(gsa_bit_window(A.qinv, start)); // offset of the first non-ACGT base at or after start
		if (bad < K) { // every search starting in [start, start+bad] misses
			uint32_t ns = min(start + bad + 1, limit);
			if (MODE == 0) vis_mark(vis, start - base, ns - base);
			if (MODE == 1) { // any visited start inside the run merges the chains
				for (uint32_t s2 = start + 1; s2 < ns; s2++) if ((vis[(s2 - base) >> 5] >> ((s2 - base) & 31)) & 1) return 0xFFFFFFFFu;
			}
			start = ns;
			continue;
		}
		if (MODE == 0) vis[(start - base) >> 5] |= 1u << ((start - base) & 31);
		uint32_t lo, size, rpos = 0;
		int len = seed_search(ix, A.qpk, A.qinv, start, stop, K, lo, size, rpos);
		if (len >= A.min_seed_len && size <= GSA_MAX_SEED_FREQ) {
			if (MODE == 2) {
				unsigned long long slot = atomicAdd(out.count, (unsigned long long)size);
				if (size == 1) emit_seed(out, slot, (int32_t)start, (int64_t)rpos, len);
				else
					for (uint32_t i = 0; i < size; i++)
						emit_seed(out, slot + i, (int32_t)start, (int64_t)(ix.n - __ldg(ix.sa + lo + i) - (uint32_t)len), len);
			}
			start += A.sensitive ? 5 : (uint32_t)len + 1;
		} else start++;
	}
	return start;
}

// One warp per 10 kb chunk, one lane per 313-bp sub-chunk.  The chain of search starts inside a chunk is serial
// (reference src/GSAlign.cpp:70-93), but chains started anywhere re-synchronise at the next mismatch, so:
//   pass A  every lane walks its sub-chunk speculatively from the sub-chunk's first base and records the visited starts;
//   resolve lane by lane, the true entry point of sub-chunk j (= exit of the true chain from sub-chunk j-1) is looked
//           up in lane j's visited set; if it is not there the lane repairs by walking from the true entry until it
//           merges with its speculative chain (rare: needs a spurious seed spanning the entry);
//   pass B  every lane re-walks from its true entry and emits.  Results are exactly the serial chain's.
__global__ void __launch_bounds__(32 * SEED_WARPS)
k_seed(DevIndex ix, SeedArgs A, SeedOut out)
{
	__shared__ uint32_t s_vis[SEED_WARPS][32][SEED_VIS_WORDS];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	uint32_t chunk = blockIdx.x * SEED_WARPS + wib;
	if (chunk >= A.nchunks) return;
	uint32_t cs = chunk * GSA_SEED_CHUNK, stop = min(cs + GSA_SEED_CHUNK, A.qlen);
	uint32_t base = min(cs + lane * SEED_SUB, stop), limit = min(base + SEED_SUB, stop);
	uint32_t *vis = s_vis[wib][lane];
#pragma unroll
	for (int w = 0; w < SEED_VIS_WORDS; w++) vis[w] = 0;
	uint32_t spec_exit = seed_walk<0>(ix, A, base, base, limit, stop, vis, out);
	__syncwarp();
	// resolve the true entry of every sub-chunk
	uint32_t entry = cs, my_entry = cs;
	for (int j = 0; j < 32; j++) {
		uint32_t ex = entry;
		if (lane == j) {
			my_entry = entry;
			if (entry < limit) {
				if ((vis[(entry - base) >> 5] >> ((entry - base) & 31)) & 1) ex = spec_exit;
				else { ex = seed_walk<1>(ix, A, entry, base, limit, stop, vis, out); if (ex == 0xFFFFFFFFu) ex = spec_exit; }
			}
		}
		entry = __shfl_sync(0xffffffffu, ex, j);
	}
	if (my_entry < limit) seed_walk<2>(ix, A, my_entry, base, limit, stop, vis, out);
}

// sort key: ((PosDiff + 2^31) << 31) | qPos -- a strict total order on seeds, identical to CompByPosDiff
// (reference src/ProcessCandidateAlignment.cpp:3-7).  PosDiff + 2^31 < 2^33 because |T| < 2^32 here.
__global__ void k_seed_keys(const int32_t *q, const int64_t *r, uint64_t *key, uint32_t *val, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint64_t pd = (uint64_t)(r[i] - q[i] + (1ll << 31));
	key[i] = (pd << 31) | (uint64_t)(uint32_t)q[i];
	val[i] = (uint32_t)i;
}

__global__ void k_seed_gather(const uint32_t *perm, const int32_t *q, const int64_t *r, const int32_t *l,
                              int32_t *oq, int64_t *orr, int32_t *ol, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t s = perm[i];
	oq[i] = q[s]; orr[i] = r[s]; ol[i] = l[s];
}

int gsa_impl_seed(gsa_ctx *ctx)
{
	int k = ctx->prm.min_seed_len < GSA_KTAB_MAX_K ? ctx->prm.min_seed_len : GSA_KTAB_MAX_K;
	GSA_TRY(gsa_impl_build_ktab(ctx, k));
	uint32_t nchunks = (ctx->qlen + GSA_SEED_CHUNK - 1) / GSA_SEED_CHUNK;
	GSA_TRY(gsa_ensure(ctx, ctx->d_counter, 1024));
	unsigned long long *d_count = (unsigned long long *)ctx->d_counter.p;
	// raw (unsorted) seeds go to scratch 0..2, sorted seeds to d_sq/d_sr/d_sl
	// capacity: 1 seed / 32 bp covers default mode on any divergence; a denser contig (sensitive mode, repeats) reruns
	// once and the density is remembered so that later contigs of the same run do not
	unsigned long long cap = (unsigned long long)ctx->qlen / 32 + (1u << 16);
	cap = std::max(cap, (unsigned long long)(1.3 * ctx->seed_density * ctx->qlen) + (1u << 16));
	unsigned long long produced = 0;
	for (int attempt = 0; attempt < 2; attempt++) {
		GSA_TRY(gsa_ensure(ctx, ctx->d_tmp[0], cap * 4));
		GSA_TRY(gsa_ensure(ctx, ctx->d_tmp[1], cap * 8));
		GSA_TRY(gsa_ensure(ctx, ctx->d_tmp[2], cap * 4));
		CUDA_TRY(ctx, cudaMemsetAsync(d_count, 0, 8, ctx->stream));
		SeedOut so; so.q = (int32_t *)ctx->d_tmp[0].p; so.r = (int64_t *)ctx->d_tmp[1].p; so.len = (int32_t *)ctx->d_tmp[2].p;
		so.count = d_count; so.capacity = cap;
		if (nchunks > 0) {
			CUDA_TRY(ctx, cudaEventRecord(ctx->ev[8], ctx->stream));
			SeedArgs sa; sa.qpk = (const uint32_t *)ctx->d_qpk.p; sa.qinv = (const uint32_t *)ctx->d_qinv.p; sa.qlen = ctx->qlen; sa.nchunks = nchunks;
			sa.min_seed_len = ctx->prm.min_seed_len; sa.sensitive = ctx->prm.sensitive;
			k_seed<<<gsa_grid(nchunks, SEED_WARPS), 32 * SEED_WARPS, 0, ctx->stream>>>(ctx->ix, sa, so);
			KERNEL_CHECK(ctx);
			CUDA_TRY(ctx, cudaEventRecord(ctx->ev[9], ctx->stream));
		}
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_small.p, d_count, 8, cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		produced = *(unsigned long long *)ctx->h_small.p;
		if (produced <= cap) break;
		cap = produced + 1024; // the kernel kept counting: rerun once with room for everything
	}
	if (produced > cap) return gsa_fail(ctx, GSA_ERR_LIMIT, "gsa_seed: seed buffer overflow");
	int64_t n = (int64_t)produced;
	ctx->n_seeds = n;
	if (ctx->qlen > 0) ctx->seed_density = std::max(ctx->seed_density, (double)produced / ctx->qlen);
	GSA_TRY(gsa_ensure(ctx, ctx->d_sq, (size_t)(n + 1) * 4));
	GSA_TRY(gsa_ensure(ctx, ctx->d_sr, (size_t)(n + 1) * 8));
	GSA_TRY(gsa_ensure(ctx, ctx->d_sl, (size_t)(n + 1) * 4));
	if (n > 0) {
		GSA_TRY(gsa_ensure(ctx, ctx->d_tmp[3], (size_t)n * 8)); // keys in
		GSA_TRY(gsa_ensure(ctx, ctx->d_tmp[4], (size_t)n * 8)); // keys out
		GSA_TRY(gsa_ensure(ctx, ctx->d_tmp[5], (size_t)n * 4)); // vals in
		GSA_TRY(gsa_ensure(ctx, ctx->d_tmp[6], (size_t)n * 4)); // vals out
		k_seed_keys<<<gsa_grid(n, 256), 256, 0, ctx->stream>>>((int32_t *)ctx->d_tmp[0].p, (int64_t *)ctx->d_tmp[1].p, (uint64_t *)ctx->d_tmp[3].p, (uint32_t *)ctx->d_tmp[5].p, n);
		KERNEL_CHECK(ctx);
		size_t tmp_bytes = 0;
		cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (uint64_t *)ctx->d_tmp[3].p, (uint64_t *)ctx->d_tmp[4].p, (uint32_t *)ctx->d_tmp[5].p, (uint32_t *)ctx->d_tmp[6].p, n, 0, 64, ctx->stream);
		GSA_TRY(gsa_ensure(ctx, ctx->d_cub, tmp_bytes));
		CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->d_cub.p, tmp_bytes, (uint64_t *)ctx->d_tmp[3].p, (uint64_t *)ctx->d_tmp[4].p, (uint32_t *)ctx->d_tmp[5].p, (uint32_t *)ctx->d_tmp[6].p, n, 0, 64, ctx->stream));
		ctx->tm.launches += 4; // cub radix sort passes (upsweep/scan/downsweep family)
		k_seed_gather<<<gsa_grid(n, 256), 256, 0, ctx->stream>>>((uint32_t *)ctx->d_tmp[6].p, (int32_t *)ctx->d_tmp[0].p, (int64_t *)ctx->d_tmp[1].p, (int32_t *)ctx->d_tmp[2].p,
		                                                        (int32_t *)ctx->d_sq.p, (int64_t *)ctx->d_sr.p, (int32_t *)ctx->d_sl.p, n);
		KERNEL_CHECK(ctx);
	}
	ctx->have_seeds = true;
	return GSA_OK;
}
