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
//   4. a streaming comparison of the 2-bit query against the 2-bit text (32 bases per step).
// One warp walks one chunk, one lane per 313-bp sub-chunk, speculatively and as a pipelined state machine
// (see seed_walk_pipe and k_seed below); all chunks of the contig are in flight together.
#include "fm.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <stdlib.h>
#include <algorithm>

#define SEED_PAD32 336   // 32-base words of padding behind the packed query and its bitmap (> SEED_IWORDS + 4, see k_seed)

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
	// padded: bases past qlen read as invalid; the padding covers the aligned bulk copy K1 stages the last chunk with
	uint32_t nw = (ctx->qlen >> 5) + 2 + SEED_PAD32;
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
	uint32_t qpk_words, qinv_words; // allocated words of qpk / qinv
	int min_seed_len, sensitive;
};

#define SEED_SPEC 0x40000000   // flag in the raw seed's len: found by a speculative walk, valid only from the lane's merge point on

__device__ __forceinline__ void emit_seed(const SeedOut &o, unsigned long long slot, int32_t q, int64_t r, int32_t len)
{
	if (slot < o.capacity) { o.q[slot] = q; o.r[slot] = r; o.len[slot] = len; }
}

// One search (BWT_Search semantics) from `start` inside a chunk ending at `stop`; the first K bases are known to be
// ACGT and inside the chunk.  Returns the match length; lo/size = row interval of revcomp(match); rpos = start of the
// match in T when it is unique.
template <bool W>
__device__ __forceinline__ typename RowT<W>::ktab_t ktab_read(const DevIndex &ix, uint32_t code)
{
	return __ldg((const typename RowT<W>::ktab_t *)ix.ktab + code);
}

template <bool W>
__device__ __forceinline__ int seed_search(const DevIndex &ix, const uint32_t *__restrict__ qpk, const uint32_t *__restrict__ qinv,
                                           uint32_t start, uint32_t stop, int K, typename RowT<W>::t &lo, typename RowT<W>::t &size, typename RowT<W>::t &rpos)
{
	typedef typename RowT<W>::t row_t;
	const row_t n = (row_t)ix.n;
	// 1. prefix table
	uint32_t code = gsa_pk_window(qpk, start) >> (32 - 2 * K);
	typename RowT<W>::ktab_t iv = ktab_read<W>(ix, code);
	lo = iv.x; size = iv.y;
	if (size == 0) return 0;
	const bool lo_is_sa = size == 1; // a k-mer occurring once: the table holds its suffix-array value, not its row
	uint32_t pos = start + K;
	// 2. backward search while the interval holds several rows
	while (size > 1 && pos < stop) {
		if ((__ldg(qinv + (pos >> 5)) >> (~pos & 31)) & 1) break;
		int c = 3 - gsa_pk_base(qpk, pos);
		uint32_t o1, o2;
		gsa_occ2<W>(ix, c, lo - 1, lo + size - 1, o1, o2);
		if (o2 == o1) break;
		lo = (row_t)ix.L2[c] + o1 + 1; size = o2 - o1; pos++;
	}
	// 3/4. unique: locate once, then compare the 2-bit query against the 2-bit text
	if (size == 1) {
		uint32_t m = pos - start;
		row_t p = n - (lo_is_sa ? lo : gsa_sa_read<W>(ix, lo)) - m;  // start of the match in T
		rpos = p;
		row_t tpos = p + m;
		while (pos < stop && tpos < n) {
			uint32_t x = gsa_pk_window(qpk, pos) ^ gsa_pk_window(ix.txt, tpos);
			uint32_t iw = gsa_bit_window(qinv, pos);
			int ext = min(min(__clz(x) >> 1, __clz(iw)), 16);   // first mismatch / first non-ACGT
			uint32_t lim = (uint32_t)min((row_t)(stop - pos), n - tpos);
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

// The repair walk: the TRUE chain of a chunk from the true entry `start` of a sub-chunk [base, limit), as a straight line
// (one lane runs it while the others wait, so it only has to be rare).  Emits its seeds unflagged and stops as soon as it
// lands on a start the speculative walk visited: from there on the speculative chain is the true chain.  Returns
// UINT_MAX then (the start in `merge`), else the first chain start >= limit.  Misses advance by one position, so a run
// of guaranteed misses (k-mer cut by a non-ACGT base or by the chunk end) is a run of consecutive starts.
template <bool W>
__device__ __forceinline__ uint32_t seed_walk_repair(const DevIndex &ix, const SeedArgs &A, uint32_t start, uint32_t base, uint32_t limit, uint32_t stop,
                                                     const uint32_t *vis, const SeedOut &out, uint32_t &merge)
{
	const int K = ix.ktab_k;
	auto visited = [&](uint32_t s) { return (vis[(s - base) >> 5] >> ((s - base) & 31)) & 1; };
	while (start < limit) {
		if (visited(start)) { merge = start; return 0xFFFFFFFFu; }
		if (start + K > stop) return limit; // fewer than K (<= MinSeedLength) bases left in the chunk: misses all the way
		int bad = __clz(gsa_bit_window(A.qinv, start)); // offset of the first non-ACGT base at or after start
		if (bad < K) { // every search starting in [start, start+bad] misses; any visited start inside the run merges the chains
			uint32_t ns = min(start + bad + 1, limit);
			for (uint32_t s2 = start + 1; s2 < ns; s2++) if (visited(s2)) { merge = s2; return 0xFFFFFFFFu; }
			start = ns;
			continue;
		}
		typename RowT<W>::t lo, size, rpos = 0;
		int len = seed_search<W>(ix, A.qpk, A.qinv, start, stop, K, lo, size, rpos);
		if (len >= A.min_seed_len && size <= GSA_MAX_SEED_FREQ) {
			unsigned long long slot = atomicAdd(out.count, (unsigned long long)size);
			if (size == 1) emit_seed(out, slot, (int32_t)start, (int64_t)rpos, len);
			else
				for (uint32_t i = 0; i < (uint32_t)size; i++)
					emit_seed(out, slot + i, (int32_t)start, (int64_t)(ix.n - (uint64_t)gsa_sa_read<W>(ix, lo + i) - (uint32_t)len), len);
			start += A.sensitive ? 5 : (uint32_t)len + 1;
		} else start++;
	}
	return start;
}

#ifdef SEED_PROFILE
__device__ unsigned long long g_seed_prof[16]; // trips A, trips B, active lanes A, active lanes B, cycles A, resolve, B, repairs, searches A/B, fails
#endif
// ---- the pipelined walk ---------------------------------------------------------------------------------------------
// The lanes of a warp walk different sub-chunks, so a straight-line search (table -> rank steps -> locate -> compare)
// leaves them in different loops and the warp executes one lane's dependent DRAM access at a time.  Here every lane is a
// small state machine instead: one trip of the warp-wide loop lets each lane consume the load it issued in the previous
// trip and issue the next one, so the random 32-byte accesses of all 32 lanes are in flight together and the warp pays
// one memory latency per trip, not one per lane.  The search semantics are those of seed_search / seed_walk above.
enum { ST_DONE = 0, ST_NEXT, ST_PRES, ST_KTAB, ST_BWD, ST_SA, ST_CMP, ST_BWD_ISSUE, ST_AFTER_BWD, ST_CMP_ISSUE };
#define SEED_LOOK 8   // starts whose presence bits are fetched together
#define SEED_PREROLL 24      // bases a speculative walk starts before its sub-chunk (default mode: chains merge at the next difference)
#define SEED_PREROLL_SEN 64  // sensitive mode steps by 5 after a hit, so chains take longer to fall in step

// the 16 bases at offset off (0..31) of the 48 held by three consecutive packed words
__device__ __forceinline__ uint32_t win48(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t off)
{
	return off < 16 ? __funnelshift_l(w1, w0, off << 1) : __funnelshift_l(w2, w1, (off - 16) << 1);
}

// The chunk's slice of the packed query and of the invalid-base bitmap, staged in shared memory by its warp: the walk
// reads them at every step, and in L1 they would compete with the random index sectors.
#define SEED_QWORDS 632   // (10000 + 112) / 16 packed words
#define SEED_IWORDS 320   // (16 + 10000 + 112) / 32 bitmap words, rounded up
struct ChunkQuery {
	const uint32_t *sq, *si;
	uint32_t qb, ib;        // first base held by sq / si (qb = chunk start, a multiple of 16; ib = chunk start rounded down to 32)
	__device__ __forceinline__ const uint32_t *words(uint32_t p) const { return sq + ((p - qb) >> 4); }
	__device__ __forceinline__ uint32_t window(uint32_t p) const { const uint32_t *w = words(p); return __funnelshift_l(w[1], w[0], (p & 15) << 1); }
	__device__ __forceinline__ int base(uint32_t p) const { return (int)(sq[(p - qb) >> 4] >> ((~p & 15) << 1)) & 3; }
	__device__ __forceinline__ uint32_t inv_window(uint32_t p) const { const uint32_t *w = si + ((p - ib) >> 5); return __funnelshift_l(w[1], w[0], p & 31); }
	__device__ __forceinline__ bool invalid(uint32_t p) const { return (si[(p - ib) >> 5] >> (~p & 31)) & 1; }
};

// Records every visited start of [base, limit) in vis and emits every seed it finds flagged SEED_SPEC.
template <bool W>
__device__ __forceinline__ uint32_t seed_walk_pipe(const DevIndex &ix, const SeedArgs &A, const ChunkQuery &Q, uint32_t start, uint32_t base, uint32_t limit, uint32_t stop,
                                                   uint32_t *vis, const SeedOut &out, bool active)
{
	typedef typename RowT<W>::t row_t;
	const int K = ix.ktab_k, KB = ix.kbits_k, KMIN = A.min_seed_len, lane = threadIdx.x & 31;
	const uint32_t *occw = (const uint32_t *)ix.occ;
	const row_t n = (row_t)ix.n, primary = (row_t)ix.primary;
	// the walk may start a little before `base` (pre-roll, see k_seed): starts below base are neither recorded nor emitted
	auto mark = [&](uint32_t a, uint32_t b) { a = max(a, base); if (a < b) vis_mark(vis, a - base, b - base); };
	uint32_t look = 0; // candidates of the presence lookahead in flight
	bool retry = false; // the previous search of this lane failed: misses come in runs (two differences closer than MinSeedLength)
	int st = (active && start < limit) ? ST_NEXT : ST_DONE;
	row_t lo = 0, size = 0, tpos = 0, rpos = 0, r1 = 0, r2 = 0;
	uint32_t pos = 0;
	int c = 0;
	// pending loads: one register set per kind of access.  Lanes in different states issue their loads from different
	// branches of the same trip; if two branches loaded into the same register the second would have to wait for the first
	// (write-after-write on the warp's scoreboard) and the branches' DRAM latencies would add up again.
	uint4 bw_s1 = make_uint4(0, 0, 0, 0), bw_s2 = bw_s1; uint32_t bw_c1 = 0, bw_c2 = 0; // rank step: symbols + count of both blocks
	row_t kt_lo = 0, kt_size = 0, sa_v = 0; uint32_t tx0 = 0, tx1 = 0, tx2 = 0;            // prefix table entry, SA entry, text words
	uint32_t pr[SEED_LOOK] = {0, 0, 0, 0, 0, 0, 0, 0};                                      // presence words
	while (__any_sync(0xffffffffu, st != ST_DONE)) {
		bool fin = false;
#ifdef SEED_PROFILE
		if (lane == 0) { atomicAdd(&g_seed_prof[0], 1ull); atomicAdd(&g_seed_prof[2], (unsigned long long)__popc(__ballot_sync(0xffffffffu, st != ST_DONE))); }
		else __ballot_sync(0xffffffffu, st != ST_DONE);
#endif
		// ---- consume the load issued in the previous trip ------------------------------------------------------------
		if (st == ST_PRES) { // presence bits of the KB-mers at start .. start+look-1: the first present one is searched, the rest miss
			const uint32_t *qw = Q.words(start);
			const uint32_t w0 = qw[0], w1 = qw[1], w2 = qw[2], o = start & 15;
			uint32_t f = look;
#pragma unroll
			for (int i = SEED_LOOK - 1; i >= 0; i--)
				if ((uint32_t)i < look && ((pr[i] >> ((win48(w0, w1, w2, o + i) >> (32 - 2 * KB)) & 31)) & 1)) f = i;
			mark(start, start + min(f + 1, look));
			start += f;
			if (f < look) {
				typename RowT<W>::ktab_t iv = ktab_read<W>(ix, win48(w0, w1, w2, o + f) >> (32 - 2 * K));
				kt_lo = iv.x; kt_size = iv.y;
				st = ST_KTAB;
			} else st = ST_NEXT;
		} else if (st == ST_KTAB) {
			lo = kt_lo; size = kt_size;
			if (size == 0) { pos = start; fin = true; }
			else if (size == 1) { // the table entry is the suffix-array value: straight to the text
				pos = start + K;
				rpos = n - lo - (uint32_t)K; tpos = rpos + (uint32_t)K;
				st = ST_CMP_ISSUE;
			} else { pos = start + K; st = ST_BWD_ISSUE; }
		} else if (st == ST_BWD) {
			uint32_t o1 = bw_c1 + gsa_block_count(bw_s1, c, (int)(r1 & 63) + 1) - (uint32_t)(c == 0 && r1 >= primary);
			uint32_t o2 = bw_c2 + gsa_block_count(bw_s2, c, (int)(r2 & 63) + 1) - (uint32_t)(c == 0 && r2 >= primary);
			if (o2 == o1) st = ST_AFTER_BWD;
			else { lo = (row_t)ix.L2[c] + o1 + 1; size = o2 - o1; pos++; st = ST_BWD_ISSUE; }
		} else if (st == ST_SA) {
			uint32_t m = pos - start;
			rpos = n - sa_v - m; tpos = rpos + m;
			st = ST_CMP_ISSUE;
		} else if (st == ST_CMP) { // 32 bases per trip: the text words were loaded last trip, the query sits in L1
			uint32_t sh = ((uint32_t)tpos & 15) << 1;
			uint32_t t0 = __funnelshift_l(tx1, tx0, sh), t1 = __funnelshift_l(tx2, tx1, sh);
			uint32_t x0 = Q.window(pos) ^ t0, x1 = Q.window(pos + 16) ^ t1;
			uint32_t ext = x0 ? (uint32_t)(__clz(x0) >> 1) : 16u + (uint32_t)(__clz(x1) >> 1);
			ext = min(ext, (uint32_t)__clz(Q.inv_window(pos)));         // first non-ACGT base
			uint32_t lim = (uint32_t)min((row_t)(stop - pos), n - tpos);
			if (ext >= lim) { pos += lim; fin = true; }
			else { pos += ext; tpos += ext; if (ext < 32) fin = true; else st = ST_CMP_ISSUE; }
		}
		// ---- decide the next access of a running search and issue it ------------------------------------------------------
		if (st == ST_BWD_ISSUE) {
			if (size > 1 && pos < stop && !Q.invalid(pos)) {
				c = 3 - Q.base(pos);
				r1 = lo - 1; r2 = lo + size - 1;
				const size_t b1 = (size_t)(r1 >> 6), b2 = (size_t)(r2 >> 6);
				bw_c1 = __ldg(occw + b1 * 8 + c); bw_s1 = __ldg(ix.occ + b1 * 2 + 1);
				bw_c2 = __ldg(occw + b2 * 8 + c); bw_s2 = __ldg(ix.occ + b2 * 2 + 1);
				st = ST_BWD;
			} else st = ST_AFTER_BWD;
		}
		if (st == ST_AFTER_BWD) {
			if (size == 1) { sa_v = gsa_sa_read<W>(ix, lo); st = ST_SA; }
			else fin = true;
		}
		if (st == ST_CMP_ISSUE) {
			if (pos < stop && tpos < n) {
				const uint32_t *tw = ix.txt + (tpos >> 4);
				tx0 = __ldg(tw); tx1 = __ldg(tw + 1); tx2 = __ldg(tw + 2);
				st = ST_CMP;
			} else fin = true;
		}
		// ---- a search ended: seeds (emitting pass), next start ------------------------------------------------------------
		uint32_t n_emit = 0; int len = 0; bool hit = false;
		if (fin) {
#ifdef SEED_PROFILE
			atomicAdd(&g_seed_prof[9], 1ull);
#endif
			len = (int)(pos - start);
			hit = len >= A.min_seed_len && size <= GSA_MAX_SEED_FREQ;
			if (hit && start >= base) n_emit = (uint32_t)size;
		}
		if (__any_sync(0xffffffffu, n_emit != 0)) { // one counter update per warp: exclusive scan of the lanes' seed counts
			uint32_t incl = n_emit;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
			unsigned long long wbase = 0;
			if (lane == 31) wbase = atomicAdd(out.count, (unsigned long long)incl);
			wbase = __shfl_sync(0xffffffffu, wbase, 31);
			if (n_emit) {
				unsigned long long slot = wbase + incl - n_emit;
				if (size == 1) emit_seed(out, slot, (int32_t)start, (int64_t)rpos, len | SEED_SPEC);
				else
					for (uint32_t i = 0; i < (uint32_t)size; i++)
						emit_seed(out, slot + i, (int32_t)start, (int64_t)(ix.n - (uint64_t)gsa_sa_read<W>(ix, lo + i) - (uint32_t)len), len | SEED_SPEC);
			}
		}
		if (fin) {
			start += hit ? (A.sensitive ? 5u : (uint32_t)len + 1u) : 1u;
			retry = !hit;
			st = ST_NEXT;
		}
		// ---- find the next search of this lane: guaranteed misses are skipped without touching the index -----------------
		// (a search yields a seed only if it reaches MinSeedLength bases: impossible when the chunk ends or a non-ACGT base
		// comes before that, or when the KB-mer (KB <= MinSeedLength) at the start does not occur in T at all)
		if (st == ST_NEXT) {
			st = ST_DONE;
			while (start < limit) {
				if (start + KMIN > stop) { // fewer than MinSeedLength bases left in the chunk: misses all the way
					mark(start, limit);
					start = limit;
					break;
				}
				int bad = __clz(Q.inv_window(start)); // offset of the first non-ACGT base at or after start
				if (bad < KMIN) { // every search starting in [start, start+bad] misses
					uint32_t ns = min(start + bad + 1, limit);
					mark(start, ns);
					start = ns;
					continue;
				}
				if (!retry || !ix.kbits) { // straight to the prefix table
					mark(start, start + 1);
					typename RowT<W>::ktab_t iv = ktab_read<W>(ix, Q.window(start) >> (32 - 2 * K));
					kt_lo = iv.x; kt_size = iv.y;
					st = ST_KTAB;
					break;
				}
				// starts start .. start+look-1 have MinSeedLength clean bases inside the chunk: fetch their presence bits together
				look = min(min((uint32_t)SEED_LOOK, limit - start), min(stop - KMIN - start + 1, (uint32_t)(bad - KMIN + 1)));
				const uint32_t *qw = Q.words(start);
				const uint32_t w0 = qw[0], w1 = qw[1], w2 = qw[2], o = start & 15;
#pragma unroll
				for (int i = 0; i < SEED_LOOK; i++) if ((uint32_t)i < look) pr[i] = __ldg(ix.kbits + (win48(w0, w1, w2, o + i) >> (32 - 2 * KB + 5)));
				st = ST_PRES;
				break;
			}
		}
	}
	return start;
}

// One warp per 10 kb chunk, one lane per 313-bp sub-chunk.  The chain of search starts inside a chunk is serial
// (reference src/GSAlign.cpp:70-93), but chains started anywhere re-synchronise at the next mismatch, so:
//   walk    every lane walks its sub-chunk speculatively from the sub-chunk's first base, records the visited starts and
//           emits the seeds it finds, flagged speculative;
//   resolve lane by lane, the true entry point of sub-chunk j (= exit of the true chain from sub-chunk j-1) is looked
//           up in lane j's visited set; if it is not there the lane walks the true chain from the true entry (emitting
//           its seeds unflagged) until it merges with the speculative chain (rare: needs a spurious seed spanning the
//           entry).  From the merge point on the speculative chain IS the true chain, so the lane's speculative seeds
//           are valid from there on and void before: the merge point goes to merge_from[] and k_seed_keys drops the rest.
// Results are exactly the serial chain's; every search of the true chain is done once.
// 80 registers (64-bit rows; 70 with 32-bit rows), 6-7 CTAs per SM.  Squeezing it to 64 registers for 8 CTAs per SM (a third
// more warps) was measured 4 % SLOWER on a C4 contig (0.90 vs 0.86 ms): the kernel sits at the memory system's
// random-sector rate, not at a lack of warps.
template <bool W>
__global__ void __launch_bounds__(32 * SEED_WARPS)
k_seed(DevIndex ix, SeedArgs A, SeedOut out, uint32_t *merge_from)
{
	__shared__ uint32_t s_vis[SEED_WARPS][32][SEED_VIS_WORDS];
	// The chunk's packed query (2.5 KB) and invalid-base bitmap (1.3 KB) are read by every search step of the warp: they are
	// staged in shared memory by two bulk copies (cp.async.bulk, the TMA unit's linear mode) that complete on the warp's own
	// mbarrier -- one elected lane issues them, no lane spends load instructions on it.  A bulk copy wants 16-byte aligned
	// addresses and sizes: it starts at the aligned word below the chunk's first word (the chunk then sits 0..3 words into
	// the buffer) and the arrays are padded by K0 far enough for the last chunk's copy to stay inside them.
	__shared__ __align__(16) uint32_t s_q[SEED_WARPS][SEED_QWORDS + 4], s_i[SEED_WARPS][SEED_IWORDS + 4];
	__shared__ __align__(8) unsigned long long s_bar[SEED_WARPS];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	uint32_t chunk = blockIdx.x * SEED_WARPS + wib;
	if (chunk >= A.nchunks) return;
	uint32_t cs = chunk * GSA_SEED_CHUNK, stop = min(cs + GSA_SEED_CHUNK, A.qlen);
	const uint32_t wq = cs >> 4, wi = cs >> 5;
	ChunkQuery Q; Q.sq = s_q[wib] + (wq & 3u); Q.si = s_i[wib] + (wi & 3u); Q.qb = cs; Q.ib = cs & ~31u;
	{
		const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar[wib]);
		constexpr uint32_t QB = (SEED_QWORDS + 4) * 4, IB = (SEED_IWORDS + 4) * 4;
		static_assert(QB % 16 == 0 && IB % 16 == 0, "bulk copies move multiples of 16 bytes");
		if (lane == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(QB + IB) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			             :: "r"((uint32_t)__cvta_generic_to_shared(s_q[wib])), "l"(A.qpk + (wq & ~3u)), "r"(QB), "r"(bar) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			             :: "r"((uint32_t)__cvta_generic_to_shared(s_i[wib])), "l"(A.qinv + (wi & ~3u)), "r"(IB), "r"(bar) : "memory");
		}
		__syncwarp();
		uint32_t done;
		do {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
		} while (!done);
	}
	uint32_t base = min(cs + lane * SEED_SUB, stop), limit = min(base + SEED_SUB, stop);
	uint32_t *vis = s_vis[wib][lane];
#pragma unroll
	for (int w = 0; w < SEED_VIS_WORDS; w++) vis[w] = 0;
#ifdef SEED_PROFILE
	long long t0 = clock64();
#endif
	// pre-roll: start a little before the sub-chunk so that the walk has usually fallen in step with the true chain by
	// the time it reaches `base` (chains merge at the first start they share); saves most of the serial repair walks
	const uint32_t pre = min((uint32_t)(A.sensitive ? SEED_PREROLL_SEN : SEED_PREROLL), base - cs);
	uint32_t spec_exit = seed_walk_pipe<W>(ix, A, Q, base - pre, base, limit, stop, vis, out, base < limit);
	__syncwarp();
#ifdef SEED_PROFILE
	long long t1 = clock64();
#endif
	// resolve the true entry of every sub-chunk
	uint32_t entry = cs, merge = 0xFFFFFFFFu;
	for (int j = 0; j < 32; j++) {
		uint32_t ex = entry;
		if (lane == j && entry < limit) {
			if ((vis[(entry - base) >> 5] >> ((entry - base) & 31)) & 1) { merge = entry; ex = spec_exit; }
			else {
#ifdef SEED_PROFILE
				atomicAdd(&g_seed_prof[7], 1ull);
#endif
				ex = seed_walk_repair<W>(ix, A, entry, base, limit, stop, vis, out, merge);
				if (ex == 0xFFFFFFFFu) ex = spec_exit;
			}
		}
		entry = __shfl_sync(0xffffffffu, ex, j);
	}
	merge_from[chunk * 32 + lane] = merge;
#ifdef SEED_PROFILE
	if (lane == 0) { long long t2 = clock64(); atomicAdd(&g_seed_prof[4], (unsigned long long)(t1 - t0)); atomicAdd(&g_seed_prof[5], (unsigned long long)(t2 - t1)); atomicAdd(&g_seed_prof[8], 1ull); }
#endif
}

#ifdef SEED_PROFILE
extern "C" int gsa_seed_profile(unsigned long long *out16, int reset)
{
	cudaDeviceSynchronize();
	if (cudaMemcpyFromSymbol(out16, g_seed_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
	if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_seed_prof, z, sizeof(z)); }
	return 0;
}
#endif

// sort key: ((PosDiff + qlen) << qbits) | qPos -- a strict total order on seeds, identical to CompByPosDiff
// (reference src/ProcessCandidateAlignment.cpp:3-7).  0 < PosDiff + qlen <= |T| + qlen needs pdbits bits and the all-ones
// value of that field is never a real one, so the void seeds (below) sort behind every real seed.  When pdbits + qbits
// exceeds 64 (|T| + qlen >= 2^33 with a contig above 1 Gbp) the two fields are sorted in two stable passes instead:
// which = 1 emits the qPos key, which = 2 the PosDiff key.
// Raw seeds flagged SEED_SPEC count only from their sub-chunk's merge point on (see k_seed); the others get the largest
// key, sort to the end and are dropped.  *n_valid receives the number of real seeds.
__global__ void k_seed_keys(const int32_t *q, const int64_t *r, const int32_t *len, const uint32_t *merge_from, uint64_t *key, uint32_t *val, int64_t n,
                            unsigned long long *n_valid, uint32_t qlen, int qbits, int which)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	bool ok = false;
	if (i < n) {
		uint32_t qi = (uint32_t)q[i], chunk = qi / GSA_SEED_CHUNK, sub = (qi - chunk * GSA_SEED_CHUNK) / SEED_SUB;
		ok = !(len[i] & SEED_SPEC) || qi >= merge_from[chunk * 32 + sub];
		uint64_t pd = (uint64_t)(r[i] - q[i] + (int64_t)qlen);
		uint64_t k = which == 0 ? (pd << qbits) | (uint64_t)qi : which == 1 ? (uint64_t)qi : pd;
		key[i] = ok ? k : ~0ull;
		val[i] = (uint32_t)i;
	}
	if (!n_valid) return;
	unsigned m = __ballot_sync(0xffffffffu, ok);
	if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_valid, (unsigned long long)__popc(m));
}

// second pass of the two-pass sort: the PosDiff key of the seeds in their qPos order
__global__ void k_seed_keys2(const uint32_t *perm, const uint64_t *pdkey, uint64_t *key, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) key[i] = pdkey[perm[i]];
}

__global__ void k_seed_gather(const uint32_t *perm, const int32_t *q, const int64_t *r, const int32_t *l,
                              int32_t *oq, int64_t *orr, int32_t *ol, int64_t n)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t s = perm[i];
	oq[i] = q[s]; orr[i] = r[s]; ol[i] = l[s] & ~SEED_SPEC;
}

int gsa_impl_seed(gsa_ctx *ctx)
{
	int k = ctx->prm.min_seed_len < GSA_KTAB_MAX_K ? ctx->prm.min_seed_len : GSA_KTAB_MAX_K;
	GSA_TRY(gsa_impl_build_ktab(ctx, k));
	GSA_TRY(gsa_impl_build_kbits(ctx, ctx->prm.min_seed_len));
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
		GSA_TRY(gsa_ensure(ctx, ctx->d_tmp[7], (size_t)(nchunks + 1) * 32 * 4)); // merge point per sub-chunk
		CUDA_TRY(ctx, cudaMemsetAsync(d_count, 0, 16, ctx->stream));
		SeedOut so; so.q = (int32_t *)ctx->d_tmp[0].p; so.r = (int64_t *)ctx->d_tmp[1].p; so.len = (int32_t *)ctx->d_tmp[2].p;
		so.count = d_count; so.capacity = cap;
		if (nchunks > 0) {
			CUDA_TRY(ctx, cudaEventRecord(ctx->ev[8], ctx->stream));
			SeedArgs sa; sa.qpk = (const uint32_t *)ctx->d_qpk.p; sa.qinv = (const uint32_t *)ctx->d_qinv.p; sa.qlen = ctx->qlen; sa.nchunks = nchunks;
			sa.qpk_words = 2 * ((ctx->qlen >> 5) + 2 + SEED_PAD32); sa.qinv_words = (ctx->qlen >> 5) + 2 + SEED_PAD32;
			sa.min_seed_len = ctx->prm.min_seed_len; sa.sensitive = ctx->prm.sensitive;
			const dim3 grid(gsa_grid(nchunks, SEED_WARPS)), block(32 * SEED_WARPS);
			uint32_t *mf = (uint32_t *)ctx->d_tmp[7].p;
			if (ctx->ix.wide) k_seed<true><<<grid, block, 0, ctx->stream>>>(ctx->ix, sa, so, mf);
			else k_seed<false><<<grid, block, 0, ctx->stream>>>(ctx->ix, sa, so, mf);
			KERNEL_CHECK(ctx);
			CUDA_TRY(ctx, cudaEventRecord(ctx->ev[9], ctx->stream));
		}
		GSA_TRY(gsa_small_d2h(ctx, ctx->h_small.p, d_count, 8));
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		produced = *(unsigned long long *)ctx->h_small.p;
		if (produced <= cap) break;
		cap = produced + 1024; // the kernel kept counting: rerun once with room for everything
	}
	if (produced > cap) return gsa_fail(ctx, GSA_ERR_LIMIT, "gsa_seed: seed buffer overflow");
	// raw seeds (speculative ones included) -> keys -> sort; the void ones sort to the end and only the real ones are gathered
	int64_t nraw = (int64_t)produced, n = 0;
	if (ctx->qlen > 0) ctx->seed_density = std::max(ctx->seed_density, (double)produced / ctx->qlen);
	if (nraw > 0) {
		GSA_TRY(gsa_ensure(ctx, ctx->d_tmp[3], (size_t)nraw * 8)); // keys in
		GSA_TRY(gsa_ensure(ctx, ctx->d_tmp[4], (size_t)nraw * 8)); // keys out
		GSA_TRY(gsa_ensure(ctx, ctx->d_tmp[5], (size_t)nraw * 4)); // vals in
		GSA_TRY(gsa_ensure(ctx, ctx->d_tmp[6], (size_t)nraw * 4)); // vals out
		uint64_t *k_in = (uint64_t *)ctx->d_tmp[3].p, *k_out = (uint64_t *)ctx->d_tmp[4].p; uint32_t *v_in = (uint32_t *)ctx->d_tmp[5].p, *v_out = (uint32_t *)ctx->d_tmp[6].p;
		int qbits = 1, pdbits = 1;
		while ((1ull << qbits) < (uint64_t)ctx->qlen) qbits++;
		while ((1ull << pdbits) <= ctx->ix.n + ctx->qlen + 1) pdbits++;
		auto sort_pairs = [&](uint64_t *ki, uint64_t *ko, uint32_t *vi, uint32_t *vo, int bits) -> int {
			size_t tmp_bytes = 0;
			cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ki, ko, vi, vo, nraw, 0, bits, ctx->stream);
			GSA_TRY(gsa_ensure(ctx, ctx->d_cub, tmp_bytes));
			CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->d_cub.p, tmp_bytes, ki, ko, vi, vo, nraw, 0, bits, ctx->stream));
			ctx->tm.launches += 4; // cub radix sort passes (histogram / onesweep family)
			return GSA_OK;
		};
		const bool two_pass = pdbits + qbits > 64 || getenv("GSA_SEED_SORT_2PASS") != nullptr;
		k_seed_keys<<<gsa_grid(nraw, 256), 256, 0, ctx->stream>>>((int32_t *)ctx->d_tmp[0].p, (int64_t *)ctx->d_tmp[1].p, (int32_t *)ctx->d_tmp[2].p, (const uint32_t *)ctx->d_tmp[7].p,
		                                                         k_in, v_in, nraw, d_count + 1, ctx->qlen, qbits, two_pass ? 1 : 0);
		KERNEL_CHECK(ctx);
		GSA_TRY(gsa_small_d2h(ctx, ctx->h_small.p, d_count + 1, 8));
		if (!two_pass) GSA_TRY(sort_pairs(k_in, k_out, v_in, v_out, pdbits + qbits));
		else { // stable LSD: by qPos first, then by PosDiff
			GSA_TRY(sort_pairs(k_in, k_out, v_in, v_out, 32));
			GSA_TRY(gsa_ensure(ctx, ctx->d_tmp[8], (size_t)nraw * 8));
			uint64_t *pdk = (uint64_t *)ctx->d_tmp[8].p;
			k_seed_keys<<<gsa_grid(nraw, 256), 256, 0, ctx->stream>>>((int32_t *)ctx->d_tmp[0].p, (int64_t *)ctx->d_tmp[1].p, (int32_t *)ctx->d_tmp[2].p, (const uint32_t *)ctx->d_tmp[7].p,
			                                                         pdk, v_in, nraw, nullptr, ctx->qlen, qbits, 2);
			KERNEL_CHECK(ctx);
			k_seed_keys2<<<gsa_grid(nraw, 256), 256, 0, ctx->stream>>>(v_out, pdk, k_in, nraw);
			KERNEL_CHECK(ctx);
			GSA_TRY(sort_pairs(k_in, k_out, v_out, v_in, 64));
			CUDA_TRY(ctx, cudaMemcpyAsync(v_out, v_in, (size_t)nraw * 4, cudaMemcpyDeviceToDevice, ctx->stream));
		}
		CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
		n = (int64_t)*(unsigned long long *)ctx->h_small.p;
	}
	ctx->n_seeds = n;
	GSA_TRY(gsa_ensure(ctx, ctx->d_sq, (size_t)(n + 1) * 4));
	GSA_TRY(gsa_ensure(ctx, ctx->d_sr, (size_t)(n + 1) * 8));
	GSA_TRY(gsa_ensure(ctx, ctx->d_sl, (size_t)(n + 1) * 4));
	if (n > 0) {
		k_seed_gather<<<gsa_grid(n, 256), 256, 0, ctx->stream>>>((uint32_t *)ctx->d_tmp[6].p, (int32_t *)ctx->d_tmp[0].p, (int64_t *)ctx->d_tmp[1].p, (int32_t *)ctx->d_tmp[2].p,
		                                                        (int32_t *)ctx->d_sq.p, (int64_t *)ctx->d_sr.p, (int32_t *)ctx->d_sl.p, n);
		KERNEL_CHECK(ctx);
	}
	ctx->have_seeds = true;
	return GSA_OK;
}
