// scan.cuh -- single-pass chained prefix sum (decoupled look-back) whose input and output are functors.
//
// K2 is a long chain of "compute a flag or a count per seed, prefix-sum it, scatter by the prefix" steps (SeedGrouping,
// the compactions of SeedGroupAnalysis / RemoveOverlaps, block and piece tables, IdentifyNormalPairs).  Each such step is
// ONE launch of k_chain here: the functor computes the element's value from whatever arrays it likes (load / value), the
// kernel prefix-sums the values across the whole array in a single pass -- tiles are handed out by an atomic ticket, every
// tile publishes its aggregate, looks back over its predecessors' aggregates until it meets an inclusive prefix and
// publishes its own (warp-shuffle scans inside the tile) -- and the functor consumes (index, exclusive prefix, item)
// straight away (emit): scatter, segment id, hash insert ...  The element COUNT is read from device memory, so a chain of
// such steps needs no host round trip: every launch is sized by a host-side upper bound and tiles past the device count
// leave at once.
//
// Values are 62-bit sums; two counts below 2^31 can share one value (CH_PACK2) and are summed in one go.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define CH_THREADS 256
#define CH_ITEMS 4
#define CH_TILE (CH_THREADS * CH_ITEMS)
#define CH_FLAG_AGG (1ull << 62)
#define CH_FLAG_PFX (2ull << 62)
#define CH_VALUE_MASK ((1ull << 62) - 1)
#define CH_PACK2(lo, hi) ((unsigned long long)(uint32_t)(lo) | ((unsigned long long)(uint32_t)(hi) << 31))
#define CH_LO(v) ((uint32_t)((v) & 0x7FFFFFFFull))
#define CH_HI(v) ((uint32_t)(((v) >> 31) & 0x7FFFFFFFull))

// per call site: status[tile] (zeroed before the launch), ticket (zeroed), total (out: the grand total, may be null)
struct ChainState {
	unsigned long long *status;
	unsigned int *ticket;
	unsigned long long *total;
};

static inline int64_t chain_tiles(int64_t n_bound) { return (n_bound + CH_TILE - 1) / CH_TILE + 1; }

// F: struct Item; Item load(int64_t i) const; unsigned long long value(const Item &) const;
//    void emit(int64_t i, unsigned long long exclusive, const Item &, bool valid) const;   all __device__; emit is called by
//    every thread of the tile (valid = i < n) so that it may use warp collectives
//    void finish(unsigned long long total) const: called once, by one thread of the tile holding the last element
template <typename F>
__global__ void __launch_bounds__(CH_THREADS) k_chain(F f, const int32_t *d_n, ChainState cs)
{
	__shared__ unsigned int s_tile;
	__shared__ unsigned long long s_warp[CH_THREADS / 32], s_base;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) s_tile = atomicAdd(cs.ticket, 1u);
	__syncthreads();
	const unsigned int tile = s_tile;
	const int64_t n = *d_n, base = (int64_t)tile * CH_TILE;
	if (base >= n && !(n == 0 && tile == 0)) return; // nobody waits for a tile past the end
	typename F::Item it[CH_ITEMS] = {};
	unsigned long long v[CH_ITEMS], sum = 0;
	const int64_t i0 = base + (int64_t)tid * CH_ITEMS;
#pragma unroll
	for (int k = 0; k < CH_ITEMS; k++) {
		v[k] = 0;
		if (i0 + k < n) { it[k] = f.load(i0 + k); v[k] = f.value(it[k]); }
		sum += v[k];
	}
	// inclusive scan of the thread sums: warp shuffles, then the warp totals
	unsigned long long incl = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	unsigned long long wbase = 0, tile_sum = 0;
#pragma unroll
	for (int w = 0; w < CH_THREADS / 32; w++) { unsigned long long x = s_warp[w]; if (w < warp) wbase += x; tile_sum += x; }
	// chain: publish the aggregate, look back for the prefix of everything before this tile, publish the inclusive prefix
	if (warp == 0) {
		unsigned long long excl = 0;
		if (tile == 0) { if (lane == 0) atomicExch(cs.status, CH_FLAG_PFX | tile_sum); }
		else {
			if (lane == 0) atomicExch(cs.status + tile, CH_FLAG_AGG | tile_sum);
			int64_t look = (int64_t)tile - 1;
			for (;;) { // the 32 tiles below `look`, nearest first in lane 0
				unsigned long long st = 0;
				const int64_t t = look - lane;
				if (t >= 0) { do { st = *(volatile unsigned long long *)(cs.status + t); } while ((st >> 62) == 0); }
				const unsigned pfx = __ballot_sync(0xffffffffu, t >= 0 && (st >> 62) == 2);
				const int stop = pfx ? __ffs(pfx) - 1 : 31;   // nearest tile holding an inclusive prefix
				unsigned long long part = (t >= 0 && lane <= stop) ? (st & CH_VALUE_MASK) : 0;
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
				excl += part;
				if (pfx || look - 32 < 0) break;
				look -= 32;
			}
			if (lane == 0) atomicExch(cs.status + tile, CH_FLAG_PFX | ((excl + tile_sum) & CH_VALUE_MASK));
		}
		if (lane == 0) s_base = excl;
	}
	__syncthreads();
	unsigned long long run = s_base + wbase + (incl - sum);
#pragma unroll
	for (int k = 0; k < CH_ITEMS; k++) {
		f.emit(i0 + k, run, it[k], i0 + k < n); // every lane calls: emit may use warp collectives
		run += v[k];
	}
	// the tile holding the last element hands out the grand total
	if (tid == CH_THREADS - 1 && base + CH_TILE >= n) {
		if (cs.total) *cs.total = s_base + tile_sum;
		f.finish(s_base + tile_sum);
	}
}
