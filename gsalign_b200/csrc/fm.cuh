// fm.cuh -- device-side primitives on the HBM index layout (see DevIndex in gsa_internal.cuh).
#pragma once
#include "gsa_internal.cuh"

// nst_nt4_table (reference src/BWT_Index/bntseq.c:40-57): A/a 0, C/c 1, G/g 2, T/t 3, else 4
__device__ __forceinline__ int gsa_nt4(unsigned char ch)
{
	unsigned char u = ch & 0xDF; // fold case (only letters reach here for valid input; others fall to 4)
	return u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : u == 'T' ? 3 : 4;
}

// number of symbols == c among the first m (1..64) symbols of a rank block's symbol quad
__device__ __forceinline__ uint32_t gsa_block_count(uint4 s, int c, int m)
{
	unsigned long long hi = ((unsigned long long)s.x << 32) | s.y, lo = ((unsigned long long)s.z << 32) | s.w;
	unsigned long long pat = 0x5555555555555555ull * (unsigned long long)c;
	hi ^= pat; lo ^= pat;                                   // matching symbols become 00
	hi = ~(hi | (hi >> 1)) & 0x5555555555555555ull;         // one bit per matching symbol
	lo = ~(lo | (lo >> 1)) & 0x5555555555555555ull;
	if (m <= 32) return __popcll(hi & (~0ull << (64 - 2 * m)));
	return __popcll(hi) + __popcll(lo & (~0ull << (128 - 2 * m)));
}

// ---- row width --------------------------------------------------------------------------------------------------
// W = false: rows, suffix-array values and text positions are u32 (|T| < 2^32); W = true: u64 (DevIndex::wide).
template <bool W> struct RowT { typedef uint32_t t; typedef uint2 ktab_t; };
template <> struct RowT<true> { typedef uint64_t t; typedef ulonglong2 ktab_t; };

// Occ(c, r): occurrences of c among the BWT characters of rows 0..r (inclusive), '$' row excluded.
// One 32-byte sector: two 128-bit loads from the same sector.  Per-base counts fit u32 in both widths.
template <bool W>
__device__ __forceinline__ uint32_t gsa_occ(const DevIndex &ix, int c, typename RowT<W>::t r)
{
	const uint4 *blk = ix.occ + 2 * (size_t)(r >> 6);
	uint4 cnt = __ldg(blk), sym = __ldg(blk + 1);
	uint32_t base = c == 0 ? cnt.x : c == 1 ? cnt.y : c == 2 ? cnt.z : cnt.w;
	return base + gsa_block_count(sym, c, (int)(r & 63) + 1) - (uint32_t)(c == 0 && r >= (typename RowT<W>::t)ix.primary);
}

// Occ(c, r1) and Occ(c, r2) for r1 <= r2; shares the block when both rows fall in the same one
template <bool W>
__device__ __forceinline__ void gsa_occ2(const DevIndex &ix, int c, typename RowT<W>::t r1, typename RowT<W>::t r2, uint32_t &o1, uint32_t &o2)
{
	typedef typename RowT<W>::t row_t;
	const row_t primary = (row_t)ix.primary;
	const uint4 *b1 = ix.occ + 2 * (size_t)(r1 >> 6);
	uint4 cnt = __ldg(b1), sym = __ldg(b1 + 1);
	uint32_t base = c == 0 ? cnt.x : c == 1 ? cnt.y : c == 2 ? cnt.z : cnt.w;
	o1 = base + gsa_block_count(sym, c, (int)(r1 & 63) + 1) - (uint32_t)(c == 0 && r1 >= primary);
	if ((r1 >> 6) != (r2 >> 6)) {
		const uint4 *b2 = ix.occ + 2 * (size_t)(r2 >> 6);
		cnt = __ldg(b2); sym = __ldg(b2 + 1);
		base = c == 0 ? cnt.x : c == 1 ? cnt.y : c == 2 ? cnt.z : cnt.w;
	}
	o2 = base + gsa_block_count(sym, c, (int)(r2 & 63) + 1) - (uint32_t)(c == 0 && r2 >= primary);
}

// BWT character of row r (0..3; the '$' row reads as 0 -- callers special-case primary)
template <typename I>
__device__ __forceinline__ int gsa_bwt_char(const DevIndex &ix, I r)
{
	const uint32_t *w = (const uint32_t *)(ix.occ + 2 * (size_t)(r >> 6) + 1);
	return (int)(__ldg(w + ((r & 63) >> 4)) >> ((~(uint32_t)r & 15) << 1)) & 3;
}

// SA[row]; the WIDE layout keeps 6 rows per 32-byte sector: {u32 lo[6]; u8 hi[6]; u8 pad[2]}
#define GSA_SA_GROUP 6
template <bool W>
__device__ __forceinline__ typename RowT<W>::t gsa_sa_read(const DevIndex &ix, typename RowT<W>::t row)
{
	if (!W) return __ldg((const uint32_t *)ix.sa + row);
	const uint64_t g = (uint64_t)row / GSA_SA_GROUP; const uint32_t k = (uint32_t)((uint64_t)row - g * GSA_SA_GROUP);
	const uint32_t *grp = (const uint32_t *)ix.sa + g * 8;
	uint32_t lo = __ldg(grp + k), hi = (__ldg(grp + 6 + (k >> 2)) >> ((k & 3) << 3)) & 0xFFu;
	return (typename RowT<W>::t)(((uint64_t)hi << 32) | lo);
}
template <bool W>
__device__ __forceinline__ void gsa_sa_write(void *sa, typename RowT<W>::t row, typename RowT<W>::t v)
{
	if (!W) { ((uint32_t *)sa)[row] = (uint32_t)v; return; }
	const uint64_t g = (uint64_t)row / GSA_SA_GROUP; const uint32_t k = (uint32_t)((uint64_t)row - g * GSA_SA_GROUP);
	uint32_t *grp = (uint32_t *)sa + g * 8;
	grp[k] = (uint32_t)v;
	((uint8_t *)(grp + 6))[k] = (uint8_t)((uint64_t)v >> 32);
}

// base i of a 2-bit MSB-first packed stream
template <typename I>
__device__ __forceinline__ int gsa_pk_base(const uint32_t *pk, I i)
{
	return (int)(__ldg(pk + (i >> 4)) >> ((~(uint32_t)i & 15) << 1)) & 3;
}

// 16 bases starting at base i (MSB first); the stream must be padded by one word
template <typename I>
__device__ __forceinline__ uint32_t gsa_pk_window(const uint32_t *pk, I i)
{
	uint32_t w0 = __ldg(pk + (i >> 4)), w1 = __ldg(pk + (i >> 4) + 1);
	return __funnelshift_l(w1, w0, ((uint32_t)i & 15) << 1);
}

// 32 flag bits starting at bit i of an MSB-first bitmap (padded by one word)
__device__ __forceinline__ uint32_t gsa_bit_window(const uint32_t *bm, uint32_t i)
{
	uint32_t w0 = __ldg(bm + (i >> 5)), w1 = __ldg(bm + (i >> 5) + 1);
	return __funnelshift_l(w1, w0, i & 31);
}

__device__ __forceinline__ char gsa_text_char(const DevIndex &ix, int64_t pos)
{
	return "ACGT"[gsa_pk_base(ix.txt, (uint64_t)pos)];
}

// ---- warp-aggregated atomics -----------------------------------------------------------------------------------
// Seeds and fragments arrive sorted, so the lanes of a warp mostly update the same counter (one giant block, one
// outlier window, one histogram bin); same-address atomics serialise, so peers are summed in registers first.
// peers = __match_any_sync(FULL, key) of the calling lane; ALL 32 lanes must call.  Returns the group total in the group's
// lowest lane (leader == true there); other lanes get a partial sum.
template <typename T>
__device__ __forceinline__ T gsa_peer_sum(unsigned peers, T v, bool &leader)
{
	const int lane = threadIdx.x & 31;
	if (peers == 0xffffffffu) { // the usual case (one block, one window, one bin per warp): a plain butterfly
		leader = lane == 0;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
		return v;
	}
	const unsigned rank = __popc(peers & ((1u << lane) - 1));
	leader = rank == 0;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		unsigned src = __fns(peers, lane, d + 1);                // the d-th peer above this lane, or 0xffffffff
		T t = __shfl_sync(0xffffffffu, v, src == 0xffffffffu ? lane : (int)src);
		if (src != 0xffffffffu && (rank & (2 * d - 1)) == 0) v += t;
	}
	return v;
}
