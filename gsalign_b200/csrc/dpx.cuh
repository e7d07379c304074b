// dpx.cuh -- problem record and size classes shared by the DP kernels of K3 (fill.cu, dpx.cu).
#pragma once
#include "gsa_internal.cuh"

#define DP_NEG (-30000)
#define DP_MAX_DIM 8000

struct DpProblem {
	const char *ref_chars; // explicit reference fragment (gsa_dp_batch) or nullptr -> read the 2-bit text at rpos
	const char *qry_chars;
	int64_t rpos;
	int64_t flag_off;      // byte offset into the direction-flag pool (multiple of 256)
	int64_t out_off;       // into the row pools: the problem owns [out_off, out_off + m + n)
	int32_t m, n;          // m = reference fragment length (columns), n = query fragment length (rows)
	int32_t frag;          // fragment index (pipeline) or pair index (batch)
	int32_t cls;           // DPX_CLS_*
};

// Size classes of the packed-int16 wavefront kernel k_dpx (the class picks the number of warps per problem and how the
// rows are written):
//   S1 / S2   one warp per problem, strips one after the other (n <= 256; m <= 240 / <= 1000: small shared memory)
//   G4..G16   one CTA per problem, up to 4 / 8 / 16 warps sweeping consecutive 64-row strips as a pipeline
// Pairs holding any letter outside ACGT (score 0 against everything, reference src/ksw2_alignment.cpp:258-262) run the
// same kernels in their HASN variant (per-lane score tables instead of the q XOR r table): class = size class + DPX_NSIZE.
enum { DPX_CLS_S1 = 0, DPX_CLS_S2 = 1, DPX_CLS_G4 = 2, DPX_CLS_G8 = 3, DPX_CLS_G16 = 4, DPX_NSIZE = 5, DPX_NCLS = 10 };

#define DPX_TBW 16   // traceback window, in 8-step groups (one 256-byte flag row each)

struct DpxLayout { // byte offsets into the shared memory slot of one problem in k_dpx
	int G;          // 8-step groups per strip
	int nstrips;    // 64-row strips
	uint32_t off_bhe, off_a16, off_prog, off_qch, off_rch, off_win, off_st, total;
};

// stage: the rows are assembled in shared memory and copied out coalesced (small problems)
__host__ __device__ inline DpxLayout dpx_layout(int m, int n, bool stage)
{
	DpxLayout L;
	L.nstrips = (n + 63) >> 6;
	int R = n < 64 ? n : 64;
	L.G = (m + R - 1 + 7) >> 3;
	uint32_t cols = 64u + 8u * (uint32_t)L.G + 8u, o = 0;
	L.off_bhe = o; o += 4u * cols;
	L.off_a16 = o; o += (2u * cols + 3u) & ~3u;
	L.off_prog = o; o += 4u * (uint32_t)L.nstrips;
	L.off_qch = o; o += ((uint32_t)n + 3u) & ~3u;
	L.off_rch = o; o += ((uint32_t)m + 3u) & ~3u;
	o = (o + 7u) & ~7u;
	L.off_win = o; o += 256u * (uint32_t)(L.G < DPX_TBW ? L.G : DPX_TBW);
	L.off_st = o; if (stage) o += 2u * (((uint32_t)(m + n) + 3u) & ~3u);
	L.total = (o + 15u) & ~15u;
	return L;
}

__host__ __device__ inline int dpx_class(int m, int n, bool has_other)
{
	int size = (n <= 256 && m <= 240) ? DPX_CLS_S1 : (n <= 256 && m <= 1000) ? DPX_CLS_S2 : n <= 256 ? DPX_CLS_G4 : n <= 512 ? DPX_CLS_G8 : DPX_CLS_G16;
	return size + (has_other ? DPX_NSIZE : 0);
}

// bytes of the direction-flag pool a problem needs: 4 bits per cell of every (64-row strip) x (8-step group) tile
__host__ __device__ inline int64_t dpx_flag_bytes(int m, int n)
{
	DpxLayout L = dpx_layout(m, n, false);
	return 256ll * L.G * L.nstrips;
}

// launches k_dpx over problems [0, nprob) that all belong to class cls and are no larger than max_m x max_n (dpx.cu).
// Rows are written right-aligned into each problem's slot [out_off, out_off + m + n) of the row pools; then either
// frag/bsum (pipeline) or out_len/out_start (batch) are filled.
int gsa_dpx_launch(gsa_ctx *ctx, cudaStream_t stream, int cls, int max_m, int max_n, const DpProblem *prob, int nprob, uint8_t *flags, char *a1, char *a2,
                   int32_t *out_len, int64_t *out_start, gsa_frag *frag, const int32_t *fblk, unsigned int *bsum);
int gsa_dpx_init_device(gsa_ctx *ctx);   // once per context: function attributes of the k_dpx variants (dpx.cu)
