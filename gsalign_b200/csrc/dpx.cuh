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

// Size classes of the packed-int16 wavefront kernels (the class picks the kernel, the lanes or warps per problem and how
// the rows are written):
//   P1 .. P32B  k_dpx_pack: LG = 1 / 2 / 4 / 8 / 16 / 32 lanes per problem (two query rows per lane, strips of 2*LG rows),
//               32/LG problems per warp, direction flags in shared memory, tracebacks of a CTA's problems walked side by
//               side.  The fragments between two seeds of a 1-2 % divergent pair are mostly 1..50 bases long:
//               a warp per problem would leave most lanes idle.
//   S1 / S2     k_dpx, one warp per problem, strips one after the other (n <= 256; m <= 240 / <= 1000), flags in HBM
//   G4..G16     k_dpx, one CTA per problem, up to 4 / 8 / 16 warps sweeping consecutive 64-row strips as a pipeline
// Pairs holding any letter outside ACGT (score 0 against everything, reference src/ksw2_alignment.cpp:258-262) run the
// same kernels in their HASN variant (per-lane score tables instead of the q XOR r table): class = size class + DPX_NSIZE.
enum { DPX_CLS_S1 = 0, DPX_CLS_S2 = 1, DPX_CLS_G4 = 2, DPX_CLS_G8 = 3, DPX_CLS_G16 = 4,
       DPX_CLS_P1 = 5, DPX_CLS_P2 = 6, DPX_CLS_P4 = 7, DPX_CLS_P8 = 8, DPX_CLS_P16 = 9, DPX_CLS_P32A = 10, DPX_CLS_P32B = 11,
       DPX_NSIZE = 12, DPX_NCLS = 24 };

// lanes per problem of a pack class
__host__ __device__ inline int dpx_pack_lanes(int size_cls)
{
	return size_cls == DPX_CLS_P1 ? 1 : size_cls == DPX_CLS_P2 ? 2 : size_cls == DPX_CLS_P4 ? 4 : size_cls == DPX_CLS_P8 ? 8 : size_cls == DPX_CLS_P16 ? 16 : 32;
}

// Shared-memory slot of one problem in k_dpx_pack, in 32-bit words; the same for every problem of a launch (sized by the
// largest m and n of the launch), odd so that the slots of a warp's problems start in different banks.
struct PackLayout {
	int Gx, nsx, cols;   // 8-step groups per strip, strips, entries of the boundary-row arrays (indexed from -2*LG)
	uint32_t off_hdr, off_bhe, off_a16, off_f0, off_f1, off_q, off_r, words;
};

__host__ __device__ inline PackLayout pack_layout(int LG, int Mx, int Nx)
{
	PackLayout L;
	const int R = 2 * LG;
	L.nsx = (Nx + R - 1) / R; if (L.nsx < 1) L.nsx = 1;
	L.Gx = (Mx + R - 1 + 7) >> 3;
	L.cols = R + 8 * L.Gx + 8;
	uint32_t o = 0;
	L.off_hdr = o; o += 4;                                           // pos, identical columns, m, n
	L.off_bhe = o; o += (uint32_t)L.cols;                            // {H, E'} of the row above the strip
	L.off_a16 = o; o += ((uint32_t)L.cols + 1u) >> 1;                // reference selector halves
	// the finished rows are assembled over bhe/a16 once the matrix is done: 2 x (m + n) characters
	uint32_t st = 2u * (((uint32_t)(Mx + Nx) + 3u) >> 2);
	if (o - L.off_bhe < st) o = L.off_bhe + st;
	L.off_f0 = o; o += (uint32_t)(LG * L.Gx * L.nsx);                // decision bits of the even rows of every lane
	L.off_f1 = o; o += (uint32_t)(LG * L.Gx * L.nsx);                // ... of the odd rows
	L.off_q = o; o += ((uint32_t)Nx + 3u) >> 2;
	L.off_r = o; o += ((uint32_t)Mx + 3u) >> 2;
	L.words = o | 1u;
	return L;
}

struct DpxLayout { // byte offsets into the shared memory slot of one problem in k_dpx
	int G;          // 8-step groups per strip
	int nstrips;    // 64-row strips
	uint32_t off_bhe, off_a16, off_prog, off_qch, off_rch, off_st, total;
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
	L.off_st = o; if (stage) o += 2u * (((uint32_t)(m + n) + 3u) & ~3u);
	L.total = (o + 15u) & ~15u;
	return L;
}

__host__ __device__ inline int dpx_class(int m, int n, bool has_other, bool use_pack = true)
{
	int size;
	if (use_pack && n <= 32 && m <= 128) {
		size = (n <= 2 && m <= 40) ? DPX_CLS_P1 : (n <= 4 && m <= 48) ? DPX_CLS_P2 : (n <= 8 && m <= 64) ? DPX_CLS_P4 : (n <= 16 && m <= 96) ? DPX_CLS_P8 : DPX_CLS_P16;
	} else size = (n <= 256 && m <= 240) ? DPX_CLS_S1 : (n <= 256 && m <= 1000) ? DPX_CLS_S2 : n <= 256 ? DPX_CLS_G4 : n <= 512 ? DPX_CLS_G8 : DPX_CLS_G16;
	return size + (has_other ? DPX_NSIZE : 0);
}

// bytes of the direction-flag pool (HBM) a problem needs: 4 bits per cell of every (64-row strip) x (8-step group) tile;
// the pack classes keep their flags in shared memory
__host__ __device__ inline int64_t dpx_flag_bytes(int m, int n, int cls)
{
	int size = cls >= DPX_NSIZE ? cls - DPX_NSIZE : cls;
	if (size >= DPX_CLS_P1) return 0;
	DpxLayout L = dpx_layout(m, n, false);
	return 256ll * L.G * L.nstrips;
}

// launches k_dpx over problems [0, nprob) that all belong to class cls and are no larger than max_m x max_n (dpx.cu).
// Rows are written right-aligned into each problem's slot [out_off, out_off + m + n) of the row pools; then either
// frag/bsum (pipeline) or out_len/out_start (batch) are filled.
int gsa_dpx_launch(gsa_ctx *ctx, cudaStream_t stream, int cls, int max_m, int max_n, const DpProblem *prob, int nprob, uint8_t *flags, char *a1, char *a2,
                   int32_t *out_len, int64_t *out_start, gsa_frag *frag, const int32_t *fblk, unsigned int *bsum);
int gsa_dpx_init_device(gsa_ctx *ctx);   // once per context: function attributes of the k_dpx variants (dpx.cu)
