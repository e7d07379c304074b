// dpx.cuh -- problem record and size classes shared by the two DP kernels of K3 (fill.cu, dpx.cu).
#pragma once
#include "gsa_internal.cuh"

#define DP_NEG (-30000)
#define DP_MAX_DIM 8000

struct DpProblem {
	const char *ref_chars; // explicit reference fragment (gsa_dp_batch) or nullptr -> read the 2-bit text at rpos
	const char *qry_chars;
	int64_t rpos;
	int64_t flag_off;      // byte offset into the direction-flag pool (multiple of 256)
	int64_t out_off;       // into the row pools
	int32_t m, n;          // m = reference fragment length (columns), n = query fragment length (rows)
	int32_t frag;          // fragment index (pipeline) or pair index (batch)
	int32_t cls;           // DPX_CLS_*
};

// Size classes.  A fragment pair made of ACGT only goes to the packed-int16 wavefront kernel k_dpx; the class picks the
// number of warps per problem and where the direction flags live.  Pairs holding any other letter (score 0 against
// everything, reference src/ksw2_alignment.cpp:258-262) take the scalar kernel k_dp.
enum { DPX_CLS_S4 = 0, DPX_CLS_S12 = 1, DPX_CLS_S48 = 2, DPX_CLS_G4 = 3, DPX_CLS_G8 = 4, DPX_CLS_G16 = 5, DPX_CLS_SCALAR = 6 };

struct DpxLayout { // byte offsets into the dynamic shared memory of k_dpx
	int G;          // 8-step groups per strip
	int nstrips;    // 64-row strips
	uint32_t off_bhe, off_a16, off_prog, off_qch, off_rch, off_flags, off_st, total;
};

__host__ __device__ inline DpxLayout dpx_layout(int m, int n, bool smem_flags)
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
	L.off_flags = o; if (smem_flags) o += 256u * (uint32_t)L.G * (uint32_t)L.nstrips;
	L.off_st = o; if (smem_flags) o += 2u * (((uint32_t)(m + n) + 3u) & ~3u);
	L.total = o;
	return L;
}

// bytes of the global direction-flag pool a problem needs
__host__ __device__ inline int64_t dpx_flag_bytes(int m, int n, int cls)
{
	if (cls == DPX_CLS_SCALAR) { int w = m < n ? m : n; return (((int64_t)(m + n - 1) * w) + 255) & ~255ll; }
	if (cls < DPX_CLS_G4) return 0;
	DpxLayout L = dpx_layout(m, n, false);
	return 256ll * L.G * L.nstrips;
}

#define DPX_SMEM_S4 (4 * 1024)
#define DPX_SMEM_S12 (12 * 1024)
#define DPX_SMEM_S48 (48 * 1024)

__host__ __device__ inline int dpx_class(int m, int n, bool has_other)
{
	if (has_other) return DPX_CLS_SCALAR;
	uint32_t t = dpx_layout(m, n, true).total;
	if (t <= DPX_SMEM_S4) return DPX_CLS_S4;
	if (t <= DPX_SMEM_S12) return DPX_CLS_S12;
	if (t <= DPX_SMEM_S48) return DPX_CLS_S48;
	return n <= 256 ? DPX_CLS_G4 : n <= 512 ? DPX_CLS_G8 : DPX_CLS_G16; // one warp per 64-row strip, up to 16
}

// launches k_dpx over problems [0, nprob) that all belong to class cls and are no larger than max_m x max_n (dpx.cu)
int gsa_dpx_launch(gsa_ctx *ctx, cudaStream_t stream, int cls, int max_m, int max_n, const DpProblem *prob, int nprob, uint8_t *flags, char *a1, char *a2, int32_t *out_len,
                   gsa_frag *frag, const int32_t *fblk, unsigned int *bsum);
