// gsa_internal.cuh -- context, device buffers and the on-device index layout shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/gsalign_b200.h"

#define GSA_SEED_CHUNK 10000   // SeedExplorationChunk, reference src/GSAlign.cpp:5 (observable: MEMs are cut at chunk ends)
#define GSA_MAX_SEED_FREQ 100  // MaxSeedFreq, reference src/bwt_search.cpp:3
#define GSA_MAX_SEED_GAP 5000  // MaxSeedGap, reference src/structure.h:23
#define GSA_NSIDE 4
#define GSA_KTAB_MAX_K 14
#define GSA_KBITS_MAX_K 16     // presence bitmap depth: 4^16 bits = 512 MB      // k-mer prefix table depth (never above MinSeedLength, see seed.cu)

// ----------------------------------------------------------------------------------------------
// Device index.  rows 0..n of the sorted suffix matrix of T$ (T = F . revcomp(F), |T| = n = 2N).
// Two row widths (the reference's bwtint_t is 64 bit everywhere, src/structure.h:28-38): texts below 2^32 symbols use
// 32-bit rows and suffix-array values (NARROW), larger ones -- human-size genomes, n = 6.2e9 at config C4 -- the WIDE
// layout.  Kernels that walk rows are compiled for both (template <bool W>, fm.cuh); everything else only addresses the
// text and takes 64-bit positions.
//   occ   one 32-byte block per 64 rows: {u32 cnt[4]; u32 sym[4]}.  sym holds the BWT characters of the
//         block's rows, 2 bit each, MSB first (row 64b at bits 31..30 of sym[0]); the row whose BWT
//         character is '$' (primary) stores 0 and is corrected at query time.  cnt[c] = number of c among
//         rows [0, 64b) INCLUDING that placeholder (per-base counts stay below 2^32 in both widths: checked at upload).
//         One rank query = one 32-byte sector.
//   txt   T itself, 2 bit per base, MSB first in u32 words (16 bases per word), padded with 2 words.
//   sa    the FULL suffix array: sa[row] = start of the suffix.  NARROW: u32 per row.  WIDE: 32-byte groups of 6 rows,
//         {u32 lo[6]; u8 hi[6]; u8 pad[2]} (40-bit values, 5.33 B per row: 33 GB at n = 6.2e9) -- one sector per locate.
//   kbits one bit per k-mer, k = min(MinSeedLength, 16): set iff the k-mer occurs in T (T is its own reverse complement).
//         A search whose first k bases do not occur cannot yield a seed: zero index accesses for it.
//   ktab  for every k-mer w (k = ktab_k, code = bases big-endian; when size == 1 lo is SA[row] itself): the row interval {lo, size} of
//         revcomp(w); size 0 = w does not occur in T.  NARROW: uint2, WIDE: ulonglong2.
// ----------------------------------------------------------------------------------------------
struct DevIndex {
	const uint4 *occ;
	const uint32_t *txt;
	const void *sa;
	const void *ktab;
	const uint32_t *kbits; // presence bitmap of the kbits_k-mers of T (bit `code`): a clear bit = the search cannot reach kbits_k bases
	uint64_t L2[5];
	uint64_t primary;
	uint64_t n;        // 2N
	int ktab_k;
	int kbits_k;
	int wide;          // 1 = WIDE layout of sa / ktab, 64-bit rows
};

struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
};

struct HostBuf { // pinned
	void *p = nullptr;
	size_t cap = 0;
};

// one contig-end table entry of ChrLocMap (reference src/bwt_index.cpp:247-252); pad = the contig's length
struct ContigEnd { int64_t end; int32_t idx; int32_t pad; };

struct BlockHdr { // host-side view of one candidate alignment block (a range of the device seed array)
	int32_t score;
	int32_t bDup;
	int64_t beg, end;  // range in the post-overlap seed arrays
	int32_t qf, ql, lenl; // first qPos, last qPos, last len
	int64_t rf, rl;       // first rPos, last rPos
	int64_t frag_beg; int32_t n_frags; int32_t aln_len; // filled by the normal-pair / fill phases
};

// one maximal run of seeds between break points (device -> host, O(#pieces))
struct Piece {
	int64_t beg, end, sumlen, rf, rl;
	int32_t qf, ql, lenl, pad;
};

struct gsa_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t stream2 = nullptr; // side stream (forked from / joined to `stream` with ev_fork / ev_join)
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
	cudaStream_t side[GSA_NSIDE] = {};   // side[0] = stream2: the DP classes of one contig run next to each other (fill.cu)
	cudaEvent_t ev_side[GSA_NSIDE] = {};
	cudaEvent_t ev[12] = {};        // 0/1 h2d, 2/3 seed, 4/5 cluster, 6/7 fill, 8/9 k_seed, 10/11 the DP launches
	bool own_stream = true; bool dp_timed = false;
	std::string err;
	gsa_params prm;
	gsa_timing tm;

	// index
	bool have_index = false;
	int force_wide = 0;            // gsa_set_wide_index / GSA_FORCE_WIDE: use the WIDE layout whatever the text size (tests)
	bool shares_index = false;     // lane created by gsa_create_shared: occ/txt/sa (and possibly ktab) belong to the owner
	DevIndex ix;
	int64_t N = 0;                 // GenomeSize
	DevBuf d_occ, d_txt, d_sa, d_ktab, d_kbits, d_cend;
	std::vector<ContigEnd> cend;   // sorted by end (forward and reverse ends of every contig)
	std::vector<int64_t> contig_off; std::vector<int32_t> contig_len;

	// query contig
	uint32_t qlen = 0;
	bool have_contig = false, have_seeds = false, have_cluster = false;
	DevBuf d_seq, d_qpk, d_qinv;   // raw chars, 2-bit packed, invalid-base bitmap
	// gsa_contig_prefetch: contigs on their way into spare buffers while the current one is processed.  Two slots: a caller
	// that announces contig C before it starts on contig B (itself announced earlier) has two uploads pending for a moment.
	struct Prefetch { DevBuf buf; const char *src = nullptr; uint32_t len = 0; cudaEvent_t ev = nullptr; };
	Prefetch pf[2];
	cudaStream_t copy_stream = nullptr;
	const char *h_seq = nullptr;   // borrowed host pointer (may be null for device-resident contigs)

	// K1 output / K2 working set
	int64_t n_seeds = 0;
	double seed_density = 0;       // highest seeds per query bp seen so far (sizes the raw seed buffer)
	DevBuf d_counter;              // small block of device counters
	DevBuf d_sq, d_sr, d_sl;       // seeds: qPos (i32), rPos (i64), len (i32), sorted by (PosDiff,qPos) after gsa_seed
	DevBuf d_tmp[64];              // scratch arrays for K2/K3 (sized on demand)
	DevBuf d_cub;                  // cub temp storage
	DevBuf d_chain;                // chained-scan states of K2 (scan.cuh)
	HostBuf h_small;               // pinned scratch for counters / piece tables
	HostBuf h_stage;               // pinned staging for dumps

	// K2 state kept for dumps and K3
	int64_t n_cseeds = 0;          // seeds after RemoveOverlaps (device arrays cq/cr/cl)
	DevBuf d_cq, d_cr, d_cl, d_cb; // qPos, rPos, len, block id
	std::vector<BlockHdr> blocks_stage[3]; // host block lists at dump stages 0..2 (stage 3 = final_blocks)
	DevBuf d_s0q, d_s0r, d_s0l;    // stage-0 seed arrays (pre-overlap), kept only when dumps are enabled
	int64_t n_s0 = 0;
	bool keep_dumps = false;
	bool host_results = true;      // gsa_fill copies fragments and rows to pinned host memory
	std::vector<BlockHdr> final_blocks;   // after dedup, reference order
	int split_hazard = 0;                 // hazard H14: split phases of this contig whose pushes crossed a power of two (gsa_split_hazard)

	// fragments (after FillAlnBlockGaps) and K3 output
	int64_t n_frags = 0;
	int64_t aln_bytes = 0;         // bytes used in each row pool
	DevBuf d_frag;                 // gsa_frag[n_frags]
	DevBuf d_fblk;                 // block index per fragment
	DevBuf d_aln1, d_aln2;         // row pools
	DevBuf d_bsum;                 // per block {aln_len, score}
	HostBuf h_frag, h_aln1, h_aln2, h_blocks;
	std::vector<gsa_block> out_blocks;
	bool have_fill = false;        // d_frag / d_aln1 / d_aln2 hold the result of gsa_fill for the current contig

	// N3 (variants.cu): variant records of the last gsa_fill result
	DevBuf d_var; HostBuf h_var, h_vrange;

	// multi-GPU record gather (gather.cu): the outbox of this GPU (owner context) and, on the root, the arrived images
	void *nccl_comm = nullptr; int comm_rank = 0, comm_size = 0;
	cudaStream_t comm_stream = nullptr;
	cudaEvent_t ev_gather = nullptr;       // end of this rank's part of the last gather (the next step's appends order after it)
	cudaEvent_t ev_outbox = nullptr;       // per lane: its last append
	DevBuf d_outbox, d_sizes;
	DevBuf d_cfrag, d_anchor, d_packst;    // per lane: a record's compact fragment list on its way into the outbox
	int64_t outbox_used = 0;
	std::mutex outbox_mu;                  // lanes of one GPU append from their own host threads
	std::vector<cudaEvent_t> outbox_pending;
	std::vector<DevBuf> d_inbox; std::vector<int64_t> inbox_bytes;
	HostBuf h_inbox, h_rec;
};

int gsa_fail(gsa_ctx *ctx, int code, const char *fmt, ...);
int gsa_ensure(gsa_ctx *ctx, DevBuf &b, size_t bytes);
int gsa_ensure_host(gsa_ctx *ctx, HostBuf &b, size_t bytes);
// small transfers between device memory and PINNED host memory, asynchronous on ctx->stream, done by a few warps over the
// mapped host pointer instead of a copy engine (capi.cu); sizes above 256 KB fall back to cudaMemcpyAsync
int gsa_small_d2h(gsa_ctx *ctx, void *host_pinned, const void *dev, size_t bytes);
int gsa_small_h2d(gsa_ctx *ctx, void *dev, const void *host_pinned, size_t bytes);
// bulk transfer between pinned host memory and HBM, queued in pieces of a few MB (capi.cu)
int gsa_bulk_copy(gsa_ctx *ctx, void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t stream);
// the first min(*d_count, cap) elements of a device table whose length is only known on the device
int gsa_small_d2h_counted(gsa_ctx *ctx, void *host_pinned, const void *dev, size_t elem_bytes, const int32_t *d_count, int cap);

#define CUDA_TRY(ctx, call)                                                                          \
	do {                                                                                             \
		cudaError_t _e = (call);                                                                     \
		if (_e != cudaSuccess)                                                                       \
			return gsa_fail((ctx), GSA_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
	} while (0)

#define GSA_TRY(call)                  \
	do {                               \
		int _r = (call);               \
		if (_r != GSA_OK) return _r;   \
	} while (0)

#define KERNEL_CHECK(ctx) do { (ctx)->tm.launches++; CUDA_TRY((ctx), cudaGetLastError()); } while (0)

static inline unsigned gsa_grid(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// phase entry points implemented per translation unit
int gsa_impl_index_upload(gsa_ctx *ctx, const gsa_index_view *v);
int gsa_impl_index_clone(gsa_ctx *dst, gsa_ctx *src);
int gsa_impl_build_ktab(gsa_ctx *ctx, int k);
int gsa_impl_build_kbits(gsa_ctx *ctx, int k);
int gsa_impl_pack_query(gsa_ctx *ctx);
int gsa_impl_seed(gsa_ctx *ctx);
int gsa_impl_cluster(gsa_ctx *ctx);
// host block logic (block_logic.cpp): exact restatement of the reference's O(#blocks) serial phases
void gsa_host_split(gsa_ctx *ctx, std::vector<BlockHdr> &vec, const std::vector<Piece> &p1, const std::vector<Piece> &p2);
void gsa_host_dedup(const gsa_ctx *ctx, std::vector<BlockHdr> &vec);
void gsa_host_remove_bad(std::vector<BlockHdr> &vec);
int gsa_host_chr_idx(const gsa_ctx *ctx, int64_t rpos, int64_t *end_out);
int gsa_impl_fill(gsa_ctx *ctx, gsa_alignment *out);
int gsa_impl_variants(gsa_ctx *ctx, gsa_variant_list *out);
int gsa_impl_dp_batch(gsa_ctx *ctx, int32_t n_pairs, const char *ref, const int64_t *ref_off, const char *qry,
                      const int64_t *qry_off, char *out1, char *out2, int32_t *out_len, int32_t *out_identical, float *kernel_ms);
