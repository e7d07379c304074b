/* gsalign_b200.h -- C ABI of the B200-native seed -> cluster/chain -> gapped-fill path.
 *
 * The reference (hsinnan75/GSAlign) has no plugin/FFI interface; its drop-in surface is the
 * `bin/GSAlign` executable and its files.  This header therefore mirrors the reference's INTERNAL
 * seams for the hot path -- one entry point per phase GenomeComparison() runs per query contig
 * (src/GSAlign.cpp:473-552) -- so that a maintainer can replace those phases call by call (see
 * INTEGRATION.md).  Plain C types only; no exceptions cross this boundary; every function returns
 * 0 on success and a negative gsa_status otherwise (gsa_last_error() has the text).
 *
 * Ownership: the caller owns every pointer it passes in (borrowed for the duration of the call).
 * The library owns device memory and the pinned host buffers behind every pointer it hands out;
 * those stay valid until the next gsa_contig_begin*() on the same context or gsa_destroy().
 * Threading: one gsa_ctx per GPU, driven by one host thread at a time.
 */
#ifndef GSALIGN_B200_H
#define GSALIGN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gsa_ctx gsa_ctx;

enum gsa_status {
	GSA_OK = 0,
	GSA_ERR_ARG = -1,      /* bad argument / call order */
	GSA_ERR_CUDA = -2,     /* CUDA runtime error (no CPU fallback exists) */
	GSA_ERR_LIMIT = -3,    /* input exceeds a documented limit of this build */
	GSA_ERR_NOMEM = -4
};

/* The BWA-format index exactly as the reference holds it after bwa_idx_load()
 * (bwt_t / bntseq_t, src/structure.h:28-66; loaders src/bwt_index.cpp:15-159). */
typedef struct {
	const uint32_t *bwt;       /* .bwt payload after the 5 x u64 header (Occ-interleaved words) */
	uint64_t bwt_size;         /* number of 32-bit words */
	uint64_t primary;          /* bwt_t::primary */
	uint64_t L2[5];            /* bwt_t::L2 */
	uint64_t seq_len;          /* bwt_t::seq_len = 2N */
	const uint64_t *sa;        /* bwt_t::sa, n_sa entries, sa[0] = (uint64_t)-1 */
	uint64_t n_sa;
	int32_t sa_intv;           /* bwt_t::sa_intv (32) */
	const uint8_t *pac;        /* bwaidx_t::pac: forward strand, 2 bit/base, MSB first */
	int64_t l_pac;             /* bntseq_t::l_pac = N */
	int32_t n_contigs;         /* bntseq_t::n_seqs */
	const int64_t *contig_off; /* bntann1_t::offset */
	const int32_t *contig_len; /* bntann1_t::len */
} gsa_index_view;

/* The globals the hot path reads (src/main.cpp:203-215,239-291). */
typedef struct {
	int32_t min_seed_len;      /* MinSeedLength     -slen  [15]; forced to 10 by -sen (src/main.cpp:323) */
	int32_t sensitive;         /* bSensitive        -sen   [0] */
	int32_t max_indel;         /* MaxIndelSize      -ind   [25] */
	int32_t min_block_score;   /* MinAlnBlockScore  -clr   [200] */
	int32_t min_aln_len;       /* MinAlnLength      -alen  [200] */
	int32_t min_idy;           /* MinSeqIdy         -idy   [70] */
	int32_t one_on_one;        /* OneOnOneMode      -one   [0] */
} gsa_params;

/* FragPair_t (src/structure.h:103-113) without the strings. */
typedef struct {
	int64_t rPos;
	int32_t qPos;
	int32_t qLen;
	int32_t rLen;
	int32_t bSeed;             /* 1 = exact seed, 0 = gap fragment between two seeds */
	int64_t aln_off;           /* gap fragments: offset of the two aligned rows in gsa_alignment::aln1/aln2 */
	int32_t aln_len;           /* gap fragments: number of alignment columns; seeds: qLen */
	int32_t reserved;          /* seeds: 0; gap fragments: how the rows were produced -- 1 deletion, 2 insertion, 3 copied (equal
	                              lengths, <= 5 mismatches), 4 gapped alignment (ksw2 semantics) */
} gsa_frag;

/* AlnBlock_t (src/structure.h:115-122); coor is left to the emitter (GenCoordinateInfo is host logic). */
typedef struct {
	int32_t score;             /* AlnBlock_t::score after GenerateFragAlignment (identical columns) */
	int32_t aln_len;           /* AlnBlock_t::aln_len */
	int32_t bDup;              /* AlnBlock_t::bDup */
	int32_t n_frags;
	int64_t frag_beg;          /* first fragment in gsa_alignment::frags */
} gsa_block;

/* Everything the emitters (OutputMAF / OutputAlignment / VariantIdentification) need for one query
 * contig: AlnBlockVec after src/GSAlign.cpp:540, in the reference's order. */
typedef struct {
	int32_t n_blocks;
	const gsa_block *blocks;
	int64_t n_frags;
	const gsa_frag *frags;
	int64_t aln_bytes;
	const char *aln1;          /* reference rows of all gap fragments ('-' = gap), not NUL-separated */
	const char *aln2;          /* query rows */
} gsa_alignment;

/* Per-phase device time of the last contig, milliseconds (CUDA events on the context's stream). */
typedef struct {
	float h2d_ms, seed_ms, cluster_ms, fill_ms, d2h_ms, host_ms;
	int64_t n_seeds, n_dp, dp_cells, n_frags;
	int64_t launches;          /* kernels launched by this library for the contig */
	float k_seed_ms;           /* the fm_seed kernel alone (K1's dominant launch) */
	float k_dp_ms;             /* the DP kernels alone (K3) */
	float total_ms;            /* first to last event of the contig on the context's stream */
} gsa_timing;

/* --- lifecycle ------------------------------------------------------------------------------- */
int gsa_create(int device, gsa_ctx **out);
/* A second lane on the owner's GPU: own stream, own scratch and result buffers, but the owner's uploaded index (read
 * only).  Query contigs are independent, so several lanes driven by one host thread each keep the GPU
 * busy across the host-side steps of a contig.  The owner must have its index uploaded and parameters set, and must
 * outlive every lane created from it. */
int gsa_create_shared(gsa_ctx *owner, gsa_ctx **out);
void gsa_destroy(gsa_ctx *ctx);
const char *gsa_last_error(const gsa_ctx *ctx);

/* Replaces bwa_idx_load() + RestoreReferenceInfo() for the device side (src/bwt_index.cpp:147-159,
 * 229-264): re-lays the index out in HBM (32-byte rank blocks, 2-bit text, full suffix array,
 * k-mer prefix table).  The view may be freed after the call returns. */
int gsa_index_upload(gsa_ctx *ctx, const gsa_index_view *view);
/* The device index has two row widths, like the reference's 64-bit bwtint_t (src/structure.h:28-38) allows: texts below
 * 2^32 symbols use 32-bit rows, larger ones (human-size genomes) 64-bit rows with a 40-bit suffix array.  enable = 1
 * forces the wide layout whatever the text size (parity tests run both); call before gsa_index_upload. */
int gsa_set_wide_index(gsa_ctx *ctx, int enable);
/* A replica of src's device index (derived structures included) on dst's GPU, copied GPU to GPU over NVLink instead of
 * being uploaded and re-derived once per GPU.  dst must be a context created with gsa_create on another device. */
int gsa_index_clone(gsa_ctx *dst, gsa_ctx *src);
/* bytes of HBM the index of this context occupies (rank blocks + text + suffix array + prefix table + presence bits) */
int64_t gsa_index_bytes(const gsa_ctx *ctx);
int gsa_set_params(gsa_ctx *ctx, const gsa_params *prm);
/* run on a caller-owned CUDA stream (a cudaStream_t passed as void*), e.g. to time with the caller's events */
int gsa_set_stream(gsa_ctx *ctx, void *cuda_stream);
void gsa_default_params(gsa_params *prm);

/* --- per query contig (the body of the loop at src/GSAlign.cpp:483-548) ----------------------- */
/* seq: the contig exactly as LoadQueryFile() stores it (any case, IUPAC allowed), host memory. */
int gsa_contig_begin(gsa_ctx *ctx, const char *seq, uint32_t len);
/* Double buffering: starts the upload of the contig that will be passed to the NEXT gsa_contig_begin / gsa_align_contig on
 * this context, on a copy stream of its own, so that it runs under the kernels of the contig being processed (the
 * reference's loop has nothing to overlap: its contigs are already in host memory).  seq must stay valid and unchanged
 * until that next call picks the buffer up (same pointer and length); pinned memory makes the copy asynchronous. */
int gsa_contig_prefetch(gsa_ctx *ctx, const char *seq, uint32_t len);
/* same, but seq is a DEVICE pointer (inputs already resident in HBM) */
int gsa_contig_begin_device(gsa_ctx *ctx, const void *dev_seq, uint32_t len);

/* K1  IdentifyLocalMEM + BWT_Search (src/GSAlign.cpp:51-107, src/bwt_search.cpp:141-185):
 * all seeds of the contig, sorted by (PosDiff, qPos).  n_seeds may be NULL. */
int gsa_seed(gsa_ctx *ctx, int64_t *n_seeds);

/* K2  SeedGrouping .. CheckAlnBlockSpanMultiSeqs, the block-level dedup and FillAlnBlockGaps
 * (src/GSAlign.cpp:495-514): candidate blocks, then the final fragment lists.  n_blocks may be NULL. */
int gsa_cluster(gsa_ctx *ctx, int32_t *n_blocks);

/* K3  GenerateFragAlignment + ksw2_alignment and the identity filter (src/GSAlign.cpp:523-540,
 * src/ProcessCandidateAlignment.cpp:290-351, src/ksw2_alignment.cpp:251-273): copies the result to
 * pinned host memory and fills *out. */
int gsa_fill(gsa_ctx *ctx, gsa_alignment *out);

/* enable = 0: gsa_fill() leaves fragments and rows in device memory (frags / aln1 / aln2 of its output stay NULL; the block
 * headers still come to the host) for consumers that read them there with gsa_result_device(), e.g. the record gather
 * of a multi-GPU job.  Default 1. */
int gsa_set_host_results(gsa_ctx *ctx, int enable);

/* The result of the last gsa_fill() where it was produced: same struct, but frags / aln1 / aln2 are DEVICE pointers
 * (blocks stays a host pointer, O(#blocks)).  Valid until the next gsa_contig_begin*() on this context.  This is what a
 * multi-GPU host packs into its outbox for the single record gather over NVLink. */
int gsa_result_device(gsa_ctx *ctx, gsa_alignment *out);

/* N3  VariantIdentification (src/SeqVariant.cpp:12-119) where the records are: one compact record per sequence variant of
 * the last gsa_fill() result, found by scanning the aligned rows on the device, in the order the reference pushes them
 * inside a block (fragment order, column order).  The alleles are not copied: they are substrings of the query contig
 * and of the reference text at the coordinates below, which the host owns (see gsa_variant_kind). */
enum gsa_variant_kind {
	GSA_VAR_SNV = 0,       /* REF = text[rPos], ALT = query[qPos]                                         (SeqVariant.cpp:57-64,97-105) */
	GSA_VAR_INS = 1,       /* insertion inside an aligned fragment: REF = query[qPos] (sic, hazard H6),
	                          ALT = query[qPos .. qPos+len]                                               (SeqVariant.cpp:69-81) */
	GSA_VAR_DEL = 2,       /* deletion inside an aligned fragment: REF = text[rPos .. rPos+len], ALT = text[rPos]   (:83-95) */
	GSA_VAR_FRAG_INS = 3,  /* fragment with rLen = 0: REF = text[rPos], ALT = query[qPos .. qPos+len]              (:43-54) */
	GSA_VAR_FRAG_DEL = 4   /* fragment with qLen = 0: REF = text[rPos .. rPos+len], ALT = query[qPos]              (:31-42) */
};
typedef struct {
	int64_t rPos;              /* text coordinate (T = F . revcomp(F)) of the variant's anchor base */
	int32_t qPos;              /* query coordinate of the anchor base */
	int32_t gPos;              /* GenCoordinateInfo(rPos).gPos: the VCF POS column (src/tools.cpp:120-140) */
	int32_t len;               /* SNV: 1; otherwise the number of inserted / deleted bases */
	int32_t kind;              /* gsa_variant_kind */
} gsa_variant;
typedef struct {
	int64_t n_variants;
	const gsa_variant *variants;   /* pinned host memory, all fragments of the contig in fragment order */
	const int64_t *block_first;    /* per block of the last gsa_fill() output (same order): first record of the block ... */
	const int64_t *block_count;    /* ... and how many (blocks with bDup set are listed too; the reference skips them) */
} gsa_variant_list;
/* call after gsa_fill() / gsa_align_contig() on the same context, before the next contig */
int gsa_variants(gsa_ctx *ctx, gsa_variant_list *out);

/* The three phases back to back on a host buffer: the call GenomeComparison() would make per contig. */
int gsa_align_contig(gsa_ctx *ctx, const char *seq, uint32_t len, gsa_alignment *out);

int gsa_get_timing(const gsa_ctx *ctx, gsa_timing *out);

/* Hazard detector: number of block-split phases of the last gsa_cluster() in which the block list grew
 * across a power of two.  There the reference splits through a reference into a std::vector it is pushing to
 * (src/ProcessCandidateAlignment.cpp:102-116,142-154): its result is undefined, so a difference against the reference on
 * such a contig is to be flagged, not counted as a mismatch.  0 = parity is defined. */
int gsa_split_hazard(const gsa_ctx *ctx);

/* keep (1) or drop (0, default) the intermediate block lists gsa_dump_blocks() serves */
int gsa_set_dump(gsa_ctx *ctx, int enable);

/* --- dump hooks for per-kernel parity tests (the seams of SURVEY.md appendix E) ----------------- */
/* seeds after gsa_seed(), sorted by (PosDiff, qPos); arrays must hold n_seeds entries */
int gsa_fetch_seeds(gsa_ctx *ctx, int32_t *qPos, int64_t *rPos, int32_t *len);
/* blocks after a stage of gsa_cluster(): 0 = SeedGroupAnalysis/AddAlnBlock, 1 = RemoveOverlaps,
 * 2 = gap + contig-span splits, 3 = dedup + FillAlnBlockGaps.  Stream layout:
 * [nblocks, {score, aln_len, bDup, nfrag, {bSeed,qPos,rPos,qLen,rLen} * nfrag} * nblocks].
 * Pass out = NULL to get the length in int64 words. */
int64_t gsa_dump_blocks(gsa_ctx *ctx, int32_t stage, int64_t *out);

/* Debug hook: checks n_samples pseudo-random rows of the uploaded device index for internal consistency (suffix order of
 * neighbouring rows by direct text comparison, BWT character = T[SA-1], SA[LF(row)] = SA[row]-1); *n_bad = violations.
 * Validates an index whose text is too large for the reference's own indexer to cross-check in test time. */
int gsa_index_selfcheck(gsa_ctx *ctx, int64_t n_samples, int64_t *n_bad);

/* --- multi-GPU: the record gather (reference: query contigs are independent and their records are only
 * appended in contig order, src/GSAlign.cpp:483-548) -----------------------------------------------------------------
 * Query contigs are dealt to GPUs, every GPU holds an index replica, and the finished records of all GPUs are collected
 * on the root GPU by ONE gather per job: grouped ncclSend / ncclRecv over NVLink.  NCCL is bound at run time (dlopen);
 * these calls fail with GSA_ERR_CUDA where it is absent, everything else works without it.
 * Outbox image (little-endian, every section padded to 16 bytes), one record per finished contig, either plain
 *   int64[4] {contig, n_blocks, n_frags, aln_bytes}, gsa_block[n_blocks], gsa_frag[n_frags], aln1[aln_bytes], aln2[aln_bytes]
 * or compact (the default; GSA_GATHER_RAW=1 keeps the plain form)
 *   int64[4] {-1 - contig, n_blocks, n_frags, aln_bytes}, int64[4] {n_anchors, 0, 0, 0}, gsa_block[n_blocks],
 *   uint64[n_frags] {qLen:21, rLen:21, aln_len:20, bSeed:1, gapped:1}, anchor[n_anchors] {int64 first_frag, rPos; int32 qPos, 0;
 *   int64 row_base}, aln1[aln_bytes], aln2[aln_bytes] */
/* one process per GPU: rank 0 makes the 128-byte id, the host hands it to every rank (MPI, torch.distributed, a file ...) */
int gsa_comm_unique_id(void *id, int32_t id_bytes);
int gsa_comm_init_rank(gsa_ctx *ctx, const void *id, int32_t rank, int32_t n_ranks);
/* one process driving n GPUs (bin/GSAlign -gpus n): ctxs[i] = the owner context of GPU i, rank i */
int gsa_comm_init_all(gsa_ctx *const *ctxs, int32_t n);
int gsa_comm_destroy(gsa_ctx *ctx);
/* the outbox of a GPU lives in its owner context; lanes append their last gsa_fill() result (device to device, on the
 * lane's stream, thread-safe) under the caller's contig index */
int gsa_outbox_reset(gsa_ctx *owner);
int gsa_outbox_reserve(gsa_ctx *owner, int64_t bytes);
int gsa_outbox_append(gsa_ctx *owner, gsa_ctx *lane, int64_t contig);
int64_t gsa_outbox_bytes(gsa_ctx *owner);
/* the gather: collective over the communicator (every rank calls it once per job) / the same for all ranks of a
 * single-process communicator.  Asynchronous on a communication stream; gsa_gather_wait blocks until it has run. */
int gsa_gather_records(gsa_ctx *owner, int32_t root);
int gsa_gather_records_all(gsa_ctx *const *ctxs, int32_t n, int32_t root);
int gsa_gather_wait(gsa_ctx *owner);
/* on the root: rank r's image where it arrived (device) / copied to pinned host memory (valid until the next call) */
int gsa_inbox_device(gsa_ctx *root, int32_t rank, const void **dev_ptr, int64_t *bytes);
int gsa_inbox_host(gsa_ctx *root, int32_t rank, const void **host_ptr, int64_t *bytes);
/* host-side walk over an image: returns 1 and fills *contig / *out (pointers into the image) for the record at *offset,
 * which it advances; 0 at the end; < 0 on a malformed image.  Records travel in a compact form by default (8 bytes per
 * fragment instead of 40: lengths only, positions and row offsets follow from an anchor every 256 fragments); for those
 * out->frags is NULL and gsa_record_frags() writes the fragment list where the caller wants it. */
int gsa_record_next(const void *image, int64_t bytes, int64_t *offset, int64_t *contig, gsa_alignment *out);
/* the n_frags fragment records of the record that STARTS at record_offset (the value *offset had before gsa_record_next
 * returned it), expanded or copied into dst; n_threads > 1 spreads the work over host threads */
int gsa_record_frags(const void *image, int64_t bytes, int64_t record_offset, gsa_frag *dst, int32_t n_threads);

/* Stand-alone batch of global alignments through K3's DP kernel (ksw2_alignment semantics):
 * pair i aligns ref[ref_off[i] .. ref_off[i+1]) with qry[qry_off[i] .. qry_off[i+1]); rows are written
 * to out1/out2 at out_off[i] = ref_off[i] + qry_off[i], lengths to out_len.  Host pointers. */
int gsa_dp_batch(gsa_ctx *ctx, int32_t n_pairs, const char *ref, const int64_t *ref_off, const char *qry,
                 const int64_t *qry_off, char *out1, char *out2, int32_t *out_len, float *kernel_ms);

/* Same, and out_identical[i] = number of identical columns of pair i as CountIdenticalPairs counts them
 * (src/ProcessCandidateAlignment.cpp:38-47: nst_nt4_table classes, so '-' equals any non-ACGT letter), i.e. what
 * gsa_fill adds to AlnBlock_t::score for a gap fragment. */
int gsa_dp_batch_identity(gsa_ctx *ctx, int32_t n_pairs, const char *ref, const int64_t *ref_off, const char *qry,
                          const int64_t *qry_off, char *out1, char *out2, int32_t *out_len, int32_t *out_identical);

/* On-box microbenchmark of the packed-int16 DPX issue rate, the roofline denominator of K3:
 * which 0 = VIADDMNMX.S16x2 (__viaddmax_s16x2), 1 = VIMNMX.S16x2, 2 = VIADD.16x2, 3 = VIMNMX3.S16x2.
 * Result in 1e9 thread-level instructions per second (each works on two int16 lanes). */
int gsa_dpx_peak(gsa_ctx *ctx, int which, double *ginstr_per_s);

#ifdef __cplusplus
}
#endif
#endif
