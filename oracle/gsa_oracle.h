/* oracle/gsa_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's seed -> cluster/chain -> gapped-fill path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load liboracle.so; the product
 * (gsalign_b200/csrc) never links, includes or calls anything in this directory.
 *
 * Parity status: PINNED.  Every function below is checked by tests/test_oracle_vs_reference.py
 * against the unmodified reference compiled into oracle/_ref/libgsref.so (see ref_shim.cpp) on
 * E. coli (config C1) and on seeded synthetic genomes; golden digests of those runs are committed
 * under tests/golden/ so the pinning also holds where /root/reference is absent.
 */
#ifndef GSA_ORACLE_H
#define GSA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* In-memory view of a BWA-format index exactly as the reference loads it
 * (src/bwt_index.cpp:15-35,102-121; layout written by src/BWT_Index/bwtindex.c:53-75). */
typedef struct {
	const uint32_t *bwt;      /* .bwt payload after the 5 x u64 header: interleaved Occ blocks */
	uint64_t bwt_size;        /* in 32-bit words */
	uint64_t primary;
	uint64_t L2[5];
	uint64_t seq_len;         /* 2N */
	const uint64_t *sa;       /* n_sa samples, sa[0] = (uint64_t)-1 */
	uint64_t n_sa;
	int32_t sa_intv;
	const uint8_t *pac;       /* forward strand, 2 bit per base, MSB first */
	int64_t l_pac;            /* N */
	int32_t n_contigs;
	const int64_t *contig_off;
	const int32_t *contig_len;
} orc_index_t;

typedef struct {
	int32_t min_seed_len;     /* -slen, default 15 (10 with -sen)   src/main.cpp:210,323 */
	int32_t sensitive;        /* -sen                                src/main.cpp:272     */
	int32_t max_indel;        /* -ind, default 25                    src/main.cpp:214     */
	int32_t min_block_score;  /* -clr, default 200 (50 with -sen)    src/main.cpp:211,276 */
	int32_t min_aln_len;      /* -alen, default 200                  src/main.cpp:212     */
	int32_t min_idy;          /* -idy, default 70                    src/main.cpp:213     */
} orc_params_t;

/* traffic counters of the reference's algorithm (the B_seed yardstick) */
typedef struct {
	uint64_t n_search, n_ext_steps, n_split, n_sa_reads, n_lf_steps, n_seedhit, n_short, n_freqskip;
} orc_counters_t;

void orc_default_params(orc_params_t *p);

/* BWT_Search (src/bwt_search.cpp:141-185): loc must hold 100 entries */
void orc_bwt_search(const orc_index_t *idx, const char *seq, int32_t start, int32_t stop, int32_t min_seed_len,
                    int32_t *len, int32_t *freq, uint64_t *loc, orc_counters_t *ctr);

/* IdentifyLocalMEM at -t 1 (src/GSAlign.cpp:51-107): seeds sorted by (PosDiff, qPos).
 * Returns the seed count; *q,*r,*l are malloc'ed (free with orc_free). */
int64_t orc_seed_contig(const orc_index_t *idx, const orc_params_t *prm, const char *seq, int64_t seqlen,
                        int32_t **q, int64_t **r, int32_t **l, orc_counters_t *ctr);

/* SeedGrouping + GenerateAlignmentBlocks + CheckAlnBlockOverlaps + LargeGaps + SpanMultiSeqs
 * (src/GSAlign.cpp:126-391, src/ProcessCandidateAlignment.cpp:81-231, src/KmerAnalysis.cpp:78-121).
 * stage: 0 after SeedGroupAnalysis/AddAlnBlock, 1 after RemoveOverlaps, 2 after the gap and
 * contig-span splits.  Output: malloc'ed int64 stream
 *   [nblocks, { score, aln_len(0), bDup(0), nfrag, { bSeed, qPos, rPos, qLen, rLen } * nfrag } * nblocks]
 * Blocks are in the reference's -t 1 push order, except that RemoveBadAlnBlocks' std::sort tie order
 * is not reproduced (compare as multisets). Returns the stream length in int64 words. */
int64_t orc_cluster(const orc_index_t *idx, const orc_params_t *prm, const char *seq, int64_t seqlen,
                    int64_t nseeds, const int32_t *q, const int64_t *r, const int32_t *l,
                    int32_t stage, int64_t **out);

/* CalGapSimilarity (src/KmerAnalysis.cpp:78-121) */
int32_t orc_gap_similarity(const orc_index_t *idx, const char *seq, int32_t q1, int32_t q2, int64_t r1, int64_t r2);

/* IdentifyNormalPairs (src/ProcessCandidateAlignment.cpp:241-265) on one block.
 * in: nfrag x {bSeed,qPos,rPos,qLen,rLen}; out: malloc'ed, same record layout. Returns new nfrag. */
int64_t orc_normal_pairs(int64_t nfrag, const int64_t *frags, int64_t **out);

/* ksw2_alignment (src/ksw2_alignment.cpp:251-273): global affine DP, match 1 / mismatch -1 / N 0,
 * gap 2 + L, reference tie-breaks.  out1/out2 need m+n+1 bytes.  Returns the aligned length. */
int32_t orc_dp_align(const char *ref_frag, int32_t m, const char *qry_frag, int32_t n, char *out1, char *out2);

/* GenerateFragAlignment for one non-seed fragment (src/ProcessCandidateAlignment.cpp:308-342).
 * Writes the two rows (ref, query); returns aligned length, *score_inc = identical/“matching” columns
 * as the reference accumulates them. */
int32_t orc_frag_align(const orc_index_t *idx, const char *seq, int32_t qPos, int64_t rPos, int32_t qLen, int32_t rLen,
                       char *out1, char *out2, int32_t *score_inc, int32_t *used_dp);

/* one character of the text T = F . revcomp(F) the reference rebuilds (src/bwt_index.cpp:193-212) */
char orc_text_char(const orc_index_t *idx, int64_t pos);

void orc_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
