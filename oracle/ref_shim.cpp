// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" window onto the UNMODIFIED reference objects (compiled from /root/reference by
// oracle/Makefile into oracle/_ref/libgsref.so).  It lets tests drive the reference's own functions
// seam by seam -- BWT_Search (src/bwt_search.cpp:141), the seeding driver IdentifyLocalMEM
// (src/GSAlign.cpp:51), the cluster/chain phases in the order GenomeComparison runs them
// (src/GSAlign.cpp:490-526) and ksw2_alignment (src/ksw2_alignment.cpp:251) -- so that
// oracle/gsa_oracle.c (our restatement) and the CUDA product can be pinned against the real thing at
// every intermediate, not only on final MAF/VCF bytes.  Nothing here re-implements reference logic:
// it sets the reference's globals, calls its functions and serialises its containers.
#include "structure.h"

// non-static symbols of src/GSAlign.cpp that structure.h does not declare
extern vector<FragPair_t> SeedVec;
extern vector<pair<int, int> > SeedGroupVec;
extern uint32_t QrySeqPos, QryChrLength;
extern int SeedNum, SeedGroupNum, GroupID;
extern int64_t *RefChrScoreArr;
extern void *IdentifyLocalMEM(void *arg);
extern int SeedGrouping();
extern void *GenerateAlignmentBlocks(void *arg);
extern void EstChromosomeSimilarity();
extern void RemoveRedundantAlnBlocks(int type);

static int g_zero = 0;

static void serialise_blocks(vector<int64_t> &out)
{
	out.clear();
	out.push_back((int64_t)AlnBlockVec.size());
	for (size_t b = 0; b < AlnBlockVec.size(); b++) {
		const AlnBlock_t &B = AlnBlockVec[b];
		out.push_back(B.score); out.push_back(B.aln_len); out.push_back(B.bDup ? 1 : 0);
		out.push_back((int64_t)B.FragPairVec.size());
		for (size_t f = 0; f < B.FragPairVec.size(); f++) {
			const FragPair_t &F = B.FragPairVec[f];
			out.push_back(F.bSeed ? 1 : 0); out.push_back(F.qPos); out.push_back(F.rPos);
			out.push_back(F.qLen); out.push_back(F.rLen);
		}
	}
}

static vector<int64_t> g_stage[6];
static string g_aln; // all aln1/aln2 strings of the final stage, '\n'-separated, block/fragment order

extern "C" {

int ref_load_index(const char *prefix, int threads)
{
	iThreadNum = threads > 0 ? threads : 1;
	IndexFileName = strdup(prefix);
	RefIdx = bwa_idx_load(prefix);
	if (RefIdx == 0) return -1;
	Refbwt = RefIdx->bwt;
	FILE *saved = stderr; (void)saved;
	RestoreReferenceInfo();
	iThreadNum = 1;
	return 0;
}

void ref_set_params(int min_seed_len, int sensitive, int max_indel, int min_block_score, int min_aln_len, int min_idy, int one_on_one)
{
	MinSeedLength = min_seed_len; bSensitive = sensitive != 0; MaxIndelSize = max_indel;
	MinAlnBlockScore = min_block_score; MinAlnLength = min_aln_len; MinSeqIdy = min_idy;
	OneOnOneMode = one_on_one != 0; bAllowDuplication = true; bVCF = true; iThreadNum = 1;
}

int64_t ref_genome_size(void) { return GenomeSize; }

// copies RefSequence[pos, pos+len) (the 2N-char text T the reference rebuilds from .pac)
void ref_text(int64_t pos, int64_t len, char *out) { memcpy(out, RefSequence + pos, (size_t)len); }

// one BWT_Search call; loc must hold 100 entries
void ref_bwt_search(const char *seq, int seqlen, int start, int stop, int *len, int *freq, uint64_t *loc)
{
	string s(seq, (size_t)seqlen);
	bwtSearchResult_t r = BWT_Search(s, start, stop);
	*len = r.len; *freq = r.freq;
	for (int i = 0; i < r.freq; i++) loc[i] = r.LocArr[i];
	if (r.LocArr) delete[] r.LocArr;
}

// makes `seq` query contig 0 and runs the reference's seeding driver single-threaded
int64_t ref_seed_contig(const char *seq, int64_t len)
{
	QueryChrVec.clear(); QueryChrVec.resize(1);
	QueryChrVec[0].name = "q"; QueryChrVec[0].seq.assign(seq, (size_t)len);
	iQueryChrNum = 1; QueryChrIdx = 0; iThreadNum = 1;
	QrySeqPos = 0; QryChrLength = (uint32_t)len;
	SeedVec.clear(); SeedGroupVec.clear(); AlnBlockVec.clear();
	IdentifyLocalMEM(NULL);
	SeedNum = (int)SeedVec.size();
	return SeedNum;
}

void ref_get_seeds(int32_t *q, int64_t *r, int32_t *l)
{
	for (size_t i = 0; i < SeedVec.size(); i++) { q[i] = SeedVec[i].qPos; r[i] = SeedVec[i].rPos; l[i] = SeedVec[i].qLen; }
}

// runs the cluster/chain/fill phases on the current SeedVec exactly in GenomeComparison's order
// (src/GSAlign.cpp:495-540), snapshotting AlnBlockVec at the seams of SURVEY.md Appendix E:
//   stage 0: after GenerateAlignmentBlocks     1: after CheckAlnBlockOverlaps
//   stage 2: after LargeGaps + SpanMultiSeqs   3: after dedup + FillAlnBlockGaps
//   stage 4: after GenerateFragAlignment       5: after the identity filter + RemoveBadAlnBlocks
void ref_cluster(void)
{
	iThreadNum = 1; QueryChrIdx = 0;
	if (RefChrScoreArr) delete[] RefChrScoreArr;
	RefChrScoreArr = new int64_t[iChromsomeNum];
	SeedGroupVec.clear(); AlnBlockVec.clear();
	SeedNum = (int)SeedVec.size(); GroupID = 0; SeedGroupNum = SeedGrouping();
	GenerateAlignmentBlocks(&g_zero);
	serialise_blocks(g_stage[0]);
	AlnBlockNum = (int)AlnBlockVec.size(); CheckAlnBlockOverlaps(&g_zero);
	serialise_blocks(g_stage[1]);
	AlnBlockNum = (int)AlnBlockVec.size(); CheckAlnBlockLargeGaps(&g_zero); RemoveBadAlnBlocks();
	AlnBlockNum = (int)AlnBlockVec.size(); CheckAlnBlockSpanMultiSeqs(&g_zero); RemoveBadAlnBlocks();
	serialise_blocks(g_stage[2]);
	for (size_t i = 0; i < AlnBlockVec.size(); i++) AlnBlockVec[i].bDup = false;
	EstChromosomeSimilarity(); RemoveRedundantAlnBlocks(1); RemoveRedundantAlnBlocks(2);
	AlnBlockNum = (int)AlnBlockVec.size(); FillAlnBlockGaps(&g_zero);
	serialise_blocks(g_stage[3]);
	for (size_t i = 0; i < AlnBlockVec.size(); i++) AlnBlockVec[i].aln_len = AlnBlockVec[i].score = 0;
	GenerateFragAlignment(&g_zero);
	serialise_blocks(g_stage[4]);
	g_aln.clear();
	for (size_t b = 0; b < AlnBlockVec.size(); b++)
		for (size_t f = 0; f < AlnBlockVec[b].FragPairVec.size(); f++) {
			const FragPair_t &F = AlnBlockVec[b].FragPairVec[f];
			if (F.bSeed) continue;
			g_aln += F.aln1; g_aln += '\n'; g_aln += F.aln2; g_aln += '\n';
		}
	for (size_t i = 0; i < AlnBlockVec.size(); i++)
		if ((int)(100 * (1.0 * AlnBlockVec[i].score / AlnBlockVec[i].aln_len)) < MinSeqIdy) AlnBlockVec[i].score = 0;
	RemoveBadAlnBlocks();
	serialise_blocks(g_stage[5]);
}

int64_t ref_stage_size(int stage) { return (int64_t)g_stage[stage].size(); }
void ref_stage_copy(int stage, int64_t *out) { memcpy(out, g_stage[stage].data(), g_stage[stage].size() * sizeof(int64_t)); }
int64_t ref_aln_size(void) { return (int64_t)g_aln.size(); }
void ref_aln_copy(char *out) { memcpy(out, g_aln.data(), g_aln.size()); }

// ksw2_alignment(m, ref_frag, n, qry_frag): returns the gapped rows; out buffers must hold m+n+1 bytes
int ref_ksw2(const char *ref_frag, int m, const char *qry_frag, int n, char *out1, char *out2)
{
	string s1(ref_frag, (size_t)m), s2(qry_frag, (size_t)n);
	ksw2_alignment(m, s1, n, s2);
	memcpy(out1, s1.data(), s1.size()); out1[s1.size()] = 0;
	memcpy(out2, s2.data(), s2.size()); out2[s2.size()] = 0;
	return (int)s1.size();
}

// CalGapSimilarity on the current query contig (set by ref_seed_contig) and the loaded reference
int ref_gap_similarity(int q1, int q2, int64_t r1, int64_t r2) { return CalGapSimilarity(q1, q2, r1, r2) ? 1 : 0; }

} // extern "C"
