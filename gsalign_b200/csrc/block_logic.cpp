// block_logic.cpp -- the O(#blocks) serial phases of GenomeComparison, host form.
//
// The reference runs these serially on a handful of block headers; they use float/double comparisons and libstdc++'s
// unstable std::sort whose tie order is observable in the MAF record order (SURVEY.md hazards H9, H14, appendix D).  To be
// byte-exact this file runs the SAME std::sort calls on the same element order with the same comparators:
//   RemoveBadAlnBlocks                  reference src/ProcessCandidateAlignment.cpp:72-79
//   CheckGapsBetweenSeeds (tail)        reference src/ProcessCandidateAlignment.cpp:140-155
//   CheckAlnBlockSpanMultipleRefChrs    reference src/ProcessCandidateAlignment.cpp:100-117
//   EstChromosomeSimilarity             reference src/GSAlign.cpp:393-407
//   RemoveRedundantAlnBlocks            reference src/GSAlign.cpp:415-471
// All O(#seeds) work (break-point detection, piece sums) was done on the device; this file only sees block headers and
// piece tables.  By default the same logic runs in a kernel (block_logic.cuh + stdsort.cuh, cluster.cu: k_block_logic) so
// that K2 needs one host wait per contig; this form is the path of contigs the kernel declines (more than 1 024 blocks,
// another RemoveOverlaps round), of GSA_BLOCK_LOGIC=host, and the checker of the array form in tests/test_boundary_cpu.py.
#include <algorithm>
#include "gsa_internal.cuh"

static bool by_score_desc(const BlockHdr &a, const BlockHdr &b) { return a.score > b.score; }                      // CompByAlnBlockScore
static bool by_query_pos(const BlockHdr &a, const BlockHdr &b) { return a.qf == b.qf ? a.score > b.score : a.qf < b.qf; } // CompByAlnBlockQueryPos, src/GSAlign.cpp:17-21
static bool by_ref_pos(const BlockHdr &a, const BlockHdr &b) { return a.rf == b.rf ? a.score > b.score : a.rf < b.rf; }   // CompByAlnBlockRefPos, src/GSAlign.cpp:23-27

void gsa_host_remove_bad(std::vector<BlockHdr> &vec)
{
	size_t num = vec.size();
	std::sort(vec.begin(), vec.end(), by_score_desc);
	while (num > 0 && vec[num - 1].score == 0) num--;
	vec.resize(num);
}

int gsa_host_chr_idx(const gsa_ctx *ctx, int64_t rpos, int64_t *end_out)
{ // ChrLocMap.lower_bound(rpos): first contig end >= rpos
	size_t lo = 0, hi = ctx->cend.size();
	while (lo < hi) { size_t m = (lo + hi) / 2; if (ctx->cend[m].end < rpos) lo = m + 1; else hi = m; }
	if (lo == ctx->cend.size()) lo = ctx->cend.size() - 1; // cannot happen for rpos < 2N
	if (end_out) *end_out = ctx->cend[lo].end;
	return ctx->cend[lo].idx;
}

static BlockHdr from_piece(const Piece &p, int32_t score)
{
	BlockHdr b; b.score = score; b.bDup = 0; b.beg = p.beg; b.end = p.end; b.qf = p.qf; b.ql = p.ql; b.lenl = p.lenl;
	b.rf = p.rf; b.rl = p.rl; b.frag_beg = 0; b.n_frags = 0; b.aln_len = 0;
	return b;
}

// Hazard H14: the reference splits a block through a reference into AlnBlockVec that it keeps across
// AlnBlockVec.push_back (src/ProcessCandidateAlignment.cpp:102-116,142-154).  When a push crosses the vector's capacity
// (libstdc++ doubles it) that reference dangles and what the reference then computes is undefined; parity is only defined
// where that does not happen.  The block count straddling a power of two during a split phase is the observable trigger:
// such contigs are counted so that a harness can flag a difference there instead of calling it a failure.
static bool crosses_power_of_two(size_t before, size_t after)
{
	if (after <= before) return false;
	size_t p = 1;
	while (p < before) p <<= 1;      // the smallest capacity that held `before` elements
	return before == 0 || after > p;
}

static void split_phase(gsa_ctx *ctx, std::vector<BlockHdr> &vec, const std::vector<Piece> &pieces)
{
	size_t n0 = vec.size();
	for (size_t i = 0; i < n0; i++) {
		int64_t beg = vec[i].beg, end = vec[i].end;
		// pieces are sorted by beg and nest inside blocks: first piece with beg >= block.beg
		size_t lo = 0, hi = pieces.size();
		while (lo < hi) { size_t m = (lo + hi) / 2; if (pieces[m].beg < beg) lo = m + 1; else hi = m; }
		size_t k = lo, cnt = 0;
		while (k + cnt < pieces.size() && pieces[k + cnt].beg < end) cnt++;
		if (cnt <= 1) continue; // no break point inside: the block keeps its score
		vec[i].score = 0;
		for (size_t t = 0; t < cnt; t++) {
			const Piece &p = pieces[k + t];
			// CalAlnBlockScore, src/ProcessCandidateAlignment.cpp:26-36
			int32_t sc = (p.ql + p.lenl - p.qf) < ctx->prm.min_aln_len ? 0 : (int32_t)p.sumlen;
			if (sc > ctx->prm.min_block_score) vec.push_back(from_piece(p, sc));
		}
	}
	if (crosses_power_of_two(n0, vec.size())) ctx->split_hazard++;
	gsa_host_remove_bad(vec);
}

void gsa_host_split(gsa_ctx *ctx, std::vector<BlockHdr> &vec, const std::vector<Piece> &p1, const std::vector<Piece> &p2)
{
	split_phase(ctx, vec, p1); // CheckAlnBlockLargeGaps + RemoveBadAlnBlocks, src/GSAlign.cpp:504-505
	split_phase(ctx, vec, p2); // CheckAlnBlockSpanMultiSeqs + RemoveBadAlnBlocks, src/GSAlign.cpp:507-508
}

static inline bool dup_chr_score(int64_t s1, int64_t s2)
{ // CheckDuplicatedChrScore(int,int), src/GSAlign.cpp:409-413 (arguments are truncated to int there)
	int a = (int)s1, b = (int)s2;
	return a > b && a >= b * 2;
}

static void dedup_pass(const gsa_ctx *ctx, std::vector<BlockHdr> &vec, int type, const std::vector<int64_t> &chr_score)
{
	const int64_t genome = ctx->N, two = 2 * ctx->N;
	int n = (int)vec.size();
	if (type == 1) std::sort(vec.begin(), vec.end(), by_query_pos);
	else std::sort(vec.begin(), vec.end(), by_ref_pos);
	for (int i = 0; i < n; i++) {
		if (vec[i].score == 0) continue;
		int64_t H1 = type == 1 ? vec[i].qf : vec[i].rf;
		int64_t T1 = type == 1 ? (int64_t)vec[i].ql + vec[i].lenl - 1 : vec[i].rl + vec[i].lenl - 1;
		int c1 = gsa_host_chr_idx(ctx, vec[i].rf, nullptr);
		if (type == 2 && H1 >= genome) { int64_t t = H1; H1 = two - 1 - T1; T1 = two - 1 - t; } // ReverseRefCoordinate
		for (int j = i + 1; j < n; j++) {
			if (vec[j].score == 0) continue;
			int64_t H2 = type == 1 ? vec[j].qf : vec[j].rf;
			int64_t T2 = type == 1 ? (int64_t)vec[j].ql + vec[j].lenl - 1 : vec[j].rl + vec[j].lenl - 1;
			if (type == 1 && H1 == H2 && T1 == T2) { vec[i].bDup = 1; vec[j].score = 0; continue; }
			int c2 = gsa_host_chr_idx(ctx, vec[j].rf, nullptr);
			if (type == 2 && H2 >= genome) { int64_t t = H2; H2 = two - 1 - T2; T2 = two - 1 - t; }
			if (H2 < T1) {
				int64_t overlap = T2 > T1 ? T1 - H2 : T2 - H2;
				float f1 = 1. * overlap / (T1 - H1), f2 = 1. * overlap / (T2 - H2);
				if ((f1 > f2 && f1 >= 0.9) || (ctx->prm.one_on_one && dup_chr_score(chr_score[c2], chr_score[c1]))) { vec[i].score = 0; break; }
				if ((f2 > f1 && f2 >= 0.9) || (ctx->prm.one_on_one && dup_chr_score(chr_score[c1], chr_score[c2]))) vec[j].score = 0;
			} else break;
		}
	}
	gsa_host_remove_bad(vec);
}

void gsa_host_dedup(const gsa_ctx *ctx, std::vector<BlockHdr> &vec)
{
	for (BlockHdr &b : vec) b.bDup = 0; // src/GSAlign.cpp:510
	std::vector<int64_t> chr_score(ctx->contig_len.size(), 0); // EstChromosomeSimilarity
	for (const BlockHdr &b : vec) chr_score[gsa_host_chr_idx(ctx, b.rf, nullptr)] += b.score;
	dedup_pass(ctx, vec, 1, chr_score);
	dedup_pass(ctx, vec, 2, chr_score);
}
