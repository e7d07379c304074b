// block_logic.cuh -- the O(#blocks) phases of GenomeComparison on plain arrays, callable on the host and in a kernel:
//   RemoveBadAlnBlocks                  reference src/ProcessCandidateAlignment.cpp:72-79
//   CheckGapsBetweenSeeds (tail)        reference src/ProcessCandidateAlignment.cpp:140-155
//   CheckAlnBlockSpanMultipleRefChrs    reference src/ProcessCandidateAlignment.cpp:100-117
//   EstChromosomeSimilarity             reference src/GSAlign.cpp:393-407
//   RemoveRedundantAlnBlocks            reference src/GSAlign.cpp:415-471   
// Same statements as block_logic.cpp (std::vector + std::sort, the host path and the checker of this file in
// tests/test_boundary_cpu.py), with gsa_std_sort (stdsort.cuh) where that file calls std::sort: float ratios and the tie
// order of libstdc++'s introsort are observable in the MAF record order, so both are kept to the letter.
#pragma once
#include "stdsort.cuh"

struct BlkParams {
	const ContigEnd *ce; int nce;      // ChrLocMap: contig ends on both strands, sorted
	int64_t genome;                    // GenomeSize N (|T| = 2N)
	int32_t min_aln_len, min_block_score, one_on_one;
	int64_t *chr_score; int n_contigs; // scratch: EstChromosomeSimilarity
};

struct BlkByScoreDesc { GSA_HD bool operator()(const BlockHdr &a, const BlockHdr &b) const { return a.score > b.score; } };
struct BlkByQueryPos { GSA_HD bool operator()(const BlockHdr &a, const BlockHdr &b) const { return a.qf == b.qf ? a.score > b.score : a.qf < b.qf; } };
struct BlkByRefPos { GSA_HD bool operator()(const BlockHdr &a, const BlockHdr &b) const { return a.rf == b.rf ? a.score > b.score : a.rf < b.rf; } };

GSA_HD inline int blk_chr_idx(const BlkParams &P, int64_t rpos)
{ // ChrLocMap.lower_bound(rpos): first contig end >= rpos
	int lo = 0, hi = P.nce;
	while (lo < hi) { int m = (lo + hi) >> 1; if (P.ce[m].end < rpos) lo = m + 1; else hi = m; }
	if (lo == P.nce) lo = P.nce - 1;
	return P.ce[lo].idx;
}

GSA_HD inline BlockHdr blk_from_piece(const Piece &p, int32_t score)
{
	BlockHdr b; b.score = score; b.bDup = 0; b.beg = p.beg; b.end = p.end; b.qf = p.qf; b.ql = p.ql; b.lenl = p.lenl;
	b.rf = p.rf; b.rl = p.rl; b.frag_beg = 0; b.n_frags = 0; b.aln_len = 0;
	return b;
}

GSA_HD inline int blk_remove_bad(BlockHdr *v, int n)
{
	gsa_std_sort(v, v + n, BlkByScoreDesc());
	while (n > 0 && v[n - 1].score == 0) n--;
	return n;
}

// one split phase; returns the new block count, -1 if the list would outgrow cap.  *hazard is incremented when the pushes
// cross a power of two (hazard H14: the reference's behaviour is undefined there)
GSA_HD inline int blk_split(const BlkParams &P, BlockHdr *v, int n, int cap, const Piece *pieces, int np, int *hazard)
{
	const int n0 = n;
	for (int i = 0; i < n0; i++) {
		const int64_t beg = v[i].beg, end = v[i].end;
		int lo = 0, hi = np;   // pieces are sorted by beg and nest inside blocks: first piece with beg >= block.beg
		while (lo < hi) { int m = (lo + hi) >> 1; if (pieces[m].beg < beg) lo = m + 1; else hi = m; }
		int cnt = 0;
		while (lo + cnt < np && pieces[lo + cnt].beg < end) cnt++;
		if (cnt <= 1) continue; // no break point inside: the block keeps its score
		v[i].score = 0;
		for (int t = 0; t < cnt; t++) {
			const Piece &p = pieces[lo + t];
			// CalAlnBlockScore, src/ProcessCandidateAlignment.cpp:26-36
			const int32_t sc = (p.ql + p.lenl - p.qf) < P.min_aln_len ? 0 : (int32_t)p.sumlen;
			if (sc > P.min_block_score) { if (n >= cap) return -1; v[n++] = blk_from_piece(p, sc); }
		}
	}
	if (n > n0) { int p2 = 1; while (p2 < n0) p2 <<= 1; if (n0 == 0 || n > p2) (*hazard)++; }
	return blk_remove_bad(v, n);
}

GSA_HD inline bool blk_dup_chr_score(int64_t s1, int64_t s2)
{ // CheckDuplicatedChrScore(int,int), src/GSAlign.cpp:409-413 (arguments are truncated to int there)
	const int a = (int)s1, b = (int)s2;
	return a > b && a >= b * 2;
}

GSA_HD inline int blk_dedup_pass(const BlkParams &P, BlockHdr *v, int n, int type)
{
	const int64_t genome = P.genome, two = 2 * P.genome;
	if (type == 1) gsa_std_sort(v, v + n, BlkByQueryPos());
	else gsa_std_sort(v, v + n, BlkByRefPos());
	for (int i = 0; i < n; i++) {
		if (v[i].score == 0) continue;
		int64_t H1 = type == 1 ? v[i].qf : v[i].rf;
		int64_t T1 = type == 1 ? (int64_t)v[i].ql + v[i].lenl - 1 : v[i].rl + v[i].lenl - 1;
		const int c1 = blk_chr_idx(P, v[i].rf);
		if (type == 2 && H1 >= genome) { int64_t t = H1; H1 = two - 1 - T1; T1 = two - 1 - t; } // ReverseRefCoordinate
		for (int j = i + 1; j < n; j++) {
			if (v[j].score == 0) continue;
			int64_t H2 = type == 1 ? v[j].qf : v[j].rf;
			int64_t T2 = type == 1 ? (int64_t)v[j].ql + v[j].lenl - 1 : v[j].rl + v[j].lenl - 1;
			if (type == 1 && H1 == H2 && T1 == T2) { v[i].bDup = 1; v[j].score = 0; continue; }
			const int c2 = blk_chr_idx(P, v[j].rf);
			if (type == 2 && H2 >= genome) { int64_t t = H2; H2 = two - 1 - T2; T2 = two - 1 - t; }
			if (H2 < T1) {
				const int64_t overlap = T2 > T1 ? T1 - H2 : T2 - H2;
				// float f = 1. * overlap / (T - H): a double division rounded to float, then compared with the double constant 0.9
				const float f1 = (float)(1. * (double)overlap / (double)(T1 - H1)), f2 = (float)(1. * (double)overlap / (double)(T2 - H2));
				if ((f1 > f2 && (double)f1 >= 0.9) || (P.one_on_one && blk_dup_chr_score(P.chr_score[c2], P.chr_score[c1]))) { v[i].score = 0; break; }
				if ((f2 > f1 && (double)f2 >= 0.9) || (P.one_on_one && blk_dup_chr_score(P.chr_score[c1], P.chr_score[c2]))) v[j].score = 0;
			} else break;
		}
	}
	return blk_remove_bad(v, n);
}

GSA_HD inline int blk_dedup(const BlkParams &P, BlockHdr *v, int n)
{
	for (int i = 0; i < n; i++) v[i].bDup = 0; // src/GSAlign.cpp:510
	for (int c = 0; c < P.n_contigs; c++) P.chr_score[c] = 0; // EstChromosomeSimilarity
	for (int i = 0; i < n; i++) P.chr_score[blk_chr_idx(P, v[i].rf)] += v[i].score;
	n = blk_dedup_pass(P, v, n, 1);
	return blk_dedup_pass(P, v, n, 2);
}
