// blocklogic_harness.cpp -- TEST INFRASTRUCTURE (CPU): pins gsalign_b200/csrc/stdsort.cuh to std::sort and
// gsalign_b200/csrc/block_logic.cuh (the array form of the block logic that also runs in a kernel) to block_logic.cpp (the
// std::vector + std::sort form) on random, tie-heavy inputs.  Prints "same" or the first difference.
#include <algorithm>
#include <random>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "gsa_internal.cuh"
#include "block_logic.cuh"

struct K { int key; int id; };
static bool same_hdr(const BlockHdr &a, const BlockHdr &b)
{
	return a.score == b.score && a.bDup == b.bDup && a.beg == b.beg && a.end == b.end && a.qf == b.qf && a.ql == b.ql && a.lenl == b.lenl && a.rf == b.rf && a.rl == b.rl;
}

static int sort_cases()
{
	std::mt19937_64 rng(12345);
	auto cmp = [](const K &a, const K &b) { return a.key < b.key; };
	for (int rep = 0; rep < 400; rep++) {
		size_t n = rep < 40 ? (size_t)rep : (size_t)(rng() % 5000);
		int distinct = 1 + (int)(rng() % (rep % 3 == 0 ? 3 : rep % 3 == 1 ? 50 : 100000));
		std::vector<K> a(n), b;
		for (size_t i = 0; i < n; i++) { a[i].key = (int)(rng() % distinct); a[i].id = (int)i; }
		if (rep % 7 == 0) std::sort(a.begin(), a.end(), cmp);                       // already sorted
		if (rep % 11 == 0) std::sort(a.begin(), a.end(), [](const K &x, const K &y) { return x.key > y.key; }); // reversed
		b = a;
		std::sort(a.begin(), a.end(), cmp);
		gsa_std_sort(b.data(), b.data() + b.size(), cmp);
		for (size_t i = 0; i < n; i++) if (a[i].key != b[i].key || a[i].id != b[i].id) { printf("sort differs: rep %d n %zu at %zu\n", rep, n, i); return 1; }
	}
	// a median-of-three killer drives introsort into its heap-sort fallback
	for (int n : {1000, 4096, 20001}) {
		std::vector<K> a((size_t)n), b;
		int k = n / 2;
		for (int i = 1; i <= k; i++) { if (i % 2 == 1) { a[(size_t)i - 1].key = i; a[(size_t)i].key = k + i; } a[(size_t)(k + i - 1)].key = 2 * i; }
		for (int i = 0; i < n; i++) a[(size_t)i].id = i;
		b = a;
		std::sort(a.begin(), a.end(), cmp);
		gsa_std_sort(b.data(), b.data() + b.size(), cmp);
		for (int i = 0; i < n; i++) if (a[(size_t)i].key != b[(size_t)i].key || a[(size_t)i].id != b[(size_t)i].id) { printf("killer differs: n %d at %d\n", n, i); return 1; }
	}
	return 0;
}

static long st_split, st_dup, st_removed, st_hz, st_final;
static int logic_cases()
{
	std::mt19937_64 rng(777);
	for (int rep = 0; rep < 3000; rep++) {
		gsa_ctx ctx;
		memset(&ctx.prm, 0, sizeof(ctx.prm));
		ctx.prm.one_on_one = rep % 4 == 3;
		ctx.prm.min_aln_len = 50; ctx.prm.min_block_score = 20;
		const int nctg = 1 + (int)(rng() % 4);
		int64_t total = 0;
		std::vector<int32_t> len((size_t)nctg);
		for (int i = 0; i < nctg; i++) { len[(size_t)i] = 500 + (int)(rng() % 2000); total += len[(size_t)i]; }
		ctx.N = total; ctx.contig_len = len;
		int64_t acc = 0;
		for (int i = 0; i < nctg; i++) {
			ContigEnd f; f.end = acc + len[(size_t)i] - 1; f.idx = i; f.pad = len[(size_t)i];
			acc += len[(size_t)i];
			ContigEnd r; r.end = (2 * total - acc) + len[(size_t)i] - 1; r.idx = i; r.pad = len[(size_t)i];
			ctx.cend.push_back(f); ctx.cend.push_back(r);
		}
		std::sort(ctx.cend.begin(), ctx.cend.end(), [](const ContigEnd &a, const ContigEnd &b) { return a.end < b.end; });
		// blocks over consecutive seed ranges; pieces nest inside them (table 1: a subset of the breaks of table 2)
		const int nb = (int)(rng() % (rep % 10 == 0 ? 200 : 24));
		std::vector<BlockHdr> vec;
		std::vector<Piece> p1, p2;
		int64_t pos = 0;
		const int span = 1 + (int)(rng() % 3) * 40;   // small coordinate ranges: many ties and overlaps
		for (int b = 0; b < nb; b++) {
			const int parts = 1 + (int)(rng() % 4);
			std::vector<Piece> mine;
			for (int t = 0; t < parts; t++) {
				Piece p; memset(&p, 0, sizeof(p));
				p.beg = pos; pos += 1 + (int64_t)(rng() % 20); p.end = pos;
				if (rep % 5 == 1) { // coarse coordinates: exact duplicates (H1 == H2 && T1 == T2) and full overlaps
					p.qf = (int32_t)(rng() % 4) * 100; p.ql = p.qf + (int32_t)(rng() % 3) * 100; p.lenl = 20;
					p.rf = (int64_t)(rng() % 6) * (2 * total / 6); p.rl = p.rf + (int64_t)(rng() % 3) * 100;
				} else {
					p.qf = (int32_t)(rng() % (uint64_t)(span * 10)); p.ql = p.qf + (int32_t)(rng() % 300); p.lenl = 15 + (int32_t)(rng() % 30);
					p.rf = (int64_t)(rng() % (uint64_t)(2 * total)); p.rl = p.rf + (int64_t)(rng() % 300);
				}
				p.sumlen = (int64_t)(rng() % 400);
				mine.push_back(p);
			}
			Piece whole = mine[0]; whole.end = mine.back().end; whole.ql = mine.back().ql; whole.lenl = mine.back().lenl; whole.rl = mine.back().rl;
			whole.sumlen = 0; for (auto &p : mine) whole.sumlen += p.sumlen;
			BlockHdr h = blk_from_piece(whole, 21 + (int32_t)(rng() % 6) * 100);
			vec.push_back(h);
			// table 2 holds every break, table 1 only those between "even" parts
			for (auto &p : mine) p2.push_back(p);
			Piece cur = mine[0];
			for (int t = 1; t < parts; t++) {
				if (rng() % 2) { p1.push_back(cur); cur = mine[(size_t)t]; }
				else { cur.end = mine[(size_t)t].end; cur.ql = mine[(size_t)t].ql; cur.lenl = mine[(size_t)t].lenl; cur.rl = mine[(size_t)t].rl; cur.sumlen += mine[(size_t)t].sumlen; }
			}
			p1.push_back(cur);
		}
		std::vector<BlockHdr> arr = vec;
		const int cap = (int)(vec.size() + p1.size() + p2.size()) + 4;
		arr.resize((size_t)cap);
		// legacy
		gsa_host_split(&ctx, vec, p1, p2);
		const int hz_legacy = ctx.split_hazard;
		gsa_host_dedup(&ctx, vec);
		// array form
		std::vector<int64_t> chr((size_t)nctg);
		BlkParams P; P.ce = ctx.cend.data(); P.nce = (int)ctx.cend.size(); P.genome = ctx.N; P.min_aln_len = ctx.prm.min_aln_len;
		P.min_block_score = ctx.prm.min_block_score; P.one_on_one = ctx.prm.one_on_one; P.chr_score = chr.data(); P.n_contigs = nctg;
		int hz = 0, n = nb;
		n = blk_split(P, arr.data(), n, cap, p1.data(), (int)p1.size(), &hz);
		if (n >= 0) n = blk_split(P, arr.data(), n, cap, p2.data(), (int)p2.size(), &hz);
		if (n >= 0) n = blk_dedup(P, arr.data(), n);
		if (n != (int)vec.size() || hz != hz_legacy) { printf("logic differs: rep %d count %d vs %zu (hazard %d vs %d)\n", rep, n, vec.size(), hz, hz_legacy); return 1; }
		st_hz += hz; st_final += n; st_removed += (nb > n); for (int i = 0; i < n; i++) st_dup += arr[(size_t)i].bDup;
		for (int i = 0; i < n; i++) if (!same_hdr(arr[(size_t)i], vec[(size_t)i])) { printf("logic differs: rep %d block %d of %d\n", rep, i, n); return 1; }
	}
	return 0;
}

int main()
{
	if (sort_cases() || logic_cases()) return 1;
	printf("same (%ld final blocks over all cases, %ld duplicate marks, %ld cases lost blocks, %ld hazard phases)\n", st_final, st_dup, st_removed, st_hz);
	return 0;
}
