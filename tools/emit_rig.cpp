// emit_rig.cpp -- DEVELOPMENT TOOL (no GPU needed): fabricates the alignment records of a human-scale contig pair so that the
// CLI's emitters (gsalign_b200/csrc/host/emit.cpp, through tests/emit_harness.cpp) can be timed at the size of a C4 contig
// on a machine without a GPU.  The records are NOT an aligner's output: a reference contig is drawn at random, SNVs and
// indels are placed at the workload's rates, every exact stretch of >= 15 bases becomes a seed fragment and what lies between
// two seeds becomes one gap fragment whose rows are the true alignment (ungapped where the reference would copy ungapped).  One block per contig, like the real output of the
// synthetic BASELINE pairs.
//   emit_rig <dir> <contigs> <bp per contig> <p_snv> <p_indel> [seed [ext]]      (ext > 0: every block but the last runs ext
//                                                                                 bases past the end of its reference contig)
// writes <dir>/ref.{pac,ann,amb,bwt,sa} (bwt/sa are header-only stubs: the emitters never read them), <dir>/qry.fa and
// <dir>/records.bin (the stream tests/emit_harness.cpp reads).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>
#include "../include/gsalign_b200.h"

static uint64_t rng_state = 88172645463325252ull;
static inline uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static inline double rnd01() { return (double)(rnd() >> 11) * (1.0 / 9007199254740992.0); }

int main(int argc, char **argv)
{
	if (argc < 6) { fprintf(stderr, "usage: emit_rig dir contigs bp p_snv p_indel [seed [ext]]\n"); return 2; }
	const std::string dir = argv[1];
	const int K = atoi(argv[2]); const int64_t L = atoll(argv[3]);
	const double p_snv = atof(argv[4]), p_indel = atof(argv[5]), p_ev = p_snv + p_indel;
	if (argc > 6) rng_state ^= (uint64_t)atoll(argv[6]) * 0x9E3779B97F4A7C15ull;
	const int ext = argc > 7 ? atoi(argv[7]) : 0;
	const int64_t l_pac = (int64_t)K * L;
	std::vector<uint8_t> pac((size_t)(l_pac / 4 + 2), 0);
	FILE *fq = fopen((dir + "/qry.fa").c_str(), "wb"), *fr = fopen((dir + "/records.bin").c_str(), "wb");
	if (!fq || !fr) return 1;
	static const char ACGT[] = "ACGT";
	std::vector<std::string> refs((size_t)K);
	for (int c = 0; c < K; c++) {
		refs[(size_t)c].assign((size_t)L, 'A');
		for (int64_t i = 0; i < L; i++) { int b = (int)(rnd() >> 62); refs[(size_t)c][(size_t)i] = ACGT[b]; int64_t g = (int64_t)c * L + i; pac[(size_t)(g >> 2)] |= (uint8_t)(b << ((~g & 3) << 1)); }
	}
	for (int c = 0; c < K; c++) {
		const std::string &ref = refs[(size_t)c];
		std::string qry, a1, a2;
		qry.reserve((size_t)(L + L / 50));
		std::vector<gsa_frag> frags;
		std::string g1, g2;                       // rows of the gap fragment being collected
		int64_t g_r = 0, g_q = 0, score = 0, cols = 0;
		auto flush_gap = [&] {
			if (g1.empty()) return;
			{ // GenerateFragAlignment's rule (src/ProcessCandidateAlignment.cpp:290-351): a fragment of equal lengths whose ungapped
			  // comparison shows at most 5 mismatches is copied ungapped -- in particular a 1 x 1 fragment is always one column
				std::string r1, r2;
				for (char ch : g1) if (ch != '-') r1 += ch;
				for (char ch : g2) if (ch != '-') r2 += ch;
				if (r1.size() == r2.size() && r1.size() != g1.size()) {
					int mis = 0;
					for (size_t i = 0; i < r1.size(); i++) mis += r1[i] != r2[i];
					if (mis <= 5) { g1 = r1; g2 = r2; }
				}
			}
			gsa_frag f; memset(&f, 0, sizeof(f));
			f.rPos = (int64_t)c * L + g_r; f.qPos = (int32_t)g_q; f.bSeed = 0; f.aln_off = (int64_t)a1.size(); f.aln_len = (int32_t)g1.size();
			for (size_t i = 0; i < g1.size(); i++) { f.rLen += g1[i] != '-'; f.qLen += g2[i] != '-'; score += g1[i] == g2[i]; }
			cols += (int64_t)g1.size();
			a1 += g1; a2 += g2; g1.clear(); g2.clear();
			frags.push_back(f);
		};
		int64_t r = 0;                            // next reference position to account for
		bool first = true;
		while (r < L) {
			// distance to the next event (geometric)
			int64_t m = p_ev > 0 ? (int64_t)(-log(1.0 - rnd01()) / p_ev) : L;
			if (first && m < 20) m = 20;
			if (r + m > L) m = L - r;
			const bool last = r + m >= L;
			if (m >= 15 || first || (last && g1.empty())) { // an exact stretch long enough to be a seed
				flush_gap();
				gsa_frag f; memset(&f, 0, sizeof(f));
				f.rPos = (int64_t)c * L + r; f.qPos = (int32_t)qry.size(); f.qLen = f.rLen = (int32_t)m; f.bSeed = 1; f.aln_len = (int32_t)m;
				frags.push_back(f);
				qry.append(ref, (size_t)r, (size_t)m); score += m; cols += m;
			} else if (m > 0) {
				if (g1.empty()) { g_r = r; g_q = (int64_t)qry.size(); }
				g1.append(ref, (size_t)r, (size_t)m); g2.append(ref, (size_t)r, (size_t)m); qry.append(ref, (size_t)r, (size_t)m);
			}
			r += m; first = false;
			if (r >= L) break;
			if (g1.empty()) { g_r = r; g_q = (int64_t)qry.size(); }
			if (rnd01() * p_ev < p_snv) { // SNV
				char x = ACGT[(strchr(ACGT, ref[(size_t)r]) - ACGT + 1 + (int)(rnd() % 3)) & 3];
				g1 += ref[(size_t)r]; g2 += x; qry += x; r++;
			} else {
				int len = 1 + (int)(rnd() % 10);
				if (rnd() & 1) { for (int i = 0; i < len; i++) { char x = ACGT[rnd() >> 62]; g1 += '-'; g2 += x; qry += x; } }
				else { if (r + len > L - 20) len = 1; for (int i = 0; i < len && r < L; i++, r++) { g1 += ref[(size_t)r]; g2 += '-'; } }
			}
		}
		flush_gap();
		if (!frags.back().bSeed) { fprintf(stderr, "rig: contig ends in a gap fragment (harmless)\n"); }
		else if (ext > 0 && c + 1 < K) { // the last seed runs `ext` bases into the next reference contig: what iExtension trims
			frags.back().qLen += ext; frags.back().rLen += ext; frags.back().aln_len += ext;
			qry.append(refs[(size_t)c + 1], 0, (size_t)ext); score += ext; cols += ext;
		}
		gsa_block b; memset(&b, 0, sizeof(b));
		b.score = (int32_t)score; b.aln_len = (int32_t)cols; b.n_frags = (int32_t)frags.size(); b.frag_beg = 0;
		int32_t nb = 1; int64_t nf = (int64_t)frags.size(), ab = (int64_t)a1.size();
		fwrite(&nb, 4, 1, fr); fwrite(&b, sizeof(b), 1, fr); fwrite(&nf, 8, 1, fr); fwrite(frags.data(), sizeof(gsa_frag), frags.size(), fr);
		fwrite(&ab, 8, 1, fr); fwrite(a1.data(), 1, a1.size(), fr); fwrite(a2.data(), 1, a2.size(), fr);
		fprintf(fq, ">qchr%d\n", c + 1);
		for (size_t i = 0; i < qry.size(); i += 80) { fwrite(qry.data() + i, 1, std::min<size_t>(80, qry.size() - i), fq); fputc('\n', fq); }
		fprintf(stderr, "rig: contig %d: %zu fragments, %lld columns, %zu row bytes\n", c + 1, frags.size(), (long long)cols, a1.size());
	}
	fclose(fq); fclose(fr);
	FILE *f = fopen((dir + "/ref.pac").c_str(), "wb"); fwrite(pac.data(), 1, (size_t)(l_pac / 4 + 1), f); uint8_t tail = (uint8_t)(l_pac % 4); fwrite(&tail, 1, 1, f); fclose(f);
	f = fopen((dir + "/ref.ann").c_str(), "w"); fprintf(f, "%lld %d 11\n", (long long)l_pac, K);
	for (int c = 0; c < K; c++) fprintf(f, "0 chr%d (null)\n%lld %lld 0\n", c + 1, (long long)c * L, (long long)L);
	fclose(f);
	f = fopen((dir + "/ref.amb").c_str(), "w"); fprintf(f, "%lld %d 0\n", (long long)l_pac, K); fclose(f);
	uint64_t h[8] = {1, (uint64_t)l_pac / 2, (uint64_t)l_pac, (uint64_t)l_pac * 3 / 2, (uint64_t)l_pac * 2, 0, 0, 0};
	f = fopen((dir + "/ref.bwt").c_str(), "wb"); fwrite(h, 8, 5, f); fclose(f);
	uint64_t s[16]; memset(s, 0, sizeof(s)); memcpy(s, h, 40); s[5] = 1u << 30; s[6] = (uint64_t)l_pac * 2;
	f = fopen((dir + "/ref.sa").c_str(), "wb"); fwrite(s, 8, 16, f); fclose(f);
	return 0;
}
