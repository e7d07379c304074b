// emit_harness.cpp -- TEST INFRASTRUCTURE: drives the CLI's emitters (gsalign_b200/csrc/host/emit.cpp) on alignment
// records read from a file, so that MAF / ALN / VCF formatting can be checked against the unmodified reference on a
// machine without a GPU (the records come from the reference's own containers, tests/test_emit_cpu.py).
//   emit_harness <index prefix as given to -i> <query.fa> <records.bin> <out prefix> <fmt 1|2> <threads>
// records.bin, per query contig in file order: int32 n_blocks; gsa_block[n_blocks]; int64 n_frags; gsa_frag[n_frags];
// int64 aln_bytes; char aln1[aln_bytes]; char aln2[aln_bytes]
#include <stdlib.h>
#include <algorithm>
#include <chrono>
#include "host.h"

template <typename T> static bool rd(FILE *f, T *p, size_t n) { return n == 0 || fread(p, sizeof(T), n, f) == n; }

void gsa_test_sort_keys(uint64_t *keys, uint32_t *idx, size_t n, int threads); // emit.cpp

// emit_harness sort <n> <distinct keys> <threads> <seed> [pattern]: the emitters' threaded sort against std::sort on the same
// (key, index) records -- the order of equal keys must come out the same (hazard H5).  pattern 0: random keys; 1: ascending
// runs with repeated neighbours (what VariantIdentification pushes: one sorted stretch per block); 2: descending; 3: sorted
static int sort_mode(char **argv, int pattern)
{
	size_t n = (size_t)atoll(argv[2]); uint64_t distinct = (uint64_t)atoll(argv[3]); int threads = atoi(argv[4]);
	uint64_t x = (uint64_t)atoll(argv[5]) * 0x9E3779B97F4A7C15ull + 1;
	struct K { uint64_t key; uint32_t idx; };
	std::vector<K> want(n);
	std::vector<uint64_t> keys(n); std::vector<uint32_t> idx(n);
	for (size_t i = 0; i < n; i++) {
		x ^= x << 13; x ^= x >> 7; x ^= x << 17;
		keys[i] = want[i].key = x % distinct; idx[i] = want[i].idx = (uint32_t)i;
	}
	if (pattern == 1) { // runs: a new run starts with probability 1/50000, inside a run the key grows by 0 (1 in 16) or a small step
		uint64_t k = 0;
		for (size_t i = 0; i < n; i++) {
			const uint64_t r = keys[i] ^ (keys[i] >> 29);
			if (i == 0 || r % 50000 == 0) k = r % distinct; else k += (r >> 20) % 16 == 0 ? 0 : 1 + (r >> 24) % 200;
			keys[i] = want[i].key = k;
		}
	} else if (pattern == 2 || pattern == 3) {
		std::sort(keys.begin(), keys.end());
		if (pattern == 2) std::reverse(keys.begin(), keys.end());
		for (size_t i = 0; i < n; i++) want[i].key = keys[i];
	}
	std::sort(want.begin(), want.end(), [](const K &a, const K &b) { return a.key < b.key; });
	gsa_test_sort_keys(keys.data(), idx.data(), n, threads);
	for (size_t i = 0; i < n; i++) if (keys[i] != want[i].key || idx[i] != want[i].idx) { printf("differs at %zu\n", i); return 1; }
	printf("same\n");
	return 0;
}

// GSA_HARNESS_VARS=1: the variant records gsa_variants() would deliver (include/gsalign_b200.h, gsa_variant_kind) are derived
// from the rows here, on the host, and handed to the emitters the way bin/GSAlign hands them the device's list -- so that
// the allele fetch of that path (records_to_variants) is checked against the row scan without a GPU.
static int nt4_of(char c) { switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; } }
static void derive_variant_records(const HostIndex &ix, ContigResult &r, std::vector<gsa_variant> &vars, std::vector<int64_t> &first, std::vector<int64_t> &count)
{
	auto push = [&](int kind, int64_t rPos, int qPos, int len) {
		gsa_variant v; v.rPos = rPos; v.qPos = qPos; v.gPos = gen_coordinate(ix, rPos).gPos; v.len = len; v.kind = kind; vars.push_back(v);
	};
	for (const gsa_block &b : r.blocks) {
		first.push_back((int64_t)vars.size());
		for (int64_t t = b.frag_beg; t < b.frag_beg + b.n_frags; t++) {
			const gsa_frag &f = r.frags[(size_t)t];
			if (f.bSeed || (f.qLen == 0 && f.rLen == 0)) continue;
			if (f.qLen == 0) push(GSA_VAR_FRAG_DEL, f.rPos - 1, f.qPos - 1, f.rLen);
			else if (f.rLen == 0) push(GSA_VAR_FRAG_INS, f.rPos - 1, f.qPos - 1, f.qLen);
			else if (f.qLen == 1 && f.rLen == 1) {
				char c1 = r.aln1[(size_t)f.aln_off], c2 = r.aln2[(size_t)f.aln_off];
				if (nt4_of(c1) != nt4_of(c2) && nt4_of(c2) != 4) push(GSA_VAR_SNV, f.rPos, f.qPos, 1);
			} else {
				const char *a1 = r.aln1.data() + f.aln_off, *a2 = r.aln2.data() + f.aln_off;
				int qpos = f.qPos; int64_t rpos = f.rPos;
				for (int i = 0; i < f.aln_len; i++) {
					if (a1[i] == '-') { int ind = 1; while (i + ind < f.aln_len && a1[i + ind] == '-') ind++; push(GSA_VAR_INS, rpos - 1, qpos - 1, ind); qpos += ind; i += ind - 1; }
					else if (a2[i] == '-') { int ind = 1; while (i + ind < f.aln_len && a2[i + ind] == '-') ind++; push(GSA_VAR_DEL, rpos - 1, qpos - 1, ind); rpos += ind; i += ind - 1; }
					else { if (nt4_of(a1[i]) != nt4_of(a2[i]) && nt4_of(a2[i]) != 4) push(GSA_VAR_SNV, rpos, qpos, 1); rpos++; qpos++; }
				}
			}
		}
		count.push_back((int64_t)vars.size() - first.back());
	}
}

// GSA_TIMING=1: wall clock per stage on stderr (tools/emit_rig.cpp makes human-scale inputs for this)
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
	const bool timing = getenv("GSA_TIMING") != nullptr;
	double t_read = 0, t_maf = 0, t_var = 0, t0 = now_s();
	if ((argc == 6 || argc == 7) && std::string(argv[1]) == "sort") return sort_mode(argv, argc == 7 ? atoi(argv[6]) : 0);
	if (argc != 7) return 2;
	Options o;
	o.index_prefix = argv[1]; o.query = argv[2];
	std::string op = argv[4];
	o.out_format = atoi(argv[5]); o.threads = atoi(argv[6]);
	if (o.out_format == 1) o.maf = op + ".maf";
	if (o.out_format == 2) o.aln = op + ".aln";
	o.vcf_name = op + ".vcf";
	HostIndex ix; std::string err;
	if (!ix.load(argv[1], err)) { fprintf(stderr, "index: %s\n", err.c_str()); return 1; }
	std::vector<QueryChr> query;
	if (!load_query_file(argv[2], query)) return 1;
	if (timing) fprintf(stderr, "[timing] index + query loaded %.3f s\n", now_s() - t0);
	FILE *f = fopen(argv[3], "rb");
	if (!f) return 1;
	EmitState st; st.threads = o.threads > 0 ? o.threads : 1;
	for (int qi = 0; qi < (int)query.size(); qi++) {
		ContigResult r;
		double t1 = now_s();
		int32_t nb = 0; int64_t nf = 0, ab = 0;
		if (!rd(f, &nb, 1)) return 1;
		r.blocks.resize((size_t)nb);
		if (!rd(f, r.blocks.data(), (size_t)nb) || !rd(f, &nf, 1)) return 1;
		r.frags.resize((size_t)nf);
		if (!rd(f, r.frags.data(), (size_t)nf) || !rd(f, &ab, 1)) return 1;
		r.aln1.resize((size_t)ab); r.aln2.resize((size_t)ab);
		if (!rd(f, &r.aln1[0], (size_t)ab) || !rd(f, &r.aln2[0], (size_t)ab)) return 1;
		if (nb == 0) continue; // src/GSAlign.cpp:541: a contig without alignments never reaches the writers
		double t2 = now_s();
		if (o.out_format == 1) output_maf(o, ix, query, qi, r);
		if (o.out_format == 2) output_aln(o, ix, query, qi, r);
		if (getenv("GSA_HARNESS_VARS")) { // (after the writer: iExtension has trimmed the last seed by now, as in bin/GSAlign's order of effects)
			std::vector<gsa_variant> vars; std::vector<int64_t> first, count;
			derive_variant_records(ix, r, vars, first, count);
			gsa_variant_list vl; vl.n_variants = (int64_t)vars.size(); vl.variants = vars.data(); vl.block_first = first.data(); vl.block_count = count.data();
			r.assign_variants(vl);
		}
		double t3 = now_s();
		variant_identification(ix, query, qi, r, st);
		t_read += t2 - t1; t_maf += t3 - t2; t_var += now_s() - t3;
	}
	fclose(f);
	double t4 = now_s();
	output_variants(o, ix, st);
	double t5 = now_s();
	bool ok = emit_drain();
	if (timing) fprintf(stderr, "[timing] records read %.3f s, alignment file %.3f s, variant scan %.3f s, output_variants %.3f s, drain %.3f s\n", t_read, t_maf, t_var, t5 - t4, now_s() - t5);
	return ok ? 0 : 1;
}
