// emit_harness.cpp -- TEST INFRASTRUCTURE: drives the CLI's emitters (gsalign_b200/csrc/host/emit.cpp) on alignment
// records read from a file, so that MAF / ALN / VCF formatting can be checked against the unmodified reference on a
// machine without a GPU (the records come from the reference's own containers, tests/test_emit_cpu.py).
//   emit_harness <index prefix as given to -i> <query.fa> <records.bin> <out prefix> <fmt 1|2> <threads>
// records.bin, per query contig in file order: int32 n_blocks; gsa_block[n_blocks]; int64 n_frags; gsa_frag[n_frags];
// int64 aln_bytes; char aln1[aln_bytes]; char aln2[aln_bytes]
#include <stdlib.h>
#include "host.h"

template <typename T> static bool rd(FILE *f, T *p, size_t n) { return n == 0 || fread(p, sizeof(T), n, f) == n; }

int main(int argc, char **argv)
{
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
	FILE *f = fopen(argv[3], "rb");
	if (!f) return 1;
	EmitState st; st.threads = o.threads > 0 ? o.threads : 1;
	for (int qi = 0; qi < (int)query.size(); qi++) {
		ContigResult r;
		int32_t nb = 0; int64_t nf = 0, ab = 0;
		if (!rd(f, &nb, 1)) return 1;
		r.blocks.resize((size_t)nb);
		if (!rd(f, r.blocks.data(), (size_t)nb) || !rd(f, &nf, 1)) return 1;
		r.frags.resize((size_t)nf);
		if (!rd(f, r.frags.data(), (size_t)nf) || !rd(f, &ab, 1)) return 1;
		r.aln1.resize((size_t)ab); r.aln2.resize((size_t)ab);
		if (!rd(f, &r.aln1[0], (size_t)ab) || !rd(f, &r.aln2[0], (size_t)ab)) return 1;
		if (nb == 0) continue; // src/GSAlign.cpp:541: a contig without alignments never reaches the writers
		if (o.out_format == 1) output_maf(o, ix, query, qi, r);
		if (o.out_format == 2) output_aln(o, ix, query, qi, r);
		variant_identification(ix, query, qi, r, st);
	}
	fclose(f);
	output_variants(o, ix, st);
	return 0;
}
