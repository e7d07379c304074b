// host.h -- host side of bin/GSAlign: index/FASTA I/O and the MAF / ALN / VCF emitters.
// Mirrors the reference's CLI surface (src/main.cpp), loaders (src/bwt_index.cpp) and emitters
// (src/tools.cpp, src/SeqVariant.cpp) so that the files it writes are byte-identical; all of the
// seed -> cluster -> fill work goes through the C ABI in include/gsalign_b200.h.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "../../../include/gsalign_b200.h"

struct HostIndex {                 // bwaidx_t as loaded by bwa_idx_load (reference src/bwt_index.cpp:147-159)
	// .bwt / .sa / .pac are mapped, not copied: the arrays below point into the mappings (private, so that sa[0] can be
	// set to -1 like bwt_restore_sa does) and go to the GPU straight from the page cache
	const uint32_t *bwt = nullptr; uint64_t bwt_size = 0;   // words after the 5 x u64 header
	uint64_t primary = 0, L2[5] = {0, 0, 0, 0, 0}, seq_len = 0;
	const uint64_t *sa = nullptr; uint64_t n_sa = 0;        // sa[0] = (uint64_t)-1
	int sa_intv = 32;
	const uint8_t *pac = nullptr;
	int64_t l_pac = 0;
	struct Mapping { void *p = nullptr; size_t n = 0; } maps[3];
	HostIndex() {}
	HostIndex(const HostIndex &) = delete;
	HostIndex &operator=(const HostIndex &) = delete;
	~HostIndex();
	std::vector<std::string> names; // ChromosomeVec[i].name
	std::vector<int64_t> offset;    // FowardLocation
	std::vector<int32_t> len;
	// ChrLocMap (reference src/bwt_index.cpp:247-252): inclusive contig ends on both strands, sorted
	std::vector<std::pair<int64_t, int> > chr_loc;
	bool load(const std::string &prefix, std::string &err);
	void view(gsa_index_view *v) const;
	int64_t genome() const { return l_pac; }
	int64_t reverse_location(int i) const { return 2 * l_pac - (offset[i] + len[i]); }
	// RefSequence[pos] (reference src/bwt_index.cpp:193-212): forward base or its complement on the mirrored half
	char text(int64_t pos) const
	{
		int64_t f = pos < l_pac ? pos : 2 * l_pac - 1 - pos;
		int b = pac[(size_t)(f >> 2)] >> ((~f & 3) << 1) & 3;
		return "ACGT"[pos < l_pac ? b : 3 - b];
	}
	// ChrLocMap.lower_bound(pos)
	const std::pair<int64_t, int> &loc(int64_t pos) const;
};

struct QueryChr { std::string name, seq; };  // QueryChr_t, reference src/structure.h:134-138

bool check_input_file(const char *path);                                   // CheckInputFile, src/main.cpp:49-64
bool load_query_file(const char *path, std::vector<QueryChr> &out);        // LoadQueryFile,  src/main.cpp:82-114
std::string trim_chromosome_name(std::string name);                        // TrimChromosomeName, src/main.cpp:35-47

struct Coordinate { bool bDir; int gPos; int ChromosomeIdx; };            // Coordinate_t
Coordinate gen_coordinate(const HostIndex &ix, int64_t rPos);              // GenCoordinateInfo, src/tools.cpp:120-140

// Variant_t without its two std::strings: the alleles (REF then ALT) live in EmitState::alleles at `off`
struct Variant { int32_t chr_idx, pos, type; uint32_t ref_len, alt_len; uint64_t off; };

struct Options {                   // the globals of src/main.cpp:10-12,203-215
	int threads = 8, out_format = 1, n_gpus = 1, lanes = 4;
	bool sensitive = false, show_plot = false, debug = false, vcf = true, allow_dup = true, one_on_one = false;
	int min_seed_len = 15, min_block_score = 200, min_aln_len = 200, min_idy = 70, max_indel = 25;
	const char *ref_fa = nullptr, *index_prefix = nullptr, *query = nullptr, *out_prefix = nullptr, *gnuplot = nullptr;
	std::string maf, aln, vcf_name;
};

// One contig's result, detached from the library's pinned buffers (needed when several GPUs run ahead of the emitter)
struct ContigResult {
	std::vector<gsa_block> blocks;
	std::vector<gsa_frag> frags;
	std::string aln1, aln2;
	// N3: the variant records found on the device (gsa_variants), with the range of every block; empty + have_vars false =
	// the host scans the rows itself (records that came through the multi-GPU gather carry no variant list)
	std::vector<gsa_variant> vars;
	std::vector<int64_t> var_first, var_count;
	bool have_vars = false;
	void assign(const gsa_alignment &a);
	// a record of a gathered outbox image (gsa_record_next): the fragment list may have travelled in the compact form
	int assign_record(const gsa_alignment &a, const void *image, int64_t bytes, int64_t record_offset);
	void assign_variants(const gsa_variant_list &v);
};

// A growing array of plain records whose new elements are NOT initialised: grow() hands out room that the caller's threads
// fill (and first touch) themselves.  A human-size pair collects 33 M variant records (1 GB): value-initialising and then
// copying them on one thread cost more than finding them.  realloc() of a large block is a remap, not a copy.
template <typename T> struct RawArray {
	T *p = nullptr; size_t n = 0, cap = 0;
	RawArray() {}
	RawArray(const RawArray &) = delete;
	RawArray &operator=(const RawArray &) = delete;
	~RawArray() { free(p); }
	T *grow(size_t k)                // room for k more elements at the end; nullptr when memory is out (size unchanged)
	{
		if (n + k > cap) {
			size_t want = cap + cap / 2; if (want < n + k) want = n + k; if (want < 4096) want = 4096;
			T *q = (T *)realloc(p, want * sizeof(T));
			if (!q) return nullptr;
			p = q; cap = want;
		}
		T *r = p + n; n += k; return r;
	}
	size_t size() const { return n; }
	const T *data() const { return p; }
	T &operator[](size_t i) { return p[i]; }
	const T &operator[](size_t i) const { return p[i]; }
};

struct EmitState {                 // running totals of GenomeComparison (src/GSAlign.cpp:14-15)
	int64_t total_aln_len = 0, total_matches = 0, local_aln_num = 0, dup_num = 0;
	int iSNV = 0, iInsertion = 0, iDeletion = 0;
	RawArray<Variant> variants;      // in the order VariantIdentification pushes them (the input order of the final unstable sort)
	RawArray<char> alleles;
	int threads = 1;                 // -t: host threads the emitters may use (row assembly, variant scan, VCF formatting)
};

// OutputMAF / OutputAlignment (src/tools.cpp:149-286); they trim a block that runs past its contig end
// (iExtension) by MUTATING the result, exactly like the reference does before VariantIdentification runs
void output_maf(const Options &o, const HostIndex &ix, const std::vector<QueryChr> &q, int qidx, ContigResult &r);
void output_aln(const Options &o, const HostIndex &ix, const std::vector<QueryChr> &q, int qidx, ContigResult &r);
void variant_identification(const HostIndex &ix, const std::vector<QueryChr> &q, int qidx, const ContigResult &r, EmitState &st); // src/SeqVariant.cpp:12-119
void output_variants(const Options &o, const HostIndex &ix, EmitState &st);                                                      // src/SeqVariant.cpp:121-143
// The MAF and VCF writers hand their buffers to a writer thread: this waits until every byte is in the page cache (false if a
// write failed).  Static destruction does the same; a process that leaves through _exit() must call it first.
bool emit_drain();
