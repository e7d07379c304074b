// emit.cpp -- MAF / ALN / VCF emitters of bin/GSAlign, byte-compatible with the reference
// (src/tools.cpp:3-44,142-286 and src/SeqVariant.cpp:6-143; format hazards H3-H8, H15 of SURVEY.md).
#include <string.h>
#include <algorithm>
#include "host.h"

static const char *VERSION_STR = "1.0.22"; // VersionStr, src/main.cpp:9 (printed in the VCF header)

void ContigResult::assign(const gsa_alignment &a)
{
	blocks.assign(a.blocks, a.blocks + a.n_blocks);
	frags.assign(a.frags, a.frags + a.n_frags);
	aln1.assign(a.aln1 ? a.aln1 : "", (size_t)a.aln_bytes);
	aln2.assign(a.aln2 ? a.aln2 : "", (size_t)a.aln_bytes);
}

static inline int nt4(char ch)
{ // nst_nt4_table
	switch (ch) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': return 3;
	default: return 4;
	}
}

// ReverseMap (src/tools.cpp:3-31): acgtnu are upper-cased while complementing, U -> A, '-' stays, the rest -> NUL
static inline char reverse_map(char c)
{
	switch (c) {
	case 'A': case 'a': return 'T';
	case 'C': case 'c': return 'G';
	case 'G': case 'g': return 'C';
	case 'T': case 't': case 'U': case 'u': return 'A';
	case 'N': case 'n': return 'N';
	case '-': return '-';
	default: return '\0';
	}
}

static void self_complementary(size_t len, char *seq)
{ // SelfComplementarySeq, src/tools.cpp:33-44
	if (len == 0) return;
	size_t i = 0, j = len - 1;
	for (; i < j; i++, j--) { char a = seq[i], b = seq[j]; seq[i] = reverse_map(b); seq[j] = reverse_map(a); }
	if (i == j) seq[i] = reverse_map(seq[i]);
}

static int count_gaps(const char *aln, int i, int stop)
{
	int n = 0;
	for (; i < stop; i++) if (aln[i] == '-') n++;
	return n;
}

// assembles the two rows of a block (src/tools.cpp:169-184): inside seeds BOTH rows are copied from the query
static void build_rows(const QueryChr &qc, const ContigResult &r, const gsa_block &b, std::vector<char> &a1, std::vector<char> &a2)
{
	a1.assign((size_t)b.aln_len + 1, '\0'); a2.assign((size_t)b.aln_len + 1, '\0');
	size_t pos = 0;
	for (int64_t t = b.frag_beg; t < b.frag_beg + b.n_frags; t++) {
		const gsa_frag &f = r.frags[(size_t)t];
		if (f.bSeed) {
			memcpy(a1.data() + pos, qc.seq.data() + f.qPos, (size_t)f.qLen);
			memcpy(a2.data() + pos, qc.seq.data() + f.qPos, (size_t)f.qLen);
			pos += (size_t)f.qLen;
		} else {
			memcpy(a1.data() + pos, r.aln1.data() + f.aln_off, (size_t)f.aln_len);
			memcpy(a2.data() + pos, r.aln2.data() + f.aln_off, (size_t)f.aln_len);
			pos += (size_t)f.aln_len;
		}
	}
}

// iExtension (src/tools.cpp:192-202): a block whose last seed runs past the end of its contig is trimmed in place
static void trim_extension(const HostIndex &ix, const Coordinate &coor, ContigResult &r, gsa_block &b, std::vector<char> &a1, std::vector<char> &a2)
{
	gsa_frag &last = r.frags[(size_t)(b.frag_beg + b.n_frags - 1)];
	int idx = coor.ChromosomeIdx;
	int64_t end = last.rPos + last.rLen, lim = (coor.bDir ? ix.offset[idx] : ix.reverse_location(idx)) + ix.len[idx];
	int ext = end > lim ? (int)(end - lim) : 0;
	if (ext > 0) {
		b.aln_len -= ext; b.score -= ext; last.rLen -= ext; last.qLen -= ext;
		a1[(size_t)b.aln_len] = a2[(size_t)b.aln_len] = '\0';
	}
}

static void padded_names(const HostIndex &ix, const QueryChr &qc, int ref_idx, std::string &qname, std::string &rname)
{ // src/tools.cpp:187-189
	qname = qc.name; rname = ix.names[(size_t)ref_idx];
	if (qname.length() > rname.length()) rname += std::string(qname.length() - rname.length(), ' ');
	else qname += std::string(rname.length() - qname.length(), ' ');
}

void output_maf(const Options &o, const HostIndex &ix, const std::vector<QueryChr> &q, int qidx, ContigResult &r)
{
	FILE *out;
	if (qidx == 0) { out = fopen(o.maf.c_str(), "w"); if (out) fprintf(out, "##maf version=1\n"); }
	else out = fopen(o.maf.c_str(), "a");
	if (!out) return;
	const QueryChr &qc = q[(size_t)qidx];
	std::vector<char> a1, a2;
	std::string qname, rname;
	for (gsa_block &b : r.blocks) {
		if (!o.allow_dup && b.bDup) continue;
		build_rows(qc, r, b, a1, a2);
		Coordinate coor = gen_coordinate(ix, r.frags[(size_t)b.frag_beg].rPos);
		int idx = coor.ChromosomeIdx;
		padded_names(ix, qc, idx, qname, rname);
		trim_extension(ix, coor, r, b, a1, a2);
		const gsa_frag &first = r.frags[(size_t)b.frag_beg], &last = r.frags[(size_t)(b.frag_beg + b.n_frags - 1)];
		if (coor.bDir) {
			fprintf(out, "a score=%d\n", b.bDup ? 1 : b.score);
			fprintf(out, "s ref.%s %d %d + %d %s\n", ix.names[(size_t)idx].c_str(), coor.gPos - 1, b.aln_len - count_gaps(a1.data(), 0, b.aln_len), ix.len[(size_t)idx], a1.data());
			fprintf(out, "s qry.%s %d %d + %d %s\n\n", qname.c_str(), first.qPos, b.aln_len - count_gaps(a2.data(), 0, b.aln_len), (uint32_t)qc.seq.length(), a2.data());
		} else {
			int64_t rpos = last.rPos + last.rLen - 1;
			self_complementary((size_t)b.aln_len, a1.data()); self_complementary((size_t)b.aln_len, a2.data());
			fprintf(out, "a score=%d\n", b.bDup ? 1 : b.score);
			fprintf(out, "s ref.%s %d %d + %d %s\n", ix.names[(size_t)idx].c_str(), gen_coordinate(ix, rpos).gPos - 1, b.aln_len - count_gaps(a1.data(), 0, b.aln_len), ix.len[(size_t)idx], a1.data());
			fprintf(out, "s qry.%s %d %d - %d %s\n\n", qname.c_str(), (uint32_t)qc.seq.length() - (last.qPos + last.qLen), b.aln_len - count_gaps(a2.data(), 0, b.aln_len), (uint32_t)qc.seq.length(), a2.data());
		}
	}
	fclose(out);
}

void output_aln(const Options &o, const HostIndex &ix, const std::vector<QueryChr> &q, int qidx, ContigResult &r)
{ // OutputAlignment, src/tools.cpp:222-286
	FILE *out = fopen(o.aln.c_str(), qidx == 0 ? "w" : "a");
	if (!out) return;
	const QueryChr &qc = q[(size_t)qidx];
	std::vector<char> a1, a2;
	std::string qname, rname;
	for (gsa_block &b : r.blocks) {
		if (!o.allow_dup && b.bDup) continue;
		build_rows(qc, r, b, a1, a2);
		uint32_t aln_len = (uint32_t)b.aln_len; // rows were assembled at the untrimmed length
		Coordinate coor = gen_coordinate(ix, r.frags[(size_t)b.frag_beg].rPos);
		padded_names(ix, qc, coor.ChromosomeIdx, qname, rname);
		trim_extension(ix, coor, r, b, a1, a2);
		fprintf(out, "#Identity = %d / %d (%.2f%%) Orientation = %s\n\n", b.score, b.aln_len, (int)(1000 * (1.0 * b.score / b.aln_len)) / 10.0, coor.bDir ? "Forward" : "Reverse");
		uint32_t pos = 0; int qpos = r.frags[(size_t)b.frag_beg].qPos + 1; long long rpos = coor.gPos;
		while (pos < aln_len) {
			int stop = (int)(pos + 80 > aln_len ? aln_len : pos + 80);
			int p = 80 - count_gaps(a1.data(), (int)pos, stop), qq = 80 - count_gaps(a2.data(), (int)pos, stop);
			fprintf(out, "ref.%s\t%12lld\t%.80s\nqry.%s\t%12d\t%.80s\n\n", rname.c_str(), rpos, a1.data() + pos, qname.c_str(), qpos, a2.data() + pos);
			pos += 80; rpos += coor.bDir ? p : -p; qpos += qq;
		}
		fprintf(out, "%s\n", std::string(100, '*').c_str());
	}
	fclose(out);
}

void variant_identification(const HostIndex &ix, const std::vector<QueryChr> &q, int qidx, const ContigResult &r, EmitState &st)
{
	const std::string &seq = q[(size_t)qidx].seq;
	Variant v;
	for (const gsa_block &b : r.blocks) {
		if (b.bDup) continue;
		v.chr_idx = gen_coordinate(ix, r.frags[(size_t)b.frag_beg].rPos).ChromosomeIdx;
		v.query_idx = qidx;
		for (int64_t t = b.frag_beg; t < b.frag_beg + b.n_frags; t++) {
			const gsa_frag &f = r.frags[(size_t)t];
			if (f.bSeed) continue;
			if (f.qLen == 0 && f.rLen == 0) continue;
			if (f.qLen == 0) { // delete
				st.iDeletion++;
				v.type = 2; v.pos = gen_coordinate(ix, f.rPos - 1).gPos;
				v.ref_frag.resize((size_t)f.rLen + 1);
				for (int k = 0; k <= f.rLen; k++) v.ref_frag[(size_t)k] = ix.text(f.rPos - 1 + k);
				v.alt_frag.assign(1, seq[(size_t)(f.qPos - 1)]);
				st.variants.push_back(v);
			} else if (f.rLen == 0) { // insert
				st.iInsertion++;
				v.type = 1; v.pos = gen_coordinate(ix, f.rPos - 1).gPos;
				v.ref_frag.assign(1, ix.text(f.rPos - 1));
				v.alt_frag = seq.substr((size_t)(f.qPos - 1), (size_t)f.qLen + 1);
				st.variants.push_back(v);
			} else if (f.qLen == 1 && f.rLen == 1) { // substitution
				char c1 = r.aln1[(size_t)f.aln_off], c2 = r.aln2[(size_t)f.aln_off];
				if (nt4(c1) != nt4(c2) && nt4(c2) != 4) {
					st.iSNV++;
					v.type = 0; v.pos = gen_coordinate(ix, f.rPos).gPos;
					v.ref_frag.assign(1, c1); v.alt_frag.assign(1, c2);
					st.variants.push_back(v);
				}
			} else {
				const char *a1 = r.aln1.data() + f.aln_off, *a2 = r.aln2.data() + f.aln_off;
				int aln_len = f.aln_len, qpos = f.qPos; int64_t rpos = f.rPos;
				for (int i = 0; i < aln_len; i++) {
					if (a1[i] == '-') { // insert: REF is the QUERY base before the insertion (hazard H6)
						st.iInsertion++;
						int ind = 1; while (i + ind < aln_len && a1[i + ind] == '-') ind++;
						std::string frag2 = seq.substr((size_t)(qpos - 1), (size_t)ind + 1);
						v.type = 1; v.pos = gen_coordinate(ix, rpos - 1).gPos;
						v.ref_frag.assign(1, frag2[0]); v.alt_frag = frag2;
						st.variants.push_back(v);
						qpos += ind; i += ind - 1;
					} else if (a2[i] == '-') { // delete
						st.iDeletion++;
						int ind = 1; while (i + ind < aln_len && a2[i + ind] == '-') ind++;
						v.type = 2; v.pos = gen_coordinate(ix, rpos - 1).gPos;
						v.ref_frag.resize((size_t)ind + 1);
						for (int k = 0; k <= ind; k++) v.ref_frag[(size_t)k] = ix.text(rpos - 1 + k);
						v.alt_frag.assign(1, v.ref_frag[0]);
						st.variants.push_back(v);
						rpos += ind; i += ind - 1;
					} else if (nt4(a1[i]) != nt4(a2[i])) {
						if (nt4(a2[i]) != 4) {
							st.iSNV++;
							v.type = 0; v.pos = gen_coordinate(ix, rpos).gPos;
							v.ref_frag.assign(1, a1[i]); v.alt_frag.assign(1, a2[i]);
							st.variants.push_back(v);
						}
						rpos++; qpos++;
					} else { rpos++; qpos++; }
				}
			}
		}
	}
}

static bool by_variant_pos(const Variant &a, const Variant &b)
{ // CompByVariantPos, src/SeqVariant.cpp:6-10
	if (a.chr_idx == b.chr_idx) return a.pos < b.pos;
	return a.chr_idx < b.chr_idx;
}

void output_variants(const Options &o, const HostIndex &ix, EmitState &st)
{
	static const char *MutType[3] = {"SUBSTITUTE", "INSERT", "DELETE"};
	std::sort(st.variants.begin(), st.variants.end(), by_variant_pos); // same unstable libstdc++ sort as the reference (H5)
	st.iSNV = st.iInsertion = st.iDeletion = 0;
	FILE *out = fopen(o.vcf_name.c_str(), "w");
	if (!out) return;
	fprintf(out, "##fileformat=VCFv4.1\n");
	fprintf(out, "##reference=%s\n", o.index_prefix ? o.index_prefix : o.ref_fa);
	fprintf(out, "##source=GSAlign %s\n", VERSION_STR);
	fprintf(out, "##INFO=<ID=TYPE,Number=1,Type=String,Description=\"The type of allele, either SUBSTITUTE, INSERT, or DELETE.\">\n");
	for (size_t i = 0; i < ix.names.size(); i++) fprintf(out, "##contig=<ID=%s,length=%d>\n", ix.names[i].c_str(), ix.len[i]);
	fprintf(out, "#CHROM	POS	ID	REF	ALT	QUAL	FILTER	INFO\n");
	for (const Variant &v : st.variants)
		fprintf(out, "%s\t%d\t.\t%s\t%s\t100\t*\tTYPE=%s\n", ix.names[(size_t)v.chr_idx].c_str(), v.pos, v.ref_frag.c_str(), v.alt_frag.c_str(), MutType[v.type]);
	fclose(out);
}
