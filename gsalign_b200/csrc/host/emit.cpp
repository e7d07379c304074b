// emit.cpp -- MAF / ALN / VCF emitters of bin/GSAlign, byte-compatible with the reference
// (src/tools.cpp:3-44,142-286 and src/SeqVariant.cpp:6-143; format hazards H3-H8, H15 of SURVEY.md).
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <thread>
#include "host.h"

// fragments per thread chunk (row assembly, variant scan) and VCF records per formatting batch; GSA_EMIT_CHUNK shrinks
// them so that the tests reach the multi-threaded paths on small inputs
static int64_t emit_chunk()
{
	static const int64_t v = [] { const char *e = getenv("GSA_EMIT_CHUNK"); int64_t x = e ? atoll(e) : 0; return x > 0 ? x : (int64_t)65536; }();
	return v;
}

// runs fn(k) for k in [0, n) on up to `threads` host threads (k = chunk index; the caller splits its range)
static void parallel_chunks(int n, int threads, const std::function<void(int)> &fn)
{
	if (n <= 1 || threads <= 1) { for (int k = 0; k < n; k++) fn(k); return; }
	std::vector<std::thread> th;
	for (int k = 1; k < n; k++) th.emplace_back(fn, k);
	fn(0);
	for (auto &t : th) t.join();
}

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

// assembles the two rows of a block (src/tools.cpp:169-184): inside seeds BOTH rows are copied from the query.
// Fragments are independent once their offsets are known, so big blocks are assembled by several threads; the number of
// gap characters of each row (needed for the "size" column) is counted on the way: only gap fragments hold any.
static void build_rows(const QueryChr &qc, const ContigResult &r, const gsa_block &b, std::vector<char> &a1, std::vector<char> &a2, int threads,
                       int64_t &gaps1, int64_t &gaps2)
{
	a1.resize((size_t)b.aln_len + 1); a2.resize((size_t)b.aln_len + 1);
	a1[(size_t)b.aln_len] = a2[(size_t)b.aln_len] = '\0';
	const int64_t nf = b.n_frags;
	int nch = (int)std::max<int64_t>(1, std::min<int64_t>(threads, nf / emit_chunk()));
	std::vector<size_t> start((size_t)nch + 1, 0);
	std::vector<int64_t> g1((size_t)nch, 0), g2((size_t)nch, 0);
	auto frag_len = [&](int64_t t) { const gsa_frag &f = r.frags[(size_t)t]; return (size_t)(f.bSeed ? f.qLen : f.aln_len); };
	auto chunk_beg = [&](int k) { return b.frag_beg + nf * k / nch; };
	parallel_chunks(nch, threads, [&](int k) { size_t n = 0; for (int64_t t = chunk_beg(k); t < chunk_beg(k + 1); t++) n += frag_len(t); start[(size_t)k + 1] = n; });
	for (int k = 0; k < nch; k++) start[(size_t)k + 1] += start[(size_t)k];
	parallel_chunks(nch, threads, [&](int k) {
		size_t pos = start[(size_t)k];
		int64_t c1 = 0, c2 = 0;
		for (int64_t t = chunk_beg(k); t < chunk_beg(k + 1); t++) {
			const gsa_frag &f = r.frags[(size_t)t];
			if (f.bSeed) {
				memcpy(a1.data() + pos, qc.seq.data() + f.qPos, (size_t)f.qLen);
				memcpy(a2.data() + pos, qc.seq.data() + f.qPos, (size_t)f.qLen);
				pos += (size_t)f.qLen;
			} else {
				const char *s1 = r.aln1.data() + f.aln_off, *s2 = r.aln2.data() + f.aln_off;
				memcpy(a1.data() + pos, s1, (size_t)f.aln_len);
				memcpy(a2.data() + pos, s2, (size_t)f.aln_len);
				for (int i = 0; i < f.aln_len; i++) { c1 += s1[i] == '-'; c2 += s2[i] == '-'; }
				pos += (size_t)f.aln_len;
			}
		}
		g1[(size_t)k] = c1; g2[(size_t)k] = c2;
	});
	gaps1 = gaps2 = 0;
	for (int k = 0; k < nch; k++) { gaps1 += g1[(size_t)k]; gaps2 += g2[(size_t)k]; }
}

// what fprintf("%s") would print of a row: everything up to the first NUL (ReverseMap turns unknown letters into NUL, H3)
static void write_row(FILE *out, const std::vector<char> &a, int aln_len)
{
	const void *z = memchr(a.data(), 0, (size_t)aln_len);
	fwrite(a.data(), 1, z ? (size_t)((const char *)z - a.data()) : (size_t)aln_len, out);
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
		int64_t gaps1 = 0, gaps2 = 0;
		build_rows(qc, r, b, a1, a2, o.threads, gaps1, gaps2);
		Coordinate coor = gen_coordinate(ix, r.frags[(size_t)b.frag_beg].rPos);
		int idx = coor.ChromosomeIdx;
		padded_names(ix, qc, idx, qname, rname);
		trim_extension(ix, coor, r, b, a1, a2); // trims inside the last seed, which holds no gap characters
		const gsa_frag &first = r.frags[(size_t)b.frag_beg], &last = r.frags[(size_t)(b.frag_beg + b.n_frags - 1)];
		if (coor.bDir) {
			fprintf(out, "a score=%d\n", b.bDup ? 1 : b.score);
			fprintf(out, "s ref.%s %d %d + %d ", ix.names[(size_t)idx].c_str(), coor.gPos - 1, (int)(b.aln_len - gaps1), ix.len[(size_t)idx]);
			write_row(out, a1, b.aln_len);
			fprintf(out, "\ns qry.%s %d %d + %d ", qname.c_str(), first.qPos, (int)(b.aln_len - gaps2), (uint32_t)qc.seq.length());
			write_row(out, a2, b.aln_len);
			fputs("\n\n", out);
		} else {
			int64_t rpos = last.rPos + last.rLen - 1;
			std::thread t2([&] { self_complementary((size_t)b.aln_len, a2.data()); });
			self_complementary((size_t)b.aln_len, a1.data());
			t2.join();
			fprintf(out, "a score=%d\n", b.bDup ? 1 : b.score);
			fprintf(out, "s ref.%s %d %d + %d ", ix.names[(size_t)idx].c_str(), gen_coordinate(ix, rpos).gPos - 1, (int)(b.aln_len - gaps1), ix.len[(size_t)idx]);
			write_row(out, a1, b.aln_len);
			fprintf(out, "\ns qry.%s %d %d - %d ", qname.c_str(), (uint32_t)qc.seq.length() - (last.qPos + last.qLen), (int)(b.aln_len - gaps2), (uint32_t)qc.seq.length());
			write_row(out, a2, b.aln_len);
			fputs("\n\n", out);
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
		int64_t gaps1 = 0, gaps2 = 0;
		build_rows(qc, r, b, a1, a2, o.threads, gaps1, gaps2);
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

// VariantIdentification (src/SeqVariant.cpp:12-119) over fragments [t_beg, t_end) of one block: records go to `out` in
// fragment order, alleles to `pool` (offsets relative to it)
struct VarCounts { int snv = 0, ins = 0, del = 0; };
static void scan_fragments(const HostIndex &ix, const std::string &seq, const ContigResult &r, int chr_idx, int64_t t_beg, int64_t t_end,
                           std::vector<Variant> &out, std::string &pool, VarCounts &cnt)
{
	Variant v; v.chr_idx = chr_idx;
	auto push = [&](int type, int pos, const char *ref, size_t ref_len, const char *alt, size_t alt_len) {
		v.type = type; v.pos = pos; v.ref_len = (uint32_t)ref_len; v.alt_len = (uint32_t)alt_len; v.off = pool.size();
		pool.append(ref, ref_len); pool.append(alt, alt_len);
		out.push_back(v);
	};
	std::string tmp;
	for (int64_t t = t_beg; t < t_end; t++) {
		const gsa_frag &f = r.frags[(size_t)t];
		if (f.bSeed) continue;
		if (f.qLen == 0 && f.rLen == 0) continue;
		if (f.qLen == 0) { // delete
			cnt.del++;
			tmp.resize((size_t)f.rLen + 1);
			for (int k = 0; k <= f.rLen; k++) tmp[(size_t)k] = ix.text(f.rPos - 1 + k);
			push(2, gen_coordinate(ix, f.rPos - 1).gPos, tmp.data(), tmp.size(), seq.data() + (f.qPos - 1), 1);
		} else if (f.rLen == 0) { // insert
			cnt.ins++;
			char rc = ix.text(f.rPos - 1);
			size_t n = std::min((size_t)f.qLen + 1, seq.size() - (size_t)(f.qPos - 1)); // substr clamps at the end of the string
			push(1, gen_coordinate(ix, f.rPos - 1).gPos, &rc, 1, seq.data() + (f.qPos - 1), n);
		} else if (f.qLen == 1 && f.rLen == 1) { // substitution
			char c1 = r.aln1[(size_t)f.aln_off], c2 = r.aln2[(size_t)f.aln_off];
			if (nt4(c1) != nt4(c2) && nt4(c2) != 4) {
				cnt.snv++;
				push(0, gen_coordinate(ix, f.rPos).gPos, &c1, 1, &c2, 1);
			}
		} else {
			const char *a1 = r.aln1.data() + f.aln_off, *a2 = r.aln2.data() + f.aln_off;
			int aln_len = f.aln_len, qpos = f.qPos; int64_t rpos = f.rPos;
			for (int i = 0; i < aln_len; i++) {
				if (a1[i] == '-') { // insert: REF is the QUERY base before the insertion (hazard H6)
					cnt.ins++;
					int ind = 1; while (i + ind < aln_len && a1[i + ind] == '-') ind++;
					size_t n = std::min((size_t)ind + 1, seq.size() - (size_t)(qpos - 1));
					push(1, gen_coordinate(ix, rpos - 1).gPos, seq.data() + (qpos - 1), 1, seq.data() + (qpos - 1), n);
					qpos += ind; i += ind - 1;
				} else if (a2[i] == '-') { // delete
					cnt.del++;
					int ind = 1; while (i + ind < aln_len && a2[i + ind] == '-') ind++;
					tmp.resize((size_t)ind + 1);
					for (int k = 0; k <= ind; k++) tmp[(size_t)k] = ix.text(rpos - 1 + k);
					push(2, gen_coordinate(ix, rpos - 1).gPos, tmp.data(), tmp.size(), tmp.data(), 1);
					rpos += ind; i += ind - 1;
				} else if (nt4(a1[i]) != nt4(a2[i])) {
					if (nt4(a2[i]) != 4) {
						cnt.snv++;
						push(0, gen_coordinate(ix, rpos).gPos, a1 + i, 1, a2 + i, 1);
					}
					rpos++; qpos++;
				} else { rpos++; qpos++; }
			}
		}
	}
}

void variant_identification(const HostIndex &ix, const std::vector<QueryChr> &q, int qidx, const ContigResult &r, EmitState &st)
{
	const std::string &seq = q[(size_t)qidx].seq;
	for (const gsa_block &b : r.blocks) {
		if (b.bDup) continue;
		const int chr_idx = gen_coordinate(ix, r.frags[(size_t)b.frag_beg].rPos).ChromosomeIdx;
		// fragments are independent: big blocks are scanned by several threads and their records appended in fragment order,
		// i.e. in exactly the order the serial loop pushes them (the order matters: the final sort is unstable)
		const int64_t nf = b.n_frags;
		const int nch = (int)std::max<int64_t>(1, std::min<int64_t>(st.threads, nf / emit_chunk()));
		std::vector<std::vector<Variant> > part((size_t)nch);
		std::vector<std::string> pool((size_t)nch);
		std::vector<VarCounts> cnt((size_t)nch);
		parallel_chunks(nch, st.threads, [&](int k) {
			scan_fragments(ix, seq, r, chr_idx, b.frag_beg + nf * k / nch, b.frag_beg + nf * (k + 1) / nch, part[(size_t)k], pool[(size_t)k], cnt[(size_t)k]);
		});
		for (int k = 0; k < nch; k++) {
			const uint64_t base = st.alleles.size();
			st.alleles += pool[(size_t)k];
			size_t at = st.variants.size();
			st.variants.insert(st.variants.end(), part[(size_t)k].begin(), part[(size_t)k].end());
			for (size_t i = at; i < st.variants.size(); i++) st.variants[i].off += base;
			st.iSNV += cnt[(size_t)k].snv; st.iInsertion += cnt[(size_t)k].ins; st.iDeletion += cnt[(size_t)k].del;
		}
	}
}

// The final order is whatever libstdc++'s (unstable) std::sort makes of the push order under CompByVariantPos
// (src/SeqVariant.cpp:6-10,126; hazard H5).  Introsort's moves depend on comparison outcomes only, so sorting 12-byte
// (chr, pos, index) keys under a comparator with the same outcomes yields the same permutation as sorting the records.
struct VarKey { uint64_t key; uint32_t idx; };   // key = chr_idx << 32 | pos (both non-negative): one compare, same outcomes
static bool by_variant_pos(const VarKey &a, const VarKey &b) { return a.key < b.key; }

static inline char *put_int(char *p, int v)
{ // "%d"
	char tmp[12]; int n = 0;
	unsigned u = v < 0 ? 0u - (unsigned)v : (unsigned)v;
	do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
	if (v < 0) *p++ = '-';
	while (n) *p++ = tmp[--n];
	return p;
}

void output_variants(const Options &o, const HostIndex &ix, EmitState &st)
{
	static const char *MutType[3] = {"SUBSTITUTE", "INSERT", "DELETE"};
	if (st.variants.size() >= 0xFFFFFFFFull) { fprintf(stderr, "too many variants for this build\n"); return; }
	std::vector<VarKey> keys(st.variants.size());
	for (size_t i = 0; i < keys.size(); i++) { keys[i].key = ((uint64_t)(uint32_t)st.variants[i].chr_idx << 32) | (uint32_t)st.variants[i].pos; keys[i].idx = (uint32_t)i; }
	std::sort(keys.begin(), keys.end(), by_variant_pos);
	st.iSNV = st.iInsertion = st.iDeletion = 0;
	FILE *out = fopen(o.vcf_name.c_str(), "w");
	if (!out) return;
	fprintf(out, "##fileformat=VCFv4.1\n");
	fprintf(out, "##reference=%s\n", o.index_prefix ? o.index_prefix : o.ref_fa);
	fprintf(out, "##source=GSAlign %s\n", VERSION_STR);
	fprintf(out, "##INFO=<ID=TYPE,Number=1,Type=String,Description=\"The type of allele, either SUBSTITUTE, INSERT, or DELETE.\">\n");
	for (size_t i = 0; i < ix.names.size(); i++) fprintf(out, "##contig=<ID=%s,length=%d>\n", ix.names[i].c_str(), ix.len[i]);
	fprintf(out, "#CHROM	POS	ID	REF	ALT	QUAL	FILTER	INFO\n");
	// records are formatted into per-thread buffers, a batch at a time, and written in order.  "%s" of an allele stops at a
	// NUL, which an allele cannot hold (query letters are alphabetic, reference letters ACGT), so lengths can be used as they are.
	const size_t batch = (size_t)emit_chunk() * 16;
	const int nth = std::max(1, st.threads);
	std::vector<std::string> buf((size_t)nth);
	for (size_t b0 = 0; b0 < keys.size(); b0 += batch * (size_t)nth) {
		const int nch = (int)std::min<size_t>((size_t)nth, (keys.size() - b0 + batch - 1) / batch);
		parallel_chunks(nch, nth, [&](int k) {
			std::string &s = buf[(size_t)k];
			s.clear();
			const size_t lo = b0 + (size_t)k * batch, hi = std::min(keys.size(), lo + batch);
			char num[16];
			for (size_t i = lo; i < hi; i++) {
				const Variant &v = st.variants[keys[i].idx];
				s += ix.names[(size_t)v.chr_idx]; s += '\t';
				s.append(num, (size_t)(put_int(num, v.pos) - num));
				s += "\t.\t";
				s.append(st.alleles, (size_t)v.off, v.ref_len); s += '\t';
				s.append(st.alleles, (size_t)(v.off + v.ref_len), v.alt_len);
				s += "\t100\t*\tTYPE="; s += MutType[v.type]; s += '\n';
			}
		});
		for (int k = 0; k < nch; k++) fwrite(buf[(size_t)k].data(), 1, buf[(size_t)k].size(), out);
	}
	fclose(out);
}
