// io.cpp -- BWA-format index reader and query FASTA reader of bin/GSAlign.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <string.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <fstream>
#include <thread>
#include "host.h"

// maps a whole file privately (copy-on-write); returns false when it cannot be opened or is empty
static bool map_file(const std::string &path, HostIndex::Mapping &m)
{
	int fd = open(path.c_str(), O_RDONLY);
	if (fd < 0) return false;
	struct stat sb;
	if (fstat(fd, &sb) != 0 || sb.st_size <= 0) { close(fd); return false; }
	void *p = mmap(nullptr, (size_t)sb.st_size, PROT_READ | PROT_WRITE, MAP_PRIVATE, fd, 0);
	close(fd);
	if (p == MAP_FAILED) return false;
	m.p = p; m.n = (size_t)sb.st_size;
	return true;
}

HostIndex::~HostIndex()
{
	for (Mapping &m : maps) if (m.p) munmap(m.p, m.n);
}

// bwt_restore_bwt / bwt_restore_sa / bns_restore_core (reference src/bwt_index.cpp:15-121)
bool HostIndex::load(const std::string &prefix, std::string &err)
{
	for (Mapping &m : maps) if (m.p) { munmap(m.p, m.n); m = Mapping(); }
	if (!map_file(prefix + ".bwt", maps[0]) || maps[0].n < 40) { err = "cannot read " + prefix + ".bwt"; return false; }
	const uint64_t *h = (const uint64_t *)maps[0].p;
	primary = h[0]; L2[0] = 0; for (int i = 1; i < 5; i++) L2[i] = h[i];
	seq_len = L2[4];
	bwt = (const uint32_t *)((const char *)maps[0].p + 40); bwt_size = (maps[0].n - 40) / 4;
	if (!map_file(prefix + ".sa", maps[1]) || maps[1].n < 56) { err = "cannot read " + prefix + ".sa"; return false; }
	h = (const uint64_t *)maps[1].p;
	sa_intv = (int)h[5];                       // hazard H13: the reference reads this u64 into an int
	if (sa_intv <= 0) { err = "bad sa_intv in " + prefix + ".sa"; return false; }
	n_sa = (seq_len + (uint64_t)sa_intv) / (uint64_t)sa_intv;
	if (maps[1].n < 56 + (n_sa - 1) * 8) { err = prefix + ".sa is truncated"; return false; }
	// the samples follow a 7-word header: seen from its last word the file IS the array, once that word is sa[0] = -1
	uint64_t *sa_w = (uint64_t *)((char *)maps[1].p + 48);
	sa_w[0] = (uint64_t)-1;
	sa = sa_w;
	FILE *fp = fopen((prefix + ".ann").c_str(), "r");
	if (!fp) { err = "cannot read " + prefix + ".ann"; return false; }
	long long xx; int n_seqs; unsigned seed;
	if (fscanf(fp, "%lld%d%u", &xx, &n_seqs, &seed) != 3) { fclose(fp); err = "bad .ann header"; return false; }
	l_pac = xx;
	names.clear(); offset.clear(); len.clear();
	char str[10240];
	for (int i = 0; i < n_seqs; i++) {
		unsigned gi; int c, l, namb;
		if (fscanf(fp, "%u%10239s", &gi, str) != 2) { fclose(fp); err = "bad .ann record"; return false; }
		names.push_back(str);
		while ((c = fgetc(fp)) != '\n' && c != EOF) {}
		if (fscanf(fp, "%lld%d%d", &xx, &l, &namb) != 3) { fclose(fp); err = "bad .ann record"; return false; }
		offset.push_back(xx); len.push_back(l);
	}
	fclose(fp);
	if (!map_file(prefix + ".pac", maps[2]) || (int64_t)maps[2].n < l_pac / 4 + 1) { err = "cannot read " + prefix + ".pac"; return false; }
	pac = (const uint8_t *)maps[2].p;
	// RestoreReferenceInfo (reference src/bwt_index.cpp:229-253): locations are cumulative lengths
	chr_loc.clear();
	int64_t total = 0;
	for (int i = 0; i < n_seqs; i++) {
		offset[i] = total; total += len[i];
		chr_loc.push_back(std::make_pair(offset[i] + len[i] - 1, i));
		chr_loc.push_back(std::make_pair(2 * l_pac - total + len[i] - 1, i));
	}
	std::sort(chr_loc.begin(), chr_loc.end());
	if (seq_len != 2 * (uint64_t)l_pac) { err = "index inconsistent: seq_len != 2*l_pac"; return false; }
	return true;
}

void HostIndex::view(gsa_index_view *v) const
{
	v->bwt = bwt; v->bwt_size = bwt_size; v->primary = primary;
	for (int i = 0; i < 5; i++) v->L2[i] = L2[i];
	v->seq_len = seq_len; v->sa = sa; v->n_sa = n_sa; v->sa_intv = sa_intv;
	v->pac = pac; v->l_pac = l_pac; v->n_contigs = (int32_t)names.size();
	v->contig_off = offset.data(); v->contig_len = len.data();
}

const std::pair<int64_t, int> &HostIndex::loc(int64_t pos) const
{
	size_t lo = 0, hi = chr_loc.size();
	while (lo < hi) { size_t m = (lo + hi) / 2; if (chr_loc[m].first < pos) lo = m + 1; else hi = m; }
	return chr_loc[lo < chr_loc.size() ? lo : chr_loc.size() - 1];
}

Coordinate gen_coordinate(const HostIndex &ix, int64_t rPos)
{
	Coordinate c;
	const std::pair<int64_t, int> &it = ix.loc(rPos);
	c.ChromosomeIdx = it.second;
	if (rPos < ix.genome()) { c.bDir = true; c.gPos = (int)(rPos + 1 - ix.offset[it.second]); }
	else { c.bDir = false; c.gPos = (int)(it.first - rPos + 1); }
	return c;
}

std::string trim_chromosome_name(std::string name)
{
	size_t i, n = name.length();
	for (i = 0; i < n; i++) {
		if (name[i] == '|') name[i] = '-';
		else if (name[i] == ' ' || name[i] == '#' || name[i] == ':' || name[i] == '=' || name[i] == '\t') break;
	}
	return name.substr(0, i);
}

bool check_input_file(const char *path)
{
	std::ifstream f(path);
	if (!f.is_open()) return false;
	std::string s;
	std::getline(f, s);
	return !s.empty() && s[0] == '>';
}

// The query FASTA is parsed in PIECES: a record (from its '>' header line to the next one) is cut, at line starts, into
// stretches of a few megabytes, so that one long contig keeps every core busy just like many short ones.  Same acceptance
// rules as LoadQueryFile / CheckQuerySeq (src/main.cpp:66-114): header trimmed by TrimChromosomeName, empty lines skipped, a
// trailing '\r' dropped, every other character must be a letter.
struct FastaPiece {
	const char *beg, *end;   // whole lines of the record's body (the header line belongs to no piece)
	size_t rec;              // index of the record
	size_t len = 0, at = 0;  // sequence letters in the piece; where they go in the record's string
	std::string bad;         // the first offending line
	bool is_bad = false;
};

// pass 0 (copy = nullptr) validates the lines of a piece and counts their letters; pass 1 copies them to `copy`
static size_t walk_piece(FastaPiece &pc, char *copy)
{
	size_t at = 0;
	for (const char *p = pc.beg; p < pc.end;) {
		const char *nl = (const char *)memchr(p, '\n', (size_t)(pc.end - p));
		const char *le = nl ? nl : pc.end;
		size_t len = (size_t)(le - p);
		if (len > 0) {
			if (p[len - 1] == '\r') len--;
			if (!copy) {
				unsigned bad = 0;
				for (size_t i = 0; i < len; i++) bad |= (unsigned)((unsigned char)((p[i] | 0x20) - 'a') >= 26); // !isalpha, C locale
				if (bad) { pc.bad.assign(p, len); pc.is_bad = true; return at; }
			} else memcpy(copy + at, p, len);
			at += len;
		}
		p = nl ? nl + 1 : pc.end;
	}
	return at;
}

// gives the string its final length.  Built as C++23 the letters are left for the copying threads to write (and their pages to
// first touch): a 125 Mbp contig is 125 MB that resize() would zero on one thread first.
static void size_string(std::string &s, size_t n)
{
#if defined(__cpp_lib_string_resize_and_overwrite)
	s.resize_and_overwrite(n, [](char *, size_t k) { return k; });
#else
	s.resize(n);
#endif
}

static size_t fasta_piece_bytes()
{ // GSA_FASTA_PIECE shrinks the pieces so that tests put their borders everywhere in small files
	static const size_t v = [] { const char *e = getenv("GSA_FASTA_PIECE"); long long x = e ? atoll(e) : 0; return x > 0 ? (size_t)x : (size_t)4 << 20; }();
	return v;
}

// runs fn(i) for i in [0, n) on up to nth threads (dynamic: items are of uneven size)
template <typename F> static void for_each_item(size_t n, size_t nth, F fn)
{
	std::atomic<size_t> next(0);
	auto work = [&] { for (size_t i; (i = next++) < n;) fn(i); };
	std::vector<std::thread> th;
	for (size_t t = 1; t < std::min(nth, n); t++) th.emplace_back(work);
	work();
	for (auto &t : th) t.join();
}

bool load_query_file(const char *path, std::vector<QueryChr> &out)
{ // LoadQueryFile, src/main.cpp:82-114: the file is mapped, cut at the header lines and the records are parsed in parallel
	int fd = open(path, O_RDONLY);
	if (fd < 0) return false;
	struct stat sb;
	if (fstat(fd, &sb) != 0) { close(fd); return false; }
	size_t size = (size_t)sb.st_size;
	const char *buf = size ? (const char *)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0) : "";
	close(fd);
	if (size && buf == (const char *)MAP_FAILED) return false;
	const char *end = buf + size;
	// record starts: every '>' that is the first character of a line
	std::vector<const char *> starts;
	bool ok = true;
	const char *p = buf;
	while (p < end && *p == '\n') p++;                       // leading empty lines are skipped like any other
	if (p < end && *p != '>') {                              // sequence before any header: the reference validates the line, then fails
		const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
		std::string line(p, (size_t)((nl ? nl : end) - p));
		if (!line.empty() && line[line.size() - 1] == '\r') line.resize(line.size() - 1);
		bool alpha = true;
		for (char c : line) alpha = alpha && isalpha((unsigned char)c);
		if (!alpha) { printf("%s\n", line.c_str()); fprintf(stderr, "The query sequence contains non-alphabet characters!\n"); }
		ok = false;
	}
	if (ok && p < end) {
		// every '>' that opens a line from p on.  The scan is the first touch of the mapping: cut into slices it is also what
		// faults the file in on all cores instead of one (3 GB of query took this loop 2 s on one).
		unsigned hw0 = std::thread::hardware_concurrency();
		const size_t span = (size_t)(end - p), nsl = std::max<size_t>(1, std::min<size_t>(hw0 ? hw0 : 4, span >> 22));
		std::vector<std::vector<const char *> > found(nsl);
		auto scan = [&](size_t k) {
			const char *a = p + span * k / nsl, *b = p + span * (k + 1) / nsl;
			for (const char *c = a; c < b;) {
				c = (const char *)memchr(c, '>', (size_t)(b - c));
				if (!c) break;
				if (c == p || c[-1] == '\n') found[k].push_back(c);
				c++;
			}
		};
		std::vector<std::thread> sc;
		for (size_t k = 1; k < nsl; k++) sc.emplace_back(scan, k);
		scan(0);
		for (auto &t : sc) t.join();
		for (auto &f : found) starts.insert(starts.end(), f.begin(), f.end());
	}
	if (ok) {
		out.resize(starts.size());
		unsigned hw = std::thread::hardware_concurrency();
		const size_t nth = std::max<size_t>(1, hw ? hw : 4);
		// the pieces of every record: the header line is taken here, the body is cut at the first line start at or after every
		// multiple of the piece size
		std::vector<FastaPiece> pieces;
		std::vector<size_t> first_piece(starts.size() + 1, 0);
		const size_t want = fasta_piece_bytes();
		for (size_t i = 0; i < starts.size(); i++) {
			const char *rb = starts[i], *re = i + 1 < starts.size() ? starts[i + 1] : end;
			const char *nl = (const char *)memchr(rb, '\n', (size_t)(re - rb));
			const char *he = nl ? nl : re;
			out[i].name = trim_chromosome_name(std::string(rb + 1, (size_t)(he - rb) - 1));
			first_piece[i] = pieces.size();
			for (const char *b = nl ? nl + 1 : re; b < re;) {
				const char *e = re;
				if ((size_t)(re - b) > want) { const char *c = (const char *)memchr(b + want - 1, '\n', (size_t)(re - (b + want - 1))); if (c) e = c + 1; }
				FastaPiece pc; pc.beg = b; pc.end = e; pc.rec = i;
				pieces.push_back(pc);
				b = e;
			}
		}
		first_piece[starts.size()] = pieces.size();
		for_each_item(pieces.size(), nth, [&](size_t k) { pieces[k].len = walk_piece(pieces[k], nullptr); });
		for (size_t k = 0; k < pieces.size() && ok; k++)
			if (pieces[k].is_bad) { // the first bad line in file order, like the serial reader
				printf("%s\n", pieces[k].bad.c_str());
				fprintf(stderr, "The query sequence contains non-alphabet characters!\n");
				ok = false;
			}
		if (ok) {
			for_each_item(starts.size(), nth, [&](size_t i) { // (the string's pages are first touched by whoever sizes it)
				size_t total = 0;
				for (size_t k = first_piece[i]; k < first_piece[i + 1]; k++) { pieces[k].at = total; total += pieces[k].len; }
				size_string(out[i].seq, total);
			});
			for_each_item(pieces.size(), nth, [&](size_t k) { if (pieces[k].len) walk_piece(pieces[k], &out[pieces[k].rec].seq[pieces[k].at]); });
		}
	}
	if (size) munmap((void *)buf, size);
	if (!ok) return false;
	fprintf(stderr, "\tLoad the query sequences (%d %s)\n", (int)out.size(), out.size() > 1 ? "chromosomes" : "chromosome");
	return !out.empty();
}
