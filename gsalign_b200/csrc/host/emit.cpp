// emit.cpp -- MAF / ALN / VCF emitters of bin/GSAlign, byte-compatible with the reference
// (src/tools.cpp:3-44,142-286 and src/SeqVariant.cpp:6-143; format hazards H3-H8, H15 of SURVEY.md).
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include "host.h"

// fragments per thread chunk (row assembly, variant scan) and VCF records per formatting batch; GSA_EMIT_CHUNK shrinks
// them so that the tests reach the multi-threaded paths on small inputs
static int64_t emit_chunk()
{
	static const int64_t v = [] { const char *e = getenv("GSA_EMIT_CHUNK"); int64_t x = e ? atoll(e) : 0; return x > 0 ? x : (int64_t)65536; }();
	return v;
}

// GSA_TIMING=1: wall clock of the emitters' inner stages on stderr (development aid, silent otherwise)
static double emit_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static bool emit_timing() { static const bool on = getenv("GSA_TIMING") != nullptr; return on; }

// runs fn(k) for k in [0, n) on up to `threads` host threads (k = chunk index; the caller splits its range)
static void parallel_chunks(int n, int threads, const std::function<void(int)> &fn)
{
	if (n <= 1 || threads <= 1) { for (int k = 0; k < n; k++) fn(k); return; }
	std::vector<std::thread> th;
	for (int k = 1; k < n; k++) th.emplace_back(fn, k);
	fn(0);
	for (auto &t : th) t.join();
}

// ---- output files ---------------------------------------------------------------------------------------------------------
// The alignment file of a human-size pair is 2 bytes per aligned base (6 GB at 3 Gbp); pushing it into the page cache is the
// floor of the whole run.  Rows are assembled by all emitter threads in one of two anonymous buffers that live as long as the
// process (250 MB for a 125 Mbp block, first touched by the assembling threads themselves) and leave with one pwrite per
// buffer on a writer thread, so that the next block is assembled while the previous one is on its way into the page cache.
// Assembling straight into a shared mapping of the file was measured too: every 4 KB page of the mapping faults on its own
// and the faults do not scale with threads -- 2.7-3.3 GB/s whatever the thread count, against 5.9 GB/s for pwrite from memory
// on the same file system.
class OutWriter {
public:
	static const int SLOTS = 2;
	static OutWriter &get() { static OutWriter w; return w; }
	// a buffer of at least n bytes that no write is reading (waits for one); its contents are dead
	int acquire(size_t n)
	{
		std::unique_lock<std::mutex> lk(mu);
		cv.wait(lk, [&] { for (int i = 0; i < SLOTS; i++) if (!busy[i]) return true; return false; });
		int s = -1;
		for (int i = 0; i < SLOTS; i++) if (!busy[i] && (s < 0 || (cap[s] < n && cap[i] > cap[s]))) s = i;
		busy[s] = true;
		lk.unlock();
		if (cap[s] < n) {
			if (mem[s]) munmap(mem[s], cap[s]);
			const size_t want = (n + n / 4 + 4095) & ~(size_t)4095;
			void *m = mmap(nullptr, want, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
			if (m == MAP_FAILED) { mem[s] = nullptr; cap[s] = 0; lk.lock(); busy[s] = false; failed = true; cv.notify_all(); return -1; }
			// (no MADV_HUGEPAGE: measured neutral here -- the two buffers are faulted in once per process by all threads -- and with
			// defrag=madvise a huge-page fault compacts memory synchronously, which on a box whose memory is full of freshly
			// written page cache can stall for longer than all the 4 KB faults together)
			mem[s] = (char *)m; cap[s] = want;
		}
		return s;
	}
	char *data(int s) const { return mem[s]; }
	size_t capacity(int s) const { return cap[s]; }
	// queues the first n bytes of buffer s for offset `off` of the file behind fd (the job keeps its own descriptor)
	void submit(int fd, int s, size_t n, size_t off)
	{
		Job j; j.fd = n ? dup(fd) : -1; j.slot = s; j.n = n; j.off = off;
		std::unique_lock<std::mutex> lk(mu);
		if (n && j.fd < 0) failed = true;
		if (!started) { started = true; worker = std::thread([this] { run(); }); }
		jobs.push_back(j);
		cv.notify_all();
	}
	// every queued byte is in the page cache; false if any write failed since the process started
	bool drain()
	{
		std::unique_lock<std::mutex> lk(mu);
		cv.wait(lk, [&] { return jobs.empty() && !active; });
		return !failed;
	}
	// logical end of an output file this process has been writing (writes may still be queued); -1 = not known
	long long known_end(const std::string &path) { std::unique_lock<std::mutex> lk(mu); auto it = ends.find(path); return it == ends.end() ? -1 : (long long)it->second; }
	void set_end(const std::string &path, size_t end) { std::unique_lock<std::mutex> lk(mu); ends[path] = end; }
	~OutWriter()
	{
		{ std::unique_lock<std::mutex> lk(mu); quit = true; cv.notify_all(); }
		if (worker.joinable()) worker.join();
	}
private:
	struct Job { int fd, slot; size_t n, off; };
	std::mutex mu; std::condition_variable cv;
	std::deque<Job> jobs;
	std::map<std::string, size_t> ends;
	std::thread worker;
	char *mem[SLOTS] = {nullptr, nullptr}; size_t cap[SLOTS] = {0, 0}; bool busy[SLOTS] = {false, false};
	bool started = false, quit = false, active = false, failed = false;
	void run()
	{
		std::unique_lock<std::mutex> lk(mu);
		for (;;) {
			cv.wait(lk, [&] { return quit || !jobs.empty(); });
			if (jobs.empty()) return;   // quit, and nothing left to write
			Job j = jobs.front(); jobs.pop_front();
			active = true;
			lk.unlock();
			bool ok = true;
			const char *p = mem[j.slot];
			for (size_t left = j.n, off = j.off; left > 0;) {
				ssize_t w = pwrite(j.fd, p, left, (off_t)off);
				if (w <= 0) { ok = false; break; }
				p += w; left -= (size_t)w; off += (size_t)w;
			}
			if (j.fd >= 0) close(j.fd);
			lk.lock();
			if (!ok) failed = true;
			busy[j.slot] = false; active = false;
			cv.notify_all();
		}
	}
};

bool emit_drain() { return OutWriter::get().drain(); }

// An output file that grows at its end.  reserve() hands out room in the current buffer (small pieces share one buffer and one
// write); a full buffer, and the last one at close, go to the writer thread.
struct MappedOut {
	int fd = -1; size_t end = 0;               // logical end of the file (bytes handed to the writer included)
	int slot = -1; size_t fill = 0;            // current buffer and the bytes of it in use
	std::string path;
	bool open_file(const char *name, bool truncate)
	{
		OutWriter &w = OutWriter::get();
		path = name;
		long long known = w.known_end(path);    // >= 0: this process has written the file before (writes may still be queued)
		if (truncate && known >= 0) w.drain();  // nothing of an earlier writer may land after the truncation
		if (truncate) known = -1;
		fd = open(name, O_RDWR | O_CREAT | (truncate ? O_TRUNC : 0), 0644);
		if (fd < 0) return false;
		struct stat sb;
		end = known >= 0 ? (size_t)known : (fstat(fd, &sb) == 0 ? (size_t)sb.st_size : 0);
		return true;
	}
	// room for n bytes at the end of the file (valid until the next reserve / close)
	char *reserve(size_t n)
	{
		if (n == 0) return nullptr;
		OutWriter &w = OutWriter::get();
		if (slot >= 0 && fill + n > w.capacity(slot)) release();
		if (slot < 0) { slot = w.acquire(std::max<size_t>(n, (size_t)8 << 20)); fill = 0; if (slot < 0) return nullptr; }
		char *p = w.data(slot) + fill;
		fill += n;
		return p;
	}
	void release() { if (slot >= 0) { OutWriter::get().submit(fd, slot, fill, end); end += fill; slot = -1; fill = 0; } }
	void close_file() { release(); if (fd >= 0) { OutWriter::get().set_end(path, end); close(fd); fd = -1; } }
};

// ---- std::sort, bit for bit, on several threads ---------------------------------------------------------------------------
// The order of records with equal keys in the reference's output is whatever libstdc++'s introsort makes of the input order
// (hazard H5), so the sort must BE that introsort.  Its structure leaves room for threads all the same: after a partition
// the right part is sorted by a recursive call that never touches the left part, and the loop carries on with the left one.
// Running the recursive call on another thread changes no comparison and no move: same partition pivots
// (std::__unguarded_partition_pivot), same depth limit and heap-sort fallback (std::__partial_sort), same final insertion
// sort, all libstdc++'s own.
//
// The partition itself runs on threads too where the range is long (the first levels would otherwise be one thread walking
// tens of millions of records).  std::__unguarded_partition(lo, hi, pivot) swaps the k-th element from the left that is not
// below the pivot (a "left stopper", i_k) with the k-th from the right that is not above it (j_k) for as long as i_k < j_k,
// and returns where its left scan stops after the last swap: min(i_K, j_{K-1}) -- the scan may run into the element the last
// swap brought there.  With L(x) = left stoppers in [lo, x) and R(x) = right stoppers in [x, hi) of the ORIGINAL range that
// position is the largest x with L(x) <= R(x), and K = L(x).  So: count both kinds per chunk, find x from the chunk sums and
// one short scan, list the K left stoppers below x and the K rightmost right stoppers, swap them pairwise.  Same swaps, same
// return value, no other moves.
static size_t sort_par_min()
{
	static const size_t v = [] { const char *e = getenv("GSA_SORT_PAR_MIN"); long long x = e ? atoll(e) : 0; return x > 0 ? (size_t)x : (size_t)1 << 20; }();
	return v;
}

template <typename It, typename Cmp>
static It partition_threads(It lo, It hi, It pivot, Cmp comp, int threads)
{
	const size_t n = (size_t)(hi - lo);
	const int nch = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, n / 16));
	auto chunk_at = [&](int c) { return n * (size_t)c / (size_t)nch; };
	std::vector<size_t> lsum((size_t)nch + 1, 0), rsum((size_t)nch + 1, 0);   // lsum[c]: left stoppers before chunk c; rsum[c]: right stoppers from chunk c on
	parallel_chunks(nch, threads, [&](int c) {
		size_t l = 0, r = 0;
		for (It p = lo + (ptrdiff_t)chunk_at(c), e = lo + (ptrdiff_t)chunk_at(c + 1); p != e; ++p) { l += !comp(p, pivot); r += !comp(pivot, p); }
		lsum[(size_t)c + 1] = l; rsum[(size_t)c] = r;
	});
	for (int c = 0; c < nch; c++) lsum[(size_t)c + 1] += lsum[(size_t)c];
	for (int c = nch - 1; c >= 0; c--) rsum[(size_t)c] += rsum[(size_t)c + 1];
	int cc = 0;                                            // the last chunk whose start still has L <= R
	while (cc + 1 < nch && lsum[(size_t)cc + 1] <= rsum[(size_t)cc + 1]) cc++;
	size_t cut = chunk_at(cc), L = lsum[(size_t)cc], R = rsum[(size_t)cc];
	for (const size_t e = chunk_at(cc + 1); cut < e; cut++) {
		It p = lo + (ptrdiff_t)cut;
		const size_t l2 = L + !comp(p, pivot), r2 = R - !comp(pivot, p);
		if (l2 > r2) break;
		L = l2; R = r2;
	}
	const size_t K = L;
	if (K == 0) return lo + (ptrdiff_t)cut;
	std::vector<uint32_t> I(K), J(K);
	parallel_chunks(nch, threads, [&](int c) {
		const size_t a = chunk_at(c), b = chunk_at(c + 1);
		if (a < cut) { // left stoppers of [a, min(b, cut)) in ascending order
			size_t k = lsum[(size_t)c];
			for (size_t x = a, e = std::min(b, cut); x < e; x++) if (!comp(lo + (ptrdiff_t)x, pivot)) I[k++] = (uint32_t)x;
		}
		if (b > cut) { // right stoppers of [max(a, cut), b) in descending order; only the K rightmost of the range take part
			size_t k = rsum[(size_t)c + 1];
			for (size_t x = b, e = std::max(a, cut); x > e && k < K; x--) if (!comp(pivot, lo + (ptrdiff_t)(x - 1))) J[k++] = (uint32_t)(x - 1);
		}
	});
	const int sch = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, K / 4096));
	parallel_chunks(sch, threads, [&](int c) {
		for (size_t k = K * (size_t)c / (size_t)sch, e = K * (size_t)(c + 1) / (size_t)sch; k < e; k++) std::iter_swap(lo + (ptrdiff_t)I[k], lo + (ptrdiff_t)J[k]);
	});
	return lo + (ptrdiff_t)cut;
}

// `threads` = the host threads this call may keep busy: long ranges are partitioned by all of them, then the budget is split
// between the two parts by their sizes
template <typename It, typename Cmp>
static void introsort_loop_threads(It first, It last, long depth_limit, Cmp comp, int threads)
{
	std::vector<std::thread> kids;
	while (last - first > 16) { // _S_threshold
		if (depth_limit == 0) { std::__partial_sort(first, last, last, comp); break; }
		--depth_limit;
		const size_t n = (size_t)(last - first);
		It cut;
		if (threads > 1 && n >= sort_par_min() && n < 0xFFFFFFFFull) { // std::__unguarded_partition_pivot, its partition on threads
			std::__move_median_to_first(first, first + 1, first + (last - first) / 2, last - 1, comp);
			cut = partition_threads(first + 1, last, first, comp, threads);
		} else cut = std::__unguarded_partition_pivot(first, last, comp);
		if (threads > 1 && last - cut > 4096) {
			int tr = (int)((double)threads * (double)(last - cut) / (double)n + 0.5);
			tr = std::max(1, std::min(threads - 1, tr));
			const long d = depth_limit;
			kids.emplace_back([cut, last, d, comp, tr] { introsort_loop_threads(cut, last, d, comp, tr); });
			threads -= tr;
		} else introsort_loop_threads(cut, last, depth_limit, comp, 1);
		last = cut;
	}
	for (auto &t : kids) t.join();
}

template <typename It, typename Compare>
static void sort_like_std(It first, It last, Compare comp, int threads)
{
	if (first == last) return;
	auto c = __gnu_cxx::__ops::__iter_comp_iter(comp);
	// twice the thread count as the budget: the parts of a partition are rarely even, and a thread whose part is done early
	// has nothing else to take
	introsort_loop_threads(first, last, (long)std::__lg(last - first) * 2, c, threads > 1 ? threads * 2 : 1);
	std::__final_insertion_sort(first, last, c);
}

// test hook (tests/test_emit_cpu.py): sorts (key, index) pairs like output_variants does
void gsa_test_sort_keys(uint64_t *keys, uint32_t *idx, size_t n, int threads)
{
	struct K { uint64_t key; uint32_t idx; };
	std::vector<K> v(n);
	for (size_t i = 0; i < n; i++) { v[i].key = keys[i]; v[i].idx = idx[i]; }
	sort_like_std(v.begin(), v.end(), [](const K &a, const K &b) { return a.key < b.key; }, threads);
	for (size_t i = 0; i < n; i++) { keys[i] = v[i].key; idx[i] = v[i].idx; }
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
static void build_rows(const QueryChr &qc, const ContigResult &r, const gsa_block &b, char *a1, char *a2, int threads, int64_t &gaps1, int64_t &gaps2);
static void build_rows(const QueryChr &qc, const ContigResult &r, const gsa_block &b, std::vector<char> &a1, std::vector<char> &a2, int threads,
                       int64_t &gaps1, int64_t &gaps2)
{
	a1.resize((size_t)b.aln_len + 1); a2.resize((size_t)b.aln_len + 1);
	build_rows(qc, r, b, a1.data(), a2.data(), threads, gaps1, gaps2);
}
// (rows of b.aln_len + 1 bytes each, NUL-terminated; the memory may be uninitialised: every byte is written here)
static void build_rows(const QueryChr &qc, const ContigResult &r, const gsa_block &b, char *a1, char *a2, int threads, int64_t &gaps1, int64_t &gaps2)
{
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
				memcpy(a1 + pos, qc.seq.data() + f.qPos, (size_t)f.qLen);
				memcpy(a2 + pos, qc.seq.data() + f.qPos, (size_t)f.qLen);
				pos += (size_t)f.qLen;
			} else {
				const char *s1 = r.aln1.data() + f.aln_off, *s2 = r.aln2.data() + f.aln_off;
				memcpy(a1 + pos, s1, (size_t)f.aln_len);
				memcpy(a2 + pos, s2, (size_t)f.aln_len);
				for (int i = 0; i < f.aln_len; i++) { c1 += s1[i] == '-'; c2 += s2[i] == '-'; }
				pos += (size_t)f.aln_len;
			}
		}
		g1[(size_t)k] = c1; g2[(size_t)k] = c2;
	});
	gaps1 = gaps2 = 0;
	for (int k = 0; k < nch; k++) { gaps1 += g1[(size_t)k]; gaps2 += g2[(size_t)k]; }
}

// iExtension (src/tools.cpp:192-202): a block whose last seed runs past the end of its contig is trimmed in place
template <typename Rows> static void trim_extension(const HostIndex &ix, const Coordinate &coor, ContigResult &r, gsa_block &b, Rows &a1, Rows &a2)
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

// number of '-' characters in either row of a block: only gap fragments hold any
static void count_block_gaps(const ContigResult &r, const gsa_block &b, int threads, int64_t &gaps1, int64_t &gaps2)
{
	const int64_t nf = b.n_frags;
	int nch = (int)std::max<int64_t>(1, std::min<int64_t>(threads, nf / emit_chunk()));
	std::vector<int64_t> g1((size_t)nch, 0), g2((size_t)nch, 0);
	parallel_chunks(nch, threads, [&](int k) {
		int64_t c1 = 0, c2 = 0;
		for (int64_t t = b.frag_beg + nf * k / nch; t < b.frag_beg + nf * (k + 1) / nch; t++) {
			const gsa_frag &f = r.frags[(size_t)t];
			if (f.bSeed) continue;
			const char *s1 = r.aln1.data() + f.aln_off, *s2 = r.aln2.data() + f.aln_off;
			for (int i = 0; i < f.aln_len; i++) { c1 += s1[i] == '-'; c2 += s2[i] == '-'; }
		}
		g1[(size_t)k] = c1; g2[(size_t)k] = c2;
	});
	gaps1 = gaps2 = 0;
	for (int k = 0; k < nch; k++) { gaps1 += g1[(size_t)k]; gaps2 += g2[(size_t)k]; }
}

// The two rows of a block assembled straight into their place in the output (src/tools.cpp:169-184: inside seeds BOTH rows
// are copied from the query), forward or -- for a block on the reverse strand -- already reverse-complemented
// (SelfComplementarySeq + ReverseMap, src/tools.cpp:3-44): fragment t then lands mirrored at the other end of the row.
// Fragments are independent once their offsets are known, so the threads take a share each.
static void assemble_rows(const QueryChr &qc, const ContigResult &r, const gsa_block &b, char *d1, char *d2, bool reverse, int threads)
{
	const int64_t nf = b.n_frags;
	const size_t alen = (size_t)b.aln_len;
	int nch = (int)std::max<int64_t>(1, std::min<int64_t>(threads, nf / emit_chunk()));
	std::vector<size_t> start((size_t)nch + 1, 0);
	auto frag_len = [&](int64_t t) { const gsa_frag &f = r.frags[(size_t)t]; return (size_t)(f.bSeed ? f.qLen : f.aln_len); };
	auto chunk_beg = [&](int k) { return b.frag_beg + nf * k / nch; };
	parallel_chunks(nch, threads, [&](int k) { size_t n = 0; for (int64_t t = chunk_beg(k); t < chunk_beg(k + 1); t++) n += frag_len(t); start[(size_t)k + 1] = n; });
	for (int k = 0; k < nch; k++) start[(size_t)k + 1] += start[(size_t)k];
	parallel_chunks(nch, threads, [&](int k) {
		size_t pos = start[(size_t)k];
		for (int64_t t = chunk_beg(k); t < chunk_beg(k + 1); t++) {
			const gsa_frag &f = r.frags[(size_t)t];
			const size_t L = frag_len(t);
			const char *s1 = f.bSeed ? qc.seq.data() + f.qPos : r.aln1.data() + f.aln_off;
			const char *s2 = f.bSeed ? qc.seq.data() + f.qPos : r.aln2.data() + f.aln_off;
			if (!reverse) { memcpy(d1 + pos, s1, L); memcpy(d2 + pos, s2, L); }
			else {
				char *e1 = d1 + (alen - pos - 1), *e2 = d2 + (alen - pos - 1);
				for (size_t i = 0; i < L; i++) { e1[-(ptrdiff_t)i] = reverse_map(s1[i]); e2[-(ptrdiff_t)i] = reverse_map(s2[i]); }
			}
			pos += L;
		}
	});
}

// every letter of seq[beg, end) survives ReverseMap (anything else becomes NUL there and cuts the printed row short)
static bool reverse_safe(const std::string &seq, size_t beg, size_t end, int threads)
{
	const size_t n = end - beg, piece = (size_t)4 << 20;
	int nch = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, n / piece));
	std::vector<int> bad((size_t)nch, 0);
	parallel_chunks(nch, threads, [&](int k) {
		int x = 0;
		for (size_t i = beg + n * (size_t)k / (size_t)nch, e = beg + n * (size_t)(k + 1) / (size_t)nch; i < e; i++) x |= reverse_map(seq[i]) == '\0';
		bad[(size_t)k] = x;
	});
	for (int x : bad) if (x) return false;
	return true;
}

static void put_bytes(MappedOut &out, const char *p, size_t n) { char *d = out.reserve(n); if (d) memcpy(d, p, n); }

void output_maf(const Options &o, const HostIndex &ix, const std::vector<QueryChr> &q, int qidx, ContigResult &r)
{ // OutputMAF, src/tools.cpp:149-220; the file is opened "w" by the first contig and "a" by the others (H15)
	MappedOut out;
	if (!out.open_file(o.maf.c_str(), qidx == 0)) return;
	if (qidx == 0) put_bytes(out, "##maf version=1\n", 16);
	const QueryChr &qc = q[(size_t)qidx];
	std::vector<char> a1, a2;
	std::string qname, rname;
	char h1[1024], h2[1024];
	for (gsa_block &b : r.blocks) {
		if (!o.allow_dup && b.bDup) continue;
		int64_t gaps1 = 0, gaps2 = 0;
		count_block_gaps(r, b, o.threads, gaps1, gaps2);
		Coordinate coor = gen_coordinate(ix, r.frags[(size_t)b.frag_beg].rPos);
		int idx = coor.ChromosomeIdx;
		padded_names(ix, qc, idx, qname, rname);
		gsa_frag &lastf = r.frags[(size_t)(b.frag_beg + b.n_frags - 1)];
		const int64_t lim = (coor.bDir ? ix.offset[idx] : ix.reverse_location(idx)) + ix.len[idx], fend = lastf.rPos + lastf.rLen;
		const int ext = fend > lim ? (int)(fend - lim) : 0;     // iExtension, src/tools.cpp:192-202
		const bool direct = (ext == 0 || (lastf.bSeed && ext < lastf.qLen)) && qname.size() < 400 && ix.names[(size_t)idx].size() < 400 &&
		                    (coor.bDir || reverse_safe(qc.seq, (size_t)r.frags[(size_t)b.frag_beg].qPos, (size_t)(lastf.qPos + lastf.qLen), o.threads));
		if (direct) { // rows go straight into the mapped file
			if (ext > 0) { b.aln_len -= ext; b.score -= ext; lastf.rLen -= ext; lastf.qLen -= ext; }
			const gsa_frag &first = r.frags[(size_t)b.frag_beg], &last = lastf;
			int n1, n2;
			if (coor.bDir) {
				n1 = snprintf(h1, sizeof(h1), "a score=%d\ns ref.%s %d %d + %d ", b.bDup ? 1 : b.score, ix.names[(size_t)idx].c_str(), coor.gPos - 1, (int)(b.aln_len - gaps1), ix.len[(size_t)idx]);
				n2 = snprintf(h2, sizeof(h2), "\ns qry.%s %d %d + %d ", qname.c_str(), first.qPos, (int)(b.aln_len - gaps2), (uint32_t)qc.seq.length());
			} else {
				int64_t rpos = last.rPos + last.rLen - 1;
				n1 = snprintf(h1, sizeof(h1), "a score=%d\ns ref.%s %d %d + %d ", b.bDup ? 1 : b.score, ix.names[(size_t)idx].c_str(), gen_coordinate(ix, rpos).gPos - 1, (int)(b.aln_len - gaps1), ix.len[(size_t)idx]);
				n2 = snprintf(h2, sizeof(h2), "\ns qry.%s %d %d - %d ", qname.c_str(), (uint32_t)qc.seq.length() - (last.qPos + last.qLen), (int)(b.aln_len - gaps2), (uint32_t)qc.seq.length());
			}
			const size_t alen = (size_t)b.aln_len;
			char *p = out.reserve((size_t)n1 + alen + (size_t)n2 + alen + 2);
			if (!p) break;
			memcpy(p, h1, (size_t)n1); memcpy(p + n1 + alen, h2, (size_t)n2); memcpy(p + n1 + alen + n2 + alen, "\n\n", 2);
			assemble_rows(qc, r, b, p + n1, p + n1 + alen + n2, !coor.bDir, o.threads);
			continue;
		}
		// the general path: rows in memory, cut at the first NUL like fprintf("%s") does
		build_rows(qc, r, b, a1, a2, o.threads, gaps1, gaps2);
		trim_extension(ix, coor, r, b, a1, a2); // trims inside the last seed, which holds no gap characters
		const gsa_frag &first = r.frags[(size_t)b.frag_beg], &last = r.frags[(size_t)(b.frag_beg + b.n_frags - 1)];
		std::string txt;
		auto row = [&](const std::vector<char> &a) { const void *z = memchr(a.data(), 0, (size_t)b.aln_len); txt.append(a.data(), z ? (size_t)((const char *)z - a.data()) : (size_t)b.aln_len); };
		char num[2048];
		if (coor.bDir) {
			snprintf(num, sizeof(num), "a score=%d\ns ref.%s %d %d + %d ", b.bDup ? 1 : b.score, ix.names[(size_t)idx].c_str(), coor.gPos - 1, (int)(b.aln_len - gaps1), ix.len[(size_t)idx]);
			txt += num; row(a1);
			txt += "\ns qry." + qname; snprintf(num, sizeof(num), " %d %d + %d ", first.qPos, (int)(b.aln_len - gaps2), (uint32_t)qc.seq.length());
			txt += num; row(a2);
		} else {
			int64_t rpos = last.rPos + last.rLen - 1;
			std::thread t2([&] { self_complementary((size_t)b.aln_len, a2.data()); });
			self_complementary((size_t)b.aln_len, a1.data());
			t2.join();
			snprintf(num, sizeof(num), "a score=%d\ns ref.", b.bDup ? 1 : b.score);
			txt += num; txt += ix.names[(size_t)idx];
			snprintf(num, sizeof(num), " %d %d + %d ", gen_coordinate(ix, rpos).gPos - 1, (int)(b.aln_len - gaps1), ix.len[(size_t)idx]);
			txt += num; row(a1);
			txt += "\ns qry." + qname; snprintf(num, sizeof(num), " %d %d - %d ", (uint32_t)qc.seq.length() - (last.qPos + last.qLen), (int)(b.aln_len - gaps2), (uint32_t)qc.seq.length());
			txt += num; row(a2);
		}
		txt += "\n\n";
		put_bytes(out, txt.data(), txt.size());
	}
	out.close_file();
}

// "%12lld" / "%12d" of a number of at most twelve characters: right-aligned in 12 columns
static inline char *put_padded_int(char *p, long long v)
{
	char tmp[24]; int n = 0;
	unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
	do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
	if (v < 0) tmp[n++] = '-';
	for (int k = n; k < 12; k++) *p++ = ' ';
	while (n) *p++ = tmp[--n];
	return p;
}

void output_aln(const Options &o, const HostIndex &ix, const std::vector<QueryChr> &q, int qidx, ContigResult &r)
{ // OutputAlignment, src/tools.cpp:222-286.  The 80-column windows of a block are independent once the running positions at
  // their starts are known (prefix sums of the non-gap counts), so they are formatted by all threads straight into the output
  // buffer: a counting pass, the offsets, a writing pass -- the same bytes the fprintf loop produces, "%.80s" stopping at the NUL
  // that iExtension may have put into a row included.
	MappedOut out;
	if (!out.open_file(o.aln.c_str(), qidx == 0)) return;
	const QueryChr &qc = q[(size_t)qidx];
	struct Rows { // the two rows of a block; grown without initialising: build_rows writes every byte, on all threads
		char *p = nullptr; size_t cap = 0;
		~Rows() { free(p); }
		bool fit(size_t n) { if (n <= cap) return true; free(p); p = (char *)malloc(n); cap = p ? n : 0; return p != nullptr; }
		char &operator[](size_t i) { return p[i]; }
	} a1, a2;
	std::string qname, rname;
	const int nth = std::max(1, o.threads);
	for (gsa_block &b : r.blocks) {
		if (!o.allow_dup && b.bDup) continue;
		int64_t gaps1 = 0, gaps2 = 0;
		if (!a1.fit((size_t)b.aln_len + 1) || !a2.fit((size_t)b.aln_len + 1)) { fprintf(stderr, "out of memory while writing the alignments\n"); break; }
		build_rows(qc, r, b, a1.p, a2.p, o.threads, gaps1, gaps2);
		const uint32_t aln_len = (uint32_t)b.aln_len; // rows were assembled at the untrimmed length
		Coordinate coor = gen_coordinate(ix, r.frags[(size_t)b.frag_beg].rPos);
		padded_names(ix, qc, coor.ChromosomeIdx, qname, rname);
		trim_extension(ix, coor, r, b, a1, a2);
		char head[256];
		const int nh = snprintf(head, sizeof(head), "#Identity = %d / %d (%.2f%%) Orientation = %s\n\n", b.score, b.aln_len, (int)(1000 * (1.0 * b.score / b.aln_len)) / 10.0, coor.bDir ? "Forward" : "Reverse");
		const size_t nw = ((size_t)aln_len + 79) / 80;                       // windows of the fprintf loop (pos < aln_len)
		const int nch = (int)std::max<size_t>(1, std::min<size_t>((size_t)nth, nw / (size_t)std::max<int64_t>(1, emit_chunk() / 16)));   // 4 096 windows per thread at least
		auto w_beg = [&](int k) { return nw * (size_t)k / (size_t)nch; };
		// non-gap counts of a window as the loop adds them: 80 minus the '-' of [pos, stop), for a short last window too
		auto steps = [&](size_t w, int &p, int &qq) {
			const int pos = (int)(w * 80), stop = (int)std::min<size_t>((size_t)aln_len, w * 80 + 80);
			p = 80 - count_gaps(a1.p, pos, stop); qq = 80 - count_gaps(a2.p, pos, stop);
		};
		// The two numbers of a window are printed "%12lld" / "%12d": twelve columns as long as they have at most twelve
		// characters, which positions inside a genome always do (|rpos| < 2^31 + 2^32).  A window's size is then known without
		// its running positions, and one pass counts both the bytes and the steps of a chunk.
		const size_t fixed = 4 + rname.size() + 1 + 12 + 1 + 1 + 4 + qname.size() + 1 + 12 + 1 + 2;   // everything of a window but the two rows
		std::vector<long long> rsum((size_t)nch + 1, 0), qsum((size_t)nch + 1, 0);
		std::vector<size_t> at((size_t)nch + 1, 0);
		parallel_chunks(nch, nth, [&](int k) {
			long long sr = 0, sq = 0; size_t bytes = 0;
			for (size_t w = w_beg(k); w < w_beg(k + 1); w++) {
				int p, qq; steps(w, p, qq); sr += p; sq += qq;
				bytes += fixed + strnlen(a1.p + w * 80, 80) + strnlen(a2.p + w * 80, 80);
			}
			rsum[(size_t)k + 1] = sr; qsum[(size_t)k + 1] = sq; at[(size_t)k + 1] = bytes;
		});
		for (int k = 0; k < nch; k++) { rsum[(size_t)k + 1] += rsum[(size_t)k]; qsum[(size_t)k + 1] += qsum[(size_t)k]; at[(size_t)k + 1] += at[(size_t)k]; }
		const long long rpos0 = coor.gPos, rdir = coor.bDir ? 1 : -1;
		const int qpos0 = r.frags[(size_t)b.frag_beg].qPos + 1;
		char *base = out.reserve((size_t)nh + at[(size_t)nch] + 101);
		if (!base) break;
		memcpy(base, head, (size_t)nh);
		parallel_chunks(nch, nth, [&](int k) {
			long long rpos = rpos0 + rdir * rsum[(size_t)k]; int qpos = qpos0 + (int)qsum[(size_t)k];
			char *p = base + nh + at[(size_t)k];
			for (size_t w = w_beg(k); w < w_beg(k + 1); w++) {
				const size_t pos = w * 80;
				const size_t l1 = strnlen(a1.p + pos, 80), l2 = strnlen(a2.p + pos, 80);
				memcpy(p, "ref.", 4); p += 4; memcpy(p, rname.data(), rname.size()); p += rname.size(); *p++ = '\t';
				p = put_padded_int(p, rpos); *p++ = '\t'; memcpy(p, a1.p + pos, l1); p += l1; *p++ = '\n';
				memcpy(p, "qry.", 4); p += 4; memcpy(p, qname.data(), qname.size()); p += qname.size(); *p++ = '\t';
				p = put_padded_int(p, qpos); *p++ = '\t'; memcpy(p, a2.p + pos, l2); p += l2; *p++ = '\n'; *p++ = '\n';
				int pp, qq; steps(w, pp, qq);
				rpos += rdir * pp; qpos += qq;
			}
		});
		memset(base + nh + at[(size_t)nch], '*', 100);
		base[(size_t)nh + at[(size_t)nch] + 100] = '\n';
	}
	out.close_file();
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

void ContigResult::assign_variants(const gsa_variant_list &v)
{
	vars.assign(v.variants, v.variants + v.n_variants);
	var_first.assign(v.block_first, v.block_first + blocks.size());
	var_count.assign(v.block_count, v.block_count + blocks.size());
	have_vars = true;
}

// Records [v_beg, v_end) of the device's variant list (gsa_variants) -> the reference's Variant_t fields: the alleles are
// substrings of the query and of the reference text at the record's coordinates (include/gsalign_b200.h, gsa_variant_kind).
// The allele lengths follow from the record alone, so a first pass sizes a stretch of records (WRITE = false) and a second
// one writes records and alleles straight into their final places: `out` = the stretch's slots in EmitState::variants,
// `pool` = the whole allele array, `at` = where this stretch's alleles start in it.
template <bool WRITE>
static size_t records_to_variants(const HostIndex &ix, const std::string &seq, const ContigResult &r, int chr_idx, int64_t v_beg, int64_t v_end,
                                  Variant *out, char *pool, size_t at, VarCounts &cnt)
{
	for (int64_t k = v_beg; k < v_end; k++) {
		const gsa_variant &d = r.vars[(size_t)k];
		const bool text_ref = d.kind != GSA_VAR_INS;          // REF comes from the reference text (an in-fragment insertion takes the query base, H6)
		const bool long_ref = d.kind == GSA_VAR_DEL || d.kind == GSA_VAR_FRAG_DEL;
		const bool long_alt = d.kind == GSA_VAR_INS || d.kind == GSA_VAR_FRAG_INS;
		const uint32_t ref_len = long_ref ? (uint32_t)d.len + 1u : 1u;
		const uint32_t alt_len = long_alt ? (uint32_t)std::min((size_t)d.len + 1, seq.size() - (size_t)d.qPos) : 1u;   // substr clamps at the end of the string
		if (WRITE) {
			Variant &v = *out++;
			v.chr_idx = chr_idx; v.pos = d.gPos; v.off = at; v.ref_len = ref_len; v.alt_len = alt_len;
			char *ref = pool + at, *alt = ref + ref_len;
			if (text_ref) for (uint32_t i = 0; i < ref_len; i++) ref[i] = ix.text(d.rPos + i);
			else ref[0] = seq[(size_t)d.qPos];
			if (long_alt) memcpy(alt, seq.data() + d.qPos, alt_len);
			else alt[0] = d.kind == GSA_VAR_DEL ? ref[0] : seq[(size_t)d.qPos];
			if (d.kind == GSA_VAR_SNV) { v.type = 0; cnt.snv++; }
			else if (long_alt) { v.type = 1; cnt.ins++; }
			else { v.type = 2; cnt.del++; }
		}
		at += (size_t)ref_len + alt_len;
	}
	return at;
}

void variant_identification(const HostIndex &ix, const std::vector<QueryChr> &q, int qidx, const ContigResult &r, EmitState &st)
{
	const std::string &seq = q[(size_t)qidx].seq;
	for (size_t bi = 0; bi < r.blocks.size(); bi++) {
		const gsa_block &b = r.blocks[bi];
		if (b.bDup) continue;
		const int chr_idx = gen_coordinate(ix, r.frags[(size_t)b.frag_beg].rPos).ChromosomeIdx;
		if (r.have_vars) { // the device found the records: the host only fetches the alleles, every thread into its own stretch
			const int64_t v0 = r.var_first[bi], nv = r.var_count[bi];
			if (nv <= 0) continue;
			const int nch = (int)std::max<int64_t>(1, std::min<int64_t>(st.threads, nv / emit_chunk()));
			std::vector<size_t> at((size_t)nch + 1, 0);
			std::vector<VarCounts> cnt((size_t)nch);
			auto lo = [&](int k) { return v0 + nv * k / nch; };
			parallel_chunks(nch, st.threads, [&](int k) { VarCounts none; at[(size_t)k + 1] = records_to_variants<false>(ix, seq, r, chr_idx, lo(k), lo(k + 1), nullptr, nullptr, 0, none); });
			for (int k = 0; k < nch; k++) at[(size_t)k + 1] += at[(size_t)k];
			const size_t abase = st.alleles.size();
			Variant *vout = st.variants.grow((size_t)nv);
			if (!vout || !st.alleles.grow(at[(size_t)nch])) { // (the arrays stay consistent: the block's records are dropped)
				if (vout) st.variants.n -= (size_t)nv;
				fprintf(stderr, "out of memory while collecting the variants\n");
				return;
			}
			char *pool = &st.alleles[0];                       // offsets count from the start of the allele array
			parallel_chunks(nch, st.threads, [&](int k) {
				VarCounts mine;                                 // (not cnt[k] itself: neighbouring counters share a cache line)
				records_to_variants<true>(ix, seq, r, chr_idx, lo(k), lo(k + 1), vout + (lo(k) - v0), pool, abase + at[(size_t)k], mine);
				cnt[(size_t)k] = mine;
			});
			for (int k = 0; k < nch; k++) { st.iSNV += cnt[(size_t)k].snv; st.iInsertion += cnt[(size_t)k].ins; st.iDeletion += cnt[(size_t)k].del; }
			continue;
		}
		// fragments are independent: big blocks are scanned by several threads and their records appended in fragment order,
		// i.e. in exactly the order the serial loop pushes them (the order matters: the final sort is unstable)
		const int64_t nf = b.n_frags;
		const int nch = (int)std::max<int64_t>(1, std::min<int64_t>(st.threads, nf / emit_chunk()));
		std::vector<std::vector<Variant> > part((size_t)nch);
		std::vector<std::string> pool((size_t)nch);
		std::vector<VarCounts> cnt((size_t)nch);
		parallel_chunks(nch, st.threads, [&](int k) {
			const int64_t t0 = b.frag_beg + nf * k / nch, t1 = b.frag_beg + nf * (k + 1) / nch;
			// a gap fragment yields at most one record per column and three allele bytes per column: reserved once, the vectors
			// never move while they fill (a move of a large vector is an unmap, and an unmap interrupts every thread)
			size_t cols = 0;
			for (int64_t t = t0; t < t1; t++) if (!r.frags[(size_t)t].bSeed) cols += (size_t)std::max(1, r.frags[(size_t)t].aln_len);
			// (the thread fills objects of its own: the headers of part[k], pool[k], cnt[k] share cache lines with their neighbours',
			// and every push writes the header)
			std::vector<Variant> my_part; std::string my_pool; VarCounts mine;
			my_part.reserve(cols); my_pool.reserve(3 * cols + 16);
			scan_fragments(ix, seq, r, chr_idx, t0, t1, my_part, my_pool, mine);
			part[(size_t)k].swap(my_part); pool[(size_t)k].swap(my_pool); cnt[(size_t)k] = mine;
		});
		// the parts move to their places at the end of the two arrays, every thread its own
		std::vector<size_t> vat((size_t)nch + 1, 0), aat((size_t)nch + 1, 0);
		for (int k = 0; k < nch; k++) { vat[(size_t)k + 1] = vat[(size_t)k] + part[(size_t)k].size(); aat[(size_t)k + 1] = aat[(size_t)k] + pool[(size_t)k].size(); }
		const size_t abase = st.alleles.size();
		Variant *vout = st.variants.grow(vat[(size_t)nch]);
		char *aout = st.alleles.grow(aat[(size_t)nch]);
		if ((!vout && vat[(size_t)nch]) || (!aout && aat[(size_t)nch])) {
			if (vout) st.variants.n -= vat[(size_t)nch];
			fprintf(stderr, "out of memory while collecting the variants\n");
			return;
		}
		parallel_chunks(nch, st.threads, [&](int k) {
			if (!pool[(size_t)k].empty()) memcpy(aout + aat[(size_t)k], pool[(size_t)k].data(), pool[(size_t)k].size());
			Variant *o = vout + vat[(size_t)k];
			const uint64_t base = abase + aat[(size_t)k];
			for (const Variant &v : part[(size_t)k]) { *o = v; o->off += base; o++; }
		});
		for (int k = 0; k < nch; k++) { st.iSNV += cnt[(size_t)k].snv; st.iInsertion += cnt[(size_t)k].ins; st.iDeletion += cnt[(size_t)k].del; }
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
	const double t_a = emit_now();
	// the key array is first touched by the threads that fill it (half a gigabyte for a human-size pair)
	struct KeyArray {
		VarKey *p; size_t n;
		explicit KeyArray(size_t n_) : p((VarKey *)malloc(std::max<size_t>(1, n_) * sizeof(VarKey))), n(n_) {}
		~KeyArray() { free(p); }
		size_t size() const { return n; }
		VarKey *begin() const { return p; }
		VarKey *end() const { return p + n; }
		VarKey &operator[](size_t i) const { return p[i]; }
	} keys(st.variants.size());
	if (!keys.p) { fprintf(stderr, "out of memory while sorting the variants\n"); return; }
	{
		const int nch = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, st.threads), keys.size() / 65536));
		parallel_chunks(nch, st.threads, [&](int k) {
			for (size_t i = keys.size() * (size_t)k / (size_t)nch, e = keys.size() * (size_t)(k + 1) / (size_t)nch; i < e; i++) {
				keys[i].key = ((uint64_t)(uint32_t)st.variants[i].chr_idx << 32) | (uint32_t)st.variants[i].pos; keys[i].idx = (uint32_t)i;
			}
		});
	}
	const double t_b = emit_now();
	sort_like_std(keys.begin(), keys.end(), by_variant_pos, st.threads); // std::sort's own moves, its recursive calls on threads
	const double t_c = emit_now();
	double t_fmt = 0, t_wr = 0;
	st.iSNV = st.iInsertion = st.iDeletion = 0;
	MappedOut out;
	if (!out.open_file(o.vcf_name.c_str(), true)) return;
	{
		std::string hdr = "##fileformat=VCFv4.1\n";
		hdr += std::string("##reference=") + (o.index_prefix ? o.index_prefix : o.ref_fa) + "\n";
		hdr += std::string("##source=GSAlign ") + VERSION_STR + "\n";
		hdr += "##INFO=<ID=TYPE,Number=1,Type=String,Description=\"The type of allele, either SUBSTITUTE, INSERT, or DELETE.\">\n";
		for (size_t i = 0; i < ix.names.size(); i++) hdr += "##contig=<ID=" + ix.names[i] + ",length=" + std::to_string(ix.len[i]) + ">\n";
		hdr += "#CHROM	POS	ID	REF	ALT	QUAL	FILTER	INFO\n";
		put_bytes(out, hdr.data(), hdr.size());
	}
	// Records are formatted straight into the process-wide output buffer (already resident after the alignment file), a
	// round of batches at a time: every thread first adds up the bytes of its batch, the batches' offsets follow, then every
	// thread writes its lines where they belong and the round leaves with one pwrite.  "%s" of an allele stops at a NUL, which
	// an allele cannot hold (query letters are alphabetic, reference letters ACGT), so lengths can be used as they are.
	static const size_t MutLen[3] = {10, 6, 6};
	const size_t batch = (size_t)emit_chunk() * 16;
	const int nth = std::max(1, st.threads);
	std::vector<size_t> name_len(ix.names.size());
	for (size_t i = 0; i < ix.names.size(); i++) name_len[i] = ix.names[i].size();
	for (size_t b0 = 0; b0 < keys.size(); b0 += batch * (size_t)nth) {
		const int nch = (int)std::min<size_t>((size_t)nth, (keys.size() - b0 + batch - 1) / batch);
		const double t_d = emit_now();
		std::vector<size_t> at((size_t)nch + 1, 0);
		parallel_chunks(nch, nth, [&](int k) {
			const size_t lo = b0 + (size_t)k * batch, hi = std::min(keys.size(), lo + batch);
			size_t n = 0;
			for (size_t i = lo; i < hi; i++) {
				const Variant &v = st.variants[keys[i].idx];
				unsigned u = v.pos < 0 ? 0u - (unsigned)v.pos : (unsigned)v.pos;
				size_t digits = 1; while (u >= 10) { u /= 10; digits++; }
				// name \t pos \t . \t ref \t alt \t 100 \t * \t TYPE= type \n
				n += name_len[(size_t)v.chr_idx] + 1 + digits + (v.pos < 0) + 3 + v.ref_len + 1 + v.alt_len + 12 + MutLen[v.type] + 1;
			}
			at[(size_t)k + 1] = n;
		});
		for (int k = 0; k < nch; k++) at[(size_t)k + 1] += at[(size_t)k];
		char *base = out.reserve(at[(size_t)nch]);
		if (!base && at[(size_t)nch]) break;
		parallel_chunks(nch, nth, [&](int k) {
			const size_t lo = b0 + (size_t)k * batch, hi = std::min(keys.size(), lo + batch);
			char *p = base + at[(size_t)k];
			for (size_t i = lo; i < hi; i++) {
				const Variant &v = st.variants[keys[i].idx];
				const std::string &nm = ix.names[(size_t)v.chr_idx];
				memcpy(p, nm.data(), nm.size()); p += nm.size(); *p++ = '\t';
				p = put_int(p, v.pos);
				memcpy(p, "\t.\t", 3); p += 3;
				memcpy(p, st.alleles.data() + v.off, v.ref_len); p += v.ref_len; *p++ = '\t';
				memcpy(p, st.alleles.data() + v.off + v.ref_len, v.alt_len); p += v.alt_len;
				memcpy(p, "\t100\t*\tTYPE=", 12); p += 12;
				memcpy(p, MutType[v.type], MutLen[v.type]); p += MutLen[v.type]; *p++ = '\n';
			}
		});
		const double t_e = emit_now();
		out.release();
		t_fmt += t_e - t_d; t_wr += emit_now() - t_e;
	}
	out.close_file();
	if (emit_timing()) fprintf(stderr, "[timing] variants: keys %.3f s, sort %.3f s, format %.3f s, write %.3f s (%zu records)\n", t_b - t_a, t_c - t_b, t_fmt, t_wr, keys.size());
}
