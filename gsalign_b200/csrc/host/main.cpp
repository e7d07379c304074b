// main.cpp -- bin/GSAlign: the reference's command line (src/main.cpp:14-33,198-334) on top of the
// B200 seed -> cluster -> fill path.  Same flags, same defaults, same files, exit code 0 always.
// The per-contig loop of GenomeComparison (src/GSAlign.cpp:473-552) becomes: gsa_align_contig() on a GPU,
// then the emitters on the host.  Query contigs are independent, so with -gpus N they are dealt to N
// GPUs (longest first); records are emitted in contig order whatever GPU produced them.
#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <algorithm>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include "host.h"

extern "C" int gsa_build_index_files(const char *fasta, const char *prefix, int device); // index_build.cu

static void usage(const char *prog, const Options &o)
{
	fprintf(stderr, "\n");
	fprintf(stderr, "GenAlign v%s (B200 build)\n", "1.0.22");
	fprintf(stderr, "Usage: %s [-i IndexFile Prefix / -r Reference file] -q QueryFile[Fasta]\n\n", prog);
	fprintf(stderr, "Options: -t     INT     number of threads [%d]\n", o.threads);
	fprintf(stderr, "         -o     STR     Set the prefix of the output files [output]\n");
	fprintf(stderr, "         -fmt   INT     Set the output format 1:maf, 2:aln [%d]\n", o.out_format);
	fprintf(stderr, "         -idy   INT     Set the minimal sequence identity (0-100) of a local alignment [%d]\n", o.min_idy);
	fprintf(stderr, "         -slen  INT     Set the minimal seed length [%d]\n", o.min_seed_len);
	fprintf(stderr, "         -alen  INT     Set the minimal alignment length [%d]\n", o.min_aln_len);
	fprintf(stderr, "         -ind   INT     Set the maximal indel size [%d]\n", o.max_indel);
	fprintf(stderr, "         -clr   INT     Set the minimal cluster size [%d]\n", o.min_block_score);
	fprintf(stderr, "         -unique        Output unique alignment only [false]\n");
	fprintf(stderr, "         -sen           Sensitive mode [False]\n");
	fprintf(stderr, "         -dp            Output Dot-plots\n");
	fprintf(stderr, "         -one           set one on one aligment mode[false]\n");
	fprintf(stderr, "         -gp    STR     Specify the path of gnuplot\n");
	fprintf(stderr, "         -gpus  INT     number of B200s to spread the query contigs over [1]\n");
	fprintf(stderr, "         -lanes INT     query contigs in flight per GPU [%d]\n", o.lanes);
	fprintf(stderr, "\n");
}

static bool check_output_prefix(const char *p)
{ // CheckOutputPrefix, src/main.cpp:116-138
	if (strcmp(p, "/dev/null") == 0) return true;
	for (size_t i = 0, n = strlen(p); i < n; i++) {
		int c = (int)p[i];
		if (!isprint((unsigned char)p[i]) || (c >= 32 && c <= 44) || (c >= 58 && c <= 64) || (c >= 123 && c <= 127)) {
			fprintf(stderr, "FatalError: Please specify a valid prefix name\n");
			return false;
		}
	}
	return true;
}

static bool index_files_present(const std::string &prefix)
{ // CheckBWAIndexFiles, src/GetData.cpp:8-24 (.bwt/.sa are not checked there either)
	const char *ext[] = {".ann", ".amb", ".pac"};
	for (const char *e : ext) { FILE *f = fopen((prefix + e).c_str(), "r"); if (!f) return false; fclose(f); }
	return true;
}

// GSA_TIMING=1: wall-clock of the host-side stages on stderr (development aid, silent otherwise)
#include <chrono>
static void tick(const char *what)
{
	static const bool on = getenv("GSA_TIMING") != nullptr;
	static auto t0 = std::chrono::steady_clock::now(), last = t0;
	if (!on) return;
	auto now = std::chrono::steady_clock::now();
	fprintf(stderr, "[timing] %-28s +%.3f s (total %.3f s)\n", what, std::chrono::duration<double>(now - last).count(), std::chrono::duration<double>(now - t0).count());
	last = now;
}

// (lives here, not in emit.cpp: the emitters build without the library -- tests/emit_harness.cpp -- and this calls into it)
int ContigResult::assign_record(const gsa_alignment &a, const void *image, int64_t bytes, int64_t record_offset)
{
	blocks.assign(a.blocks, a.blocks + a.n_blocks);
	frags.resize((size_t)a.n_frags);
	aln1.assign(a.aln1 ? a.aln1 : "", (size_t)a.aln_bytes);
	aln2.assign(a.aln2 ? a.aln2 : "", (size_t)a.aln_bytes);
	return gsa_record_frags(image, bytes, record_offset, frags.data(), 4);
}

struct Worker {
	gsa_ctx *ctx = nullptr;
	std::thread th;
};

int main(int argc, char *argv[])
{
	Options o;
	if (argc == 1 || strcmp(argv[1], "-h") == 0) { usage(argv[0], o); return 0; }
	if (strcmp(argv[1], "update") == 0) { fprintf(stderr, "self-update is not supported by this build\n"); return 0; }
	if (strcmp(argv[1], "index") == 0) {
		if (argc == 4) { if (gsa_build_index_files(argv[2], argv[3], 0) != 0) fprintf(stderr, "index construction failed\n"); }
		else fprintf(stderr, "usage: %s index ref.fa prefix\n", argv[0]);
		return 0;
	}
	for (int i = 1; i < argc; i++) {
		std::string p = argv[i];
		if (p == "-i") o.index_prefix = argv[++i];
		else if (p == "-r" && i + 1 < argc) o.ref_fa = argv[++i];
		else if (p == "-q" && i + 1 < argc) o.query = argv[++i];
		else if (p == "-t" && i + 1 < argc) { if ((o.threads = atoi(argv[++i])) < 0) { fprintf(stderr, "Warning! Thread number should be greater than 0!\n"); o.threads = 16; } }
		else if (p == "-slen" && i + 1 < argc) { o.min_seed_len = atoi(argv[++i]); if (o.min_seed_len < 10 || o.min_seed_len > 30) { fprintf(stderr, "Warning! minimal seed length is between 10~20!\n"); return 0; } }
		else if (p == "-ind" && i + 1 < argc) { o.max_indel = atoi(argv[++i]); if (o.max_indel < 10 || o.max_indel > 100) { fprintf(stderr, "Warning! maximal indel size is between 10~100!\n"); return 0; } }
		else if (p == "-sen" || p == "-sensitive") { o.sensitive = true; o.min_aln_len = 200; o.min_block_score = 50; }
		else if (p == "-unique") o.allow_dup = false;
		else if (p == "-no_vcf") o.vcf = false;
		else if (p == "-one") o.one_on_one = true;
		else if (p == "-idy" && i + 1 < argc) o.min_idy = atoi(argv[++i]);
		else if (p == "-alen" && i + 1 < argc) o.min_aln_len = atoi(argv[++i]);
		else if (p == "-clr" && i + 1 < argc) o.min_block_score = atoi(argv[++i]);
		else if (p == "-dp") o.show_plot = true;
		else if (p == "-gp" && i + 1 < argc) o.gnuplot = argv[++i];
		else if (p == "-fmt" && i + 1 < argc) o.out_format = atoi(argv[++i]);
		else if (p == "-o") o.out_prefix = argv[++i];
		else if (p == "-d" || p == "-debug") o.debug = true;
		else if (p == "-obr") ++i;
		else if (p == "-gpus" && i + 1 < argc) o.n_gpus = std::max(1, atoi(argv[++i]));
		else if (p == "-lanes" && i + 1 < argc) o.lanes = std::max(1, atoi(argv[++i]));
		else fprintf(stderr, "Warning! Unknow parameter: %s\n", argv[i]);
	}
	if ((!o.index_prefix && !o.ref_fa) || !o.query) { usage(argv[0], o); return 0; }
	if (!o.out_prefix) o.out_prefix = "output";
	else if (!check_output_prefix(o.out_prefix)) return 0;

	// every lane drives its own stream plus side streams: ask for the maximum of hardware work queues before CUDA initialises,
	// so that a launch of one lane does not queue behind another lane's bulk copy
	setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
	time_t t_start = time(NULL);
	tick("start");
	fprintf(stderr, "Step1. Load the two genome sequences...\n");
	if (o.sensitive) o.min_seed_len = 10; // src/main.cpp:323
	gsa_params prm; gsa_default_params(&prm);
	prm.min_seed_len = o.min_seed_len; prm.sensitive = o.sensitive; prm.max_indel = o.max_indel; prm.min_block_score = o.min_block_score;
	prm.min_aln_len = o.min_aln_len; prm.min_idy = o.min_idy; prm.one_on_one = o.one_on_one;

	// The reference side (index files -> host -> every GPU's HBM) is prepared by a helper thread while this thread
	// parses the query FASTA; the messages keep the reference's order.
	HostIndex ix;
	std::string idx_err;   // empty = ok; otherwise the message to print
	int n_dev = std::max(1, o.n_gpus);
	std::vector<gsa_ctx *> owners((size_t)n_dev, nullptr);
	std::thread idx_thread([&] {
		std::string err, prefix;
		if (o.index_prefix && index_files_present(o.index_prefix)) prefix = o.index_prefix;
		else if (o.ref_fa && check_input_file(o.ref_fa)) {
			prefix = o.ref_fa;
			size_t p = prefix.find_last_of('.');
			if (p != std::string::npos && p > 0) prefix.resize(p);
			if (gsa_build_index_files(o.ref_fa, prefix.c_str(), 0) != 0) { idx_err = "\n\nError! Please check your input!\n"; return; }
		} else { idx_err = "Please specify a valid reference genome\n"; return; }
		if (!ix.load(prefix, err)) { idx_err = "\n\nError! Please check your input! (" + err + ")\n"; return; }
		tick("index files loaded");
		gsa_index_view view; ix.view(&view);
		// GPU 0 gets the index files and derives the HBM layout; the other GPUs receive a copy of the finished layout over
		// NVLink (the derived structures are several times the size of the files), all of them at the same time
		{ // a CUDA context takes a second or two to come up: all GPUs at once
			std::vector<std::thread> mk;
			std::vector<int> mk_rc((size_t)n_dev, 0);
			for (int g = 0; g < n_dev; g++) mk.emplace_back([&, g] { mk_rc[(size_t)g] = gsa_create(g, &owners[g]); });
			for (auto &t : mk) t.join();
			for (int g = 0; g < n_dev; g++)
				if (mk_rc[(size_t)g] != 0) { idx_err = "FatalError: cannot open CUDA device " + std::to_string(g) + " (this build has no CPU path)\n"; return; }
		}
		if (gsa_set_params(owners[0], &prm) != 0 || gsa_index_upload(owners[0], &view) != 0) { idx_err = std::string("FatalError: ") + gsa_last_error(owners[0]) + "\n"; return; }
		tick("index uploaded");
		std::vector<std::thread> cl;
		std::vector<std::string> cl_err((size_t)n_dev);
		for (int g = 1; g < n_dev; g++)
			cl.emplace_back([&, g] { if (gsa_set_params(owners[g], &prm) != 0 || gsa_index_clone(owners[g], owners[0]) != 0) cl_err[(size_t)g] = gsa_last_error(owners[g]); });
		for (auto &t : cl) t.join();
		for (int g = 1; g < n_dev; g++) if (!cl_err[(size_t)g].empty()) { idx_err = "FatalError: " + cl_err[(size_t)g] + "\n"; return; }
		if (n_dev > 1) tick("index replicated");
	});
	std::vector<QueryChr> query;
	bool query_ok = check_input_file(o.query) && load_query_file(o.query, query);
	tick("query loaded");
	idx_thread.join();
	if (!query_ok) { fprintf(stderr, "Please check the query file: %s\n", o.query); return 0; }
	if (!idx_err.empty()) { fprintf(stderr, "%s", idx_err.c_str()); return 0; }
	fprintf(stderr, "\tLoad the reference sequences (%d %s)\n", (int)ix.names.size(), ix.names.size() > 1 ? "chromosomes" : "chromosome");
	if (o.show_plot) fprintf(stderr, "Warning! dot-plots need gnuplot and are not produced by this build\n");
	std::string op = o.out_prefix;
	if (o.out_format == 1) o.maf = op + ".maf";
	if (o.out_format == 2) o.aln = op + ".aln";
	o.vcf_name = op + ".vcf";

	// ---- one index replica per GPU; on every GPU `lanes` contexts share it (gsa_create_shared), one host thread each ---------
	int nq = (int)query.size();
	int ngpu = std::min<int>(n_dev, nq);
	// longest-processing-time dealing of contigs to GPUs
	std::vector<int> order((size_t)nq); for (int i = 0; i < nq; i++) order[i] = i;
	std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return query[a].seq.size() > query[b].seq.size(); });
	std::vector<std::vector<int> > work((size_t)ngpu); std::vector<size_t> load((size_t)ngpu, 0);
	for (int qi : order) { int g = (int)(std::min_element(load.begin(), load.end()) - load.begin()); work[g].push_back(qi); load[g] += query[qi].seq.size(); }
	for (auto &w : work) std::sort(w.begin(), w.end()); // each GPU walks its share in contig order so that the emitter is never starved
	std::vector<std::vector<gsa_ctx *> > ctx((size_t)ngpu);
	for (int g = 0; g < ngpu; g++) {
		int nl = std::max(1, std::min<int>(o.lanes, (int)work[g].size()));
		ctx[g].assign((size_t)nl, nullptr);
		ctx[g][0] = owners[g];
		for (int l = 1; l < nl; l++)
			if (gsa_create_shared(ctx[g][0], &ctx[g][l]) != 0) { fprintf(stderr, "FatalError: %s\n", gsa_last_error(ctx[g][0])); return 0; }
	}
	// several GPUs: the finished records stay in HBM, are packed into one outbox per GPU and collected on GPU 0 by a single
	// NCCL gather over NVLink once every contig is aligned (GSA_GATHER=host: every GPU copies its records to the host itself)
	const char *gmode = getenv("GSA_GATHER");
	const bool nccl_gather = ngpu > 1 && !(gmode && strcmp(gmode, "host") == 0);
	if (nccl_gather) {
		if (gsa_comm_init_all(owners.data(), ngpu) != 0) { fprintf(stderr, "FatalError: %s\n", gsa_last_error(owners[0])); return 1; }
		for (int g = 0; g < ngpu; g++) for (gsa_ctx *c : ctx[g]) gsa_set_host_results(c, 0);
		// a first guess at every outbox (about one record byte per query base for closely related genomes); it grows on demand
		// (GSA_OUTBOX_RESERVE=0 starts from nothing: the growth path, for tests)
		const char *rs = getenv("GSA_OUTBOX_RESERVE");
		if (!(rs && strcmp(rs, "0") == 0)) for (int g = 0; g < ngpu; g++) gsa_outbox_reserve(owners[g], (int64_t)load[g] + ((int64_t)64 << 20));
	}
	// GSA_VARIANTS=host: VariantIdentification scans the rows on the host instead of taking the device's records (gsa_variants)
	const char *vmode = getenv("GSA_VARIANTS");
	const bool device_variants = !(vmode && strcmp(vmode, "host") == 0);
	tick("lanes ready");

	// ---- GenomeComparison ------------------------------------------------------------------------------------------------
	fprintf(stderr, "Step2. Sequence analysis for all query chromosomes\n");
	std::vector<ContigResult> results((size_t)nq);
	std::vector<int> done((size_t)nq, 0);
	std::mutex mu; std::condition_variable cv;
	std::vector<size_t> cursor((size_t)ngpu, 0);
	bool failed = false;
	auto run_lane = [&](int g, gsa_ctx *c) { // a lane takes the next contig of its GPU's share until none is left
		for (;;) {
			int qi;
			{ std::unique_lock<std::mutex> lk(mu); if (cursor[g] >= work[g].size()) return; qi = work[g][cursor[g]++]; }
			gsa_alignment al;
			int rc = gsa_align_contig(c, query[qi].seq.data(), (uint32_t)query[qi].seq.size(), &al);
			if (rc == 0 && nccl_gather) rc = gsa_outbox_append(owners[g], c, qi);
			else if (rc == 0) { // the copy out of the pinned buffers runs outside the lock: slot qi is this lane's alone
				results[qi].assign(al);
				gsa_variant_list vl;
				if (o.vcf && device_variants && al.n_blocks > 0 && (rc = gsa_variants(c, &vl)) == 0) results[qi].assign_variants(vl);
			}
			std::unique_lock<std::mutex> lk(mu);
			if (rc != 0) { fprintf(stderr, "FatalError: %s\n", gsa_last_error(c)); failed = true; }
			if (!nccl_gather || rc != 0) done[qi] = 1;
			cv.notify_all();
		}
	};
	std::vector<std::thread> threads;
	for (int g = 0; g < ngpu; g++) for (gsa_ctx *c : ctx[g]) threads.emplace_back(run_lane, g, c);
	std::thread collector;
	if (nccl_gather) { // the collector: all lanes, then the one gather, then the images come to the host rank by rank
		collector = std::thread([&] {
			for (auto &t : threads) t.join();
			bool bad;
			{ std::unique_lock<std::mutex> lk(mu); bad = failed; }
			if (!bad && (gsa_gather_records_all(owners.data(), ngpu, 0) != 0 || gsa_gather_wait(owners[0]) != 0)) bad = true;
			for (int r = 0; r < ngpu && !bad; r++) {
				const void *img = nullptr; int64_t bytes = 0, off = 0, contig = 0;
				if (gsa_inbox_host(owners[0], r, &img, &bytes) != 0) { bad = true; break; }
				std::vector<std::thread> cp; std::vector<int64_t> got;
				gsa_alignment al; int rc;
				std::deque<int> cp_rc;   // one result per copy thread; a deque keeps their addresses stable
				for (int64_t start = off; (rc = gsa_record_next(img, bytes, &off, &contig, &al)) == 1; start = off) {
					if (contig < 0 || contig >= nq) { rc = -1; break; }
					got.push_back(contig);
					cp_rc.push_back(0);
					int *prc = &cp_rc.back();
					// blocks and rows are copied out of the image; the fragment list is expanded from its compact form (or copied)
					cp.emplace_back([&results, contig, al, img, bytes, start, prc] { *prc = results[(size_t)contig].assign_record(al, img, bytes, start); });
				}
				for (auto &t : cp) t.join();
				for (int x : cp_rc) if (x != 0) rc = -1;
				if (rc < 0) { bad = true; break; }
				std::unique_lock<std::mutex> lk(mu);
				for (int64_t qi : got) done[(size_t)qi] = 1;
				cv.notify_all();
			}
			std::unique_lock<std::mutex> lk(mu);
			if (bad) { if (!failed) fprintf(stderr, "FatalError: record gather: %s\n", gsa_last_error(owners[0])); failed = true; for (auto &d : done) d = 1; }
			cv.notify_all();
		});
	}

	EmitState st;
	st.threads = std::max(1, o.threads);
	for (int qi = 0; qi < nq; qi++) {
		fprintf(stderr, "\tProcess query chromsomoe: %s...\n", query[qi].name.c_str());
		ContigResult &r = results[(size_t)qi];
		bool stop;
		{ std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return done[qi] != 0; }); stop = failed; }
		if (stop) break;
		int n = 0; int64_t aln_score = 0, aln_len = 0;
		for (const gsa_block &b : r.blocks) { // src/GSAlign.cpp:529-539 (the identity filter itself ran inside gsa_fill)
			if (b.bDup) st.dup_num++;
			n++; aln_len += b.aln_len; aln_score += b.score;
			st.local_aln_num++; st.total_aln_len += b.aln_len; st.total_matches += b.score;
		}
		if (n == 0) { r = ContigResult(); continue; }
		fprintf(stderr, "\t\tProduce %d local alignments (length = %lld), ANI=%.2f%%\n", n, (long long)aln_len, 100 * (1.0 * aln_score / aln_len));
		// the alignment file is written by this thread while a helper scans the same records for variants: the scan skips seed
		// fragments, the only thing the writer touches (iExtension trims the last seed of a block)
		if (o.out_format == 1 || o.out_format == 2) fprintf(stderr, "\t\tOutput alignments for query sequence (%s)\n", o.out_format == 1 ? o.maf.c_str() : o.aln.c_str());
		if (o.vcf) fprintf(stderr, "\t\tIdentify sequence variants for query sequence...\n");
		std::thread var_thread;
		if (o.vcf) var_thread = std::thread([&] { variant_identification(ix, query, qi, r, st); });
		if (o.out_format == 1) output_maf(o, ix, query, qi, r);
		if (o.out_format == 2) output_aln(o, ix, query, qi, r);
		if (var_thread.joinable()) var_thread.join();
		fprintf(stderr, "\n");
		r = ContigResult();
	}
	if (nccl_gather) collector.join(); // the collector has joined the lanes
	else for (auto &t : threads) t.join();
	tick("align + emit");
	if (failed) { // a device or limit failure mid-run: no partial files are left behind and the exit code says so (the reference's
		// always-0 convention covers usage errors, not an aborted run)
		emit_drain();
		if (!o.maf.empty()) remove(o.maf.c_str());
		if (!o.aln.empty()) remove(o.aln.c_str());
		fflush(NULL);
		_exit(1);
	}
	if (st.local_aln_num > 0)
		fprintf(stderr, "\tAlignment#=%d (total alignment length=%lld) ANI=%.2f%%, unique alignment#=%d\n", (int)st.local_aln_num, (long long)st.total_aln_len,
		        100 * (1.0 * st.total_matches / st.total_aln_len), (int)(st.local_aln_num - st.dup_num));
	fprintf(stderr, "\tIt took %lld seconds for genome sequence alignment.\n", (long long)(time(NULL) - t_start));
	if (o.vcf) {
		fprintf(stderr, "\nGSAlign identifies %d SNVs, %d insertions, and %d deletions [%s].\n\n", st.iSNV, st.iInsertion, st.iDeletion, o.vcf_name.c_str());
		output_variants(o, ix, st);
	}
	if (!emit_drain()) { fprintf(stderr, "FatalError: cannot write the output files\n"); fflush(NULL); _exit(1); }
	tick("variants written");
	// the process ends here: files are closed, device and pinned memory go back with the process (tearing contexts down one
	// buffer at a time costs more than the whole alignment of a small genome)
	fflush(NULL);
	tick("done");
	_exit(0);
}
