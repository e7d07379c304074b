/* oracle/gsa_oracle.c -- TEST INFRASTRUCTURE ONLY (see gsa_oracle.h).
 *
 * CPU restatement, in plain C, of the reference's hot path.  Each function cites the reference
 * file:line it follows.  It is deliberately simple and serial: it exists to be obviously faithful,
 * not fast.  Pinned against the compiled reference by tests/test_oracle_vs_reference.py.
 */
#include "gsa_oracle.h"
#include <stdlib.h>
#include <string.h>

#define SEED_CHUNK    10000   /* SeedExplorationChunk, src/GSAlign.cpp:5 */
#define MAX_SEED_FREQ 100     /* MaxSeedFreq, src/bwt_search.cpp:3 */
#define MAX_SEED_GAP  5000    /* MaxSeedGap, src/structure.h:23 */

void orc_free(void *p) { free(p); }

void orc_default_params(orc_params_t *p)
{ /* src/main.cpp:203-215 */
	p->min_seed_len = 15; p->sensitive = 0; p->max_indel = 25;
	p->min_block_score = 200; p->min_aln_len = 200; p->min_idy = 70;
}

/* nst_nt4_table (src/BWT_Index/bntseq.c:40-57): A/a 0, C/c 1, G/g 2, T/t 3, everything else 4 */
static inline int nt4(char ch)
{
	switch (ch) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': return 3;
	default: return 4;
	}
}

static inline int64_t i64abs(int64_t x) { return x < 0 ? -x : x; }

/* ---------------------------------------------------------------------------------------------
 * reference text
 * ------------------------------------------------------------------------------------------- */
static inline int pac_base(const orc_index_t *idx, int64_t pos)
{ /* src/bwt_index.cpp:201 */
	return idx->pac[pos >> 2] >> ((~pos & 3) << 1) & 3;
}

static inline int text_code(const orc_index_t *idx, int64_t pos)
{ /* RefSequence[f] = base, RefSequence[2N-1-f] = complement (src/bwt_index.cpp:199-209) */
	if (pos < idx->l_pac) return pac_base(idx, pos);
	return 3 - pac_base(idx, 2 * idx->l_pac - 1 - pos);
}

char orc_text_char(const orc_index_t *idx, int64_t pos) { return "ACGT"[text_code(idx, pos)]; }

/* ---------------------------------------------------------------------------------------------
 * FM-index rank queries on the BWA layout (src/bwt_search.cpp:36-119)
 * ------------------------------------------------------------------------------------------- */
static inline int popc32(uint32_t x) { return __builtin_popcount(x); }

/* number of symbols == c among the first m (1..16) symbols (MSB first) of word w */
static inline int word_count(uint32_t w, int m, int c)
{
	uint32_t x = w ^ ((uint32_t)c * 0x55555555u);   /* matching symbols become 00 */
	uint32_t z = ~(x | (x >> 1)) & 0x55555555u;     /* 1 at the low bit of every matching symbol */
	if (m < 16) z &= ~((1u << ((16 - m) << 1)) - 1); /* keep the top m symbols */
	return popc32(z);
}

/* Occ(c, k) for all four c; k is a row, (uint64_t)-1 allowed (bwt_occ4, src/bwt_search.cpp:69-86) */
static void occ4(const orc_index_t *idx, uint64_t k, uint64_t cnt[4])
{
	if (k == (uint64_t)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
	k -= (k >= idx->primary);
	const uint32_t *p = idx->bwt + ((k >> 7) << 4);
	memcpy(cnt, p, 4 * sizeof(uint64_t));
	p += 8;
	int nsym = (int)(k & 127) + 1; /* symbols of this block up to and including k */
	for (int w = 0; nsym > 0; w++, nsym -= 16) {
		int m = nsym >= 16 ? 16 : nsym;
		for (int c = 0; c < 4; c++) cnt[c] += (uint64_t)word_count(p[w], m, c);
	}
}

static uint64_t occ1(const orc_index_t *idx, uint64_t k, int c)
{ /* bwt_occ, src/bwt_search.cpp:45-67 */
	uint64_t cnt[4];
	if (k == idx->seq_len) return idx->L2[c + 1] - idx->L2[c];
	if (k == (uint64_t)-1) return 0;
	occ4(idx, k, cnt);
	return cnt[c];
}

static inline int bwt_char(const orc_index_t *idx, uint64_t x)
{ /* bwt_B0, src/bwt_search.cpp:32-34: x is an index into the $-less BWT */
	uint32_t w = idx->bwt[((x >> 7) << 4) + 8 + ((x & 127) >> 4)];
	return (int)(w >> ((~x & 15) << 1) & 3);
}

static uint64_t inv_psi(const orc_index_t *idx, uint64_t k)
{ /* bwt_invPsi, src/bwt_search.cpp:121-127 */
	if (k == idx->primary) return 0;
	uint64_t x = k - (k > idx->primary);
	int c = bwt_char(idx, x);
	return idx->L2[c] + occ1(idx, k, c);
}

static uint64_t sa_lookup(const orc_index_t *idx, uint64_t k, orc_counters_t *ctr)
{ /* bwt_sa, src/bwt_search.cpp:129-139 */
	uint64_t steps = 0, mask = (uint64_t)idx->sa_intv - 1;
	while (k & mask) {
		steps++;
		k = inv_psi(idx, k);
		if (ctr) ctr->n_lf_steps++;
	}
	if (ctr) ctr->n_sa_reads++;
	return steps + idx->sa[k / (uint64_t)idx->sa_intv];
}

void orc_bwt_search(const orc_index_t *idx, const char *seq, int32_t start, int32_t stop, int32_t min_seed_len,
                    int32_t *len, int32_t *freq, uint64_t *loc, orc_counters_t *ctr)
{ /* BWT_Search, src/bwt_search.cpp:141-185: FMD forward extension of seq[start..stop) */
	uint64_t x0, x1, x2, tk[4], tl[4];
	int p = nt4(seq[start]), pos;
	x0 = idx->L2[p] + 1; x1 = idx->L2[3 - p] + 1; x2 = idx->L2[p + 1] - idx->L2[p];
	if (ctr) ctr->n_search++;
	for (pos = start + 1; pos < stop; pos++) {
		int nt = nt4(seq[pos]);
		if (nt > 3) break;
		uint64_t k = x1 - 1, l = x1 - 1 + x2;
		if (ctr) {
			uint64_t _k = k - (k >= idx->primary), _l = l - (l >= idx->primary);
			ctr->n_ext_steps++;
			if ((_k >> 7) != (_l >> 7) || k == (uint64_t)-1 || l == (uint64_t)-1) ctr->n_split++;
		}
		occ4(idx, k, tk); occ4(idx, l, tl);
		uint64_t o1[4], o2[4], o0[4];
		for (int i = 0; i < 4; i++) { o1[i] = idx->L2[i] + 1 + tk[i]; o2[i] = tl[i] - tk[i]; }
		o0[3] = x0 + (x1 <= idx->primary && x1 + x2 - 1 >= idx->primary);
		o0[2] = o0[3] + o2[3]; o0[1] = o0[2] + o2[2]; o0[0] = o0[1] + o2[1];
		int i = 3 - nt;
		if (o2[i] == 0) break;
		x0 = o0[i]; x1 = o1[i]; x2 = o2[i];
	}
	*len = pos - start; *freq = 0;
	if (*len < min_seed_len) { if (ctr) ctr->n_short++; return; }
	if (x2 > MAX_SEED_FREQ) { if (ctr) ctr->n_freqskip++; return; }
	*freq = (int32_t)x2;
	if (ctr) ctr->n_seedhit++;
	for (int32_t i = 0; i < *freq; i++) loc[i] = sa_lookup(idx, x0 + (uint64_t)i, ctr);
}

/* ---------------------------------------------------------------------------------------------
 * seeding driver
 * ------------------------------------------------------------------------------------------- */
typedef struct { int32_t q; int64_t r; int32_t len; int64_t pd; int alive; } seed_t;

static int cmp_pd_q(const void *a, const void *b)
{ /* CompByPosDiff, src/ProcessCandidateAlignment.cpp:3-7 */
	const seed_t *x = (const seed_t *)a, *y = (const seed_t *)b;
	if (x->pd != y->pd) return x->pd < y->pd ? -1 : 1;
	return (x->q > y->q) - (x->q < y->q);
}

static int cmp_q_r(const void *a, const void *b)
{ /* CompByQueryPos, src/ProcessCandidateAlignment.cpp:9-13 */
	const seed_t *x = (const seed_t *)a, *y = (const seed_t *)b;
	if (x->q != y->q) return x->q < y->q ? -1 : 1;
	return (x->r > y->r) - (x->r < y->r);
}

int64_t orc_seed_contig(const orc_index_t *idx, const orc_params_t *prm, const char *seq, int64_t seqlen,
                        int32_t **q, int64_t **r, int32_t **l, orc_counters_t *ctr)
{ /* IdentifyLocalMEM, src/GSAlign.cpp:51-107 */
	int64_t cap = 1024, n = 0;
	seed_t *v = (seed_t *)malloc((size_t)cap * sizeof(seed_t));
	uint64_t loc[MAX_SEED_FREQ];
	for (int64_t cs = 0; cs < seqlen; cs += SEED_CHUNK) {
		int64_t start = cs, stop = cs + SEED_CHUNK;
		if (stop > seqlen) stop = seqlen;
		while (start < stop) {
			if (nt4(seq[start]) > 3) { start++; continue; }
			int32_t len, freq;
			orc_bwt_search(idx, seq, (int32_t)start, (int32_t)stop, prm->min_seed_len, &len, &freq, loc, ctr);
			if (freq > 0) {
				for (int32_t i = 0; i < freq; i++) {
					if (n == cap) { cap *= 2; v = (seed_t *)realloc(v, (size_t)cap * sizeof(seed_t)); }
					v[n].q = (int32_t)start; v[n].r = (int64_t)loc[i]; v[n].len = len;
					v[n].pd = v[n].r - v[n].q; v[n].alive = 1; n++;
				}
				start += prm->sensitive ? 5 : len + 1;
			} else start++;
		}
	}
	qsort(v, (size_t)n, sizeof(seed_t), cmp_pd_q);
	*q = (int32_t *)malloc((size_t)(n ? n : 1) * sizeof(int32_t));
	*r = (int64_t *)malloc((size_t)(n ? n : 1) * sizeof(int64_t));
	*l = (int32_t *)malloc((size_t)(n ? n : 1) * sizeof(int32_t));
	for (int64_t i = 0; i < n; i++) { (*q)[i] = v[i].q; (*r)[i] = v[i].r; (*l)[i] = v[i].len; }
	free(v);
	return n;
}

/* ---------------------------------------------------------------------------------------------
 * cluster / chain
 * ------------------------------------------------------------------------------------------- */
typedef struct { int32_t q; int64_t r; int32_t ql, rl; int alive; } frag_t;
typedef struct { int32_t score; int64_t beg, end; } blk_t; /* frags[beg,end) of the pool */

typedef struct {
	frag_t *f; int64_t nf, capf;
	blk_t *b; int64_t nb, capb;
} pool_t;

static void pool_push_frag(pool_t *P, frag_t x)
{
	if (P->nf == P->capf) { P->capf = P->capf ? P->capf * 2 : 1024; P->f = (frag_t *)realloc(P->f, (size_t)P->capf * sizeof(frag_t)); }
	P->f[P->nf++] = x;
}

static void pool_push_blk(pool_t *P, blk_t x)
{
	if (P->nb == P->capb) { P->capb = P->capb ? P->capb * 2 : 64; P->b = (blk_t *)realloc(P->b, (size_t)P->capb * sizeof(blk_t)); }
	P->b[P->nb++] = x;
}

static int cmp_i64(const void *a, const void *b)
{
	int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
	return (x > y) - (x < y);
}

/* RemoveOutlierSeeds + RefinePDFmap + Check_PD_Frequency, src/GSAlign.cpp:145-153,245-296.
 * s[beg,end) is one window; uniq is indexed like s. */
static void remove_outliers(seed_t *s, const char *uniq, int64_t beg, int64_t end, int64_t genome_size, int max_indel)
{
	int64_t nu = 0;
	for (int64_t i = beg; i < end; i++) if (uniq[i]) nu++;
	if (nu == 0) return;
	int64_t *bins = (int64_t *)malloc((size_t)nu * sizeof(int64_t));
	int64_t k = 0;
	for (int64_t i = beg; i < end; i++) if (uniq[i]) bins[k++] = (int64_t)(int)(s[i].pd >> 4);
	qsort(bins, (size_t)nu, sizeof(int64_t), cmp_i64);
	/* mode = smallest key with the maximal count (src/GSAlign.cpp:250-251) */
	int64_t mode = 0, best = 0;
	for (int64_t i = 0; i < nu;) {
		int64_t j = i; while (j < nu && bins[j] == bins[i]) j++;
		if (j - i > best) { best = j - i; mode = bins[i]; }
		i = j;
	}
	/* average PosDiff over unique seeds whose bin survives |key - mode| < 3 (src/GSAlign.cpp:254-282) */
	int64_t sum = 0, n = 0;
	for (int64_t i = beg; i < end; i++) if (uniq[i]) {
		int64_t key = (int64_t)(int)(s[i].pd >> 4);
		if (i64abs(key - mode) < 3) { sum += s[i].pd; n++; }
	}
	int64_t avg = n > 0 ? sum / n : genome_size;
	for (int64_t i = beg; i < end; i++) if (uniq[i]) {
		int64_t key = (int64_t)(int)(s[i].pd >> 4), cnt = 0;
		if (i64abs(key - mode) < 3) { /* surviving bins keep their count, others are zeroed */
			int64_t lo = 0, hi = nu; /* count occurrences of key in the sorted bins */
			while (lo < hi) { int64_t m = (lo + hi) / 2; if (bins[m] < key) lo = m + 1; else hi = m; }
			int64_t a = lo; hi = nu;
			while (lo < hi) { int64_t m = (lo + hi) / 2; if (bins[m] <= key) lo = m + 1; else hi = m; }
			cnt = lo - a;
		}
		if (i64abs(avg - s[i].pd) > max_indel && cnt < 3) s[i].alive = 0; /* Min_PD_Freq 3, src/GSAlign.cpp:4,290 */
	}
	free(bins);
}

/* AddAlnBlock, src/GSAlign.cpp:29-49 (seeds s[i,j) are live, ordered by qPos) */
static void add_block(pool_t *P, const seed_t *s, int64_t i, int64_t j, const orc_params_t *prm)
{
	int32_t score = 0;
	for (int64_t k = i; k < j; k++) score += s[k].len;
	int32_t region = (s[j - 1].q + s[j - 1].len) - s[i].q;
	if (score < prm->min_block_score || region < prm->min_aln_len || (score < 1000 && score < region * 0.05)) return;
	blk_t b; b.score = score; b.beg = P->nf;
	for (int64_t k = i; k < j; k++) {
		frag_t f; f.q = s[k].q; f.r = s[k].r; f.ql = f.rl = s[k].len; f.alive = 1;
		pool_push_frag(P, f);
	}
	b.end = P->nf;
	pool_push_blk(P, b);
}

/* SeedGroupAnalysis, src/GSAlign.cpp:305-375 on the group s[0,n) */
static void group_analysis(pool_t *P, seed_t *s, int64_t n, const orc_index_t *idx, const orc_params_t *prm)
{
	qsort(s, (size_t)n, sizeof(seed_t), cmp_q_r);
	char *uniq = (char *)calloc((size_t)n, 1);
	for (int64_t i = 0; i < n;) { /* :316-325 */
		int64_t j = i + 1; while (j < n && s[j].q == s[i].q) j++;
		if (j == i + 1) uniq[i] = 1;
		i = j;
	}
	int64_t i0 = 0; int cnt = uniq[0] ? 1 : 0;
	for (int64_t j = 1; j < n; j++) { /* :326-337 */
		if (!uniq[j]) continue;
		if (s[j].pd == s[j - 1].pd) cnt++;
		else if (++cnt >= 30 && s[j].q - s[i0].q > 3000) {
			remove_outliers(s, uniq, i0, j, idx->l_pac, prm->max_indel);
			i0 = j; cnt = 0;
		}
	}
	remove_outliers(s, uniq, i0, n, idx->l_pac, prm->max_indel);
	for (int64_t i = 0; i < n;) { /* multi-hit runs, :341-350 with :178-225 */
		int64_t j = i + 1; while (j < n && s[j].q == s[i].q) j++;
		if (j > i + 1) {
			int64_t sum1 = 0, sum2 = 0; int n1 = 0, n2 = 0;
			for (int64_t p = i - 1; p >= 0; p--) if (uniq[p] && s[p].alive) { n1++; sum1 += s[p].pd; if (n1 == 5) break; }
			for (int64_t p = j; p < n; p++) if (uniq[p] && s[p].alive) { n2++; sum2 += s[p].pd; if (n2 == 5) break; }
			int64_t avg = (n1 > 0 || n2 > 0) ? (sum1 + sum2) / (n1 + n2) : s[i].pd;
			int64_t keep = -1, min_diff = idx->l_pac;
			for (int64_t k = i; k < j; k++) {
				int64_t d = i64abs(s[k].pd - avg);
				if (d < prm->max_indel && d < min_diff) { min_diff = d; keep = k; }
			}
			for (int64_t k = i; k < j; k++) if (k != keep) s[k].alive = 0;
		}
		i = j;
	}
	free(uniq);
	/* compact (CompByRemoval sort + trim, :353) */
	int64_t m = 0;
	for (int64_t i = 0; i < n; i++) if (s[i].alive) s[m++] = s[i];
	/* noise, :355-362: evaluated on the pre-filter neighbour list */
	for (int64_t j = 1; j + 1 < m; j++)
		if (i64abs(s[j].pd - s[j - 1].pd) > 5 && i64abs(s[j].pd - s[j + 1].pd) > 5) s[j].alive = 0;
	int64_t m2 = 0;
	for (int64_t i = 0; i < m; i++) if (s[i].alive) s[m2++] = s[i];
	if (m2 == 0) return; /* the reference under-runs here (SURVEY hazard H14); no block is the sane reading */
	int64_t p = 0;
	for (int64_t j = 1; j < m2; j++) /* :364-374 */
		if (s[j].q - s[j - 1].q - s[j - 1].len > MAX_SEED_GAP || i64abs(s[j - 1].pd - s[j].pd) > 100) { add_block(P, s, p, j, prm); p = j; }
	add_block(P, s, p, m2, prm);
}

/* RemoveOverlaps, src/ProcessCandidateAlignment.cpp:189-231; compacts f[0,n) in place, returns new n */
static int64_t remove_overlaps(frag_t *f, int64_t n)
{
	for (;;) {
		int modified = 0;
		for (int64_t i = 0, j = 1; j < n; i++, j++) {
			int32_t ov;
			if (f[j].r <= f[i].r) { modified = 1; f[i].alive = 0; continue; }
			if ((ov = (int32_t)(f[i].r + f[i].rl - f[j].r)) > 0) {
				f[i].ql -= ov; f[i].rl -= ov;
				if (f[i].ql <= 0 || f[i].rl <= 0) { modified = 1; f[i].alive = 0; continue; }
			}
			if ((ov = f[i].q + f[i].ql - f[j].q) > 0) {
				f[i].ql -= ov; f[i].rl -= ov;
				if (f[i].ql <= 0 || f[i].rl <= 0) { modified = 1; f[i].alive = 0; continue; }
			}
		}
		if (!modified) break;
		int64_t m = 0;
		for (int64_t i = 0; i < n; i++) if (f[i].alive) f[m++] = f[i];
		n = m;
	}
	return n;
}

/* CreateKmerVecFromReadSeq, src/KmerAnalysis.cpp:32-76 -- returns a histogram instead of a sorted
 * vector (ids < 2048); quirks kept: only the byte 'N' restarts, stale `head` after a restart */
static void kmer_hist(const char *seq, int len, int *hist)
{
	uint32_t wid, count = 0, head = 0, tail = 0;
	while (count < 5 && tail < (uint32_t)len) { if (seq[tail++] != 'N') count++; else count = 0; }
	if (count != 5) return;
#define KMER_ID(h) do { wid = 0; for (uint32_t _i = (h); _i < (h) + 5; _i++) wid = (wid << 2) + (uint32_t)nt4(seq[_i]); } while (0)
	KMER_ID(head); hist[wid]++;
	for (head += 1; tail < (uint32_t)len; head++, tail++) {
		if (seq[tail] != 'N') { wid = ((wid & 0xFF) << 2) + (uint32_t)nt4(seq[tail]); hist[wid]++; }
		else {
			count = 0; tail++;
			while (count < 5 && tail < (uint32_t)len) { if (seq[tail++] != 'N') count++; else count = 0; }
			if (count == 5) { KMER_ID(head); hist[wid]++; }
			else break;
		}
	}
#undef KMER_ID
}

int32_t orc_gap_similarity(const orc_index_t *idx, const char *seq, int32_t q1, int32_t q2, int64_t r1, int64_t r2)
{ /* CalGapSimilarity, src/KmerAnalysis.cpp:78-121 */
	int q_len = q2 - q1, r_len = (int)(r2 - r1), similar = 0;
	if (r1 - q1 == r2 - q2) {
		int idy = 0; int64_t r = r1;
		for (int q = q1; q < q2; q++, r++) {
			int a = text_code(idx, r), b = nt4(seq[q]);
			if (a == b || b == 4) idy++; /* the reference text never holds N */
		}
		if (idy >= q_len * 0.5) similar = 1;
	}
	if (!similar && q_len <= MAX_SEED_GAP && r_len <= MAX_SEED_GAP) {
		int *h1 = (int *)calloc(4096, sizeof(int)), *h2 = h1 + 2048;
		char *rf = (char *)malloc((size_t)(r_len > 0 ? r_len : 1));
		for (int i = 0; i < r_len; i++) rf[i] = orc_text_char(idx, r1 + i);
		kmer_hist(seq + q1, q_len, h1);
		kmer_hist(rf, r_len, h2);
		int common = 0;
		for (int i = 0; i < 2048; i++) common += h1[i] < h2[i] ? h1[i] : h2[i];
		if (common > (q_len + r_len) * 0.1) similar = 1;
		free(h1); free(rf);
	}
	return similar;
}

static int32_t block_score(const frag_t *f, int64_t n, const orc_params_t *prm)
{ /* CalAlnBlockScore, src/ProcessCandidateAlignment.cpp:26-36 */
	if (n == 0) return 0;
	if ((f[n - 1].q + f[n - 1].ql - f[0].q) < prm->min_aln_len) return 0;
	int32_t s = 0;
	for (int64_t i = 0; i < n; i++) s += f[i].ql;
	return s;
}

static int64_t contig_end(const orc_index_t *idx, int64_t rpos)
{ /* ChrLocMap.lower_bound(rpos)->first, src/bwt_index.cpp:247-252 */
	int64_t best = -1, total = 0, n2 = 2 * idx->l_pac;
	for (int i = 0; i < idx->n_contigs; i++) {
		int64_t fe = total + idx->contig_len[i] - 1;
		total += idx->contig_len[i];
		int64_t re = (n2 - total) + idx->contig_len[i] - 1;
		if (fe >= rpos && (best < 0 || fe < best)) best = fe;
		if (re >= rpos && (best < 0 || re < best)) best = re;
	}
	return best;
}

/* splits block bi of P at the given break points (CheckGapsBetweenSeeds / CheckAlnBlockSpanMultipleRefChrs
 * tails, src/ProcessCandidateAlignment.cpp:100-117,140-155): parent dies, pieces with score > clr are pushed */
static void split_block(pool_t *P, int64_t bi, const int64_t *brk, int64_t nbrk, const orc_params_t *prm)
{
	if (nbrk == 0) return;
	int64_t beg = P->b[bi].beg, end = P->b[bi].end;
	P->b[bi].score = 0;
	int64_t i = 0;
	for (int64_t k = 0; k <= nbrk; k++) {
		int64_t j = k < nbrk ? brk[k] : end - beg;
		int32_t sc = block_score(P->f + beg + i, j - i, prm);
		if (sc > prm->min_block_score) {
			blk_t nb; nb.score = sc; nb.beg = P->nf;
			for (int64_t t = i; t < j; t++) pool_push_frag(P, P->f[beg + t]);
			nb.end = P->nf;
			pool_push_blk(P, nb);
		}
		i = j;
	}
}

static int cmp_blk_score_desc(const void *a, const void *b)
{ /* CompByAlnBlockScore; beg as a deterministic tie-break (the reference's order on ties is libstdc++'s) */
	const blk_t *x = (const blk_t *)a, *y = (const blk_t *)b;
	if (x->score != y->score) return x->score > y->score ? -1 : 1;
	return (x->beg > y->beg) - (x->beg < y->beg);
}

static void remove_bad_blocks(pool_t *P)
{ /* RemoveBadAlnBlocks, src/ProcessCandidateAlignment.cpp:72-79 */
	qsort(P->b, (size_t)P->nb, sizeof(blk_t), cmp_blk_score_desc);
	while (P->nb > 0 && P->b[P->nb - 1].score == 0) P->nb--;
}

static int64_t serialise(const pool_t *P, int64_t **out)
{
	int64_t words = 1;
	for (int64_t b = 0; b < P->nb; b++) words += 4 + 5 * (P->b[b].end - P->b[b].beg);
	int64_t *o = (int64_t *)malloc((size_t)words * sizeof(int64_t)), w = 0;
	o[w++] = P->nb;
	for (int64_t b = 0; b < P->nb; b++) {
		o[w++] = P->b[b].score; o[w++] = 0; o[w++] = 0; o[w++] = P->b[b].end - P->b[b].beg;
		for (int64_t t = P->b[b].beg; t < P->b[b].end; t++) {
			o[w++] = 1; o[w++] = P->f[t].q; o[w++] = P->f[t].r; o[w++] = P->f[t].ql; o[w++] = P->f[t].rl;
		}
	}
	*out = o;
	return words;
}

int64_t orc_cluster(const orc_index_t *idx, const orc_params_t *prm, const char *seq, int64_t seqlen,
                    int64_t nseeds, const int32_t *q, const int64_t *r, const int32_t *l,
                    int32_t stage, int64_t **out)
{
	(void)seqlen;
	pool_t P; memset(&P, 0, sizeof(P));
	seed_t *s = (seed_t *)malloc((size_t)(nseeds ? nseeds : 1) * sizeof(seed_t));
	for (int64_t i = 0; i < nseeds; i++) { s[i].q = q[i]; s[i].r = r[i]; s[i].len = l[i]; s[i].pd = r[i] - q[i]; s[i].alive = 1; }
	/* SeedGrouping (src/GSAlign.cpp:126-143) + GenerateAlignmentBlocks (:377-391) */
	for (int64_t p = 0; p < nseeds;) {
		int64_t j = p + 1;
		while (j < nseeds && s[j].pd - s[j - 1].pd <= prm->max_indel) j++;
		int64_t score = 0;
		for (int64_t k = p; k < j; k++) score += s[k].len;
		if (score >= prm->min_block_score) group_analysis(&P, s + p, j - p, idx, prm);
		p = j;
	}
	free(s);
	if (stage >= 1) { /* CheckAlnBlockOverlaps; blocks keep their score (src/ProcessCandidateAlignment.cpp:232-239) */
		for (int64_t b = 0; b < P.nb; b++) P.b[b].end = P.b[b].beg + remove_overlaps(P.f + P.b[b].beg, P.b[b].end - P.b[b].beg);
	}
	if (stage >= 2) {
		int64_t nb0 = P.nb; /* CheckAlnBlockLargeGaps, :120-156,166-172 */
		for (int64_t b = 0; b < nb0; b++) {
			int64_t n = P.b[b].end - P.b[b].beg, nbrk = 0;
			int64_t *brk = (int64_t *)malloc((size_t)(n ? n : 1) * sizeof(int64_t));
			for (int64_t i = 0, j = 1; j < n; i++, j++) {
				const frag_t *a = &P.f[P.b[b].beg + i], *c = &P.f[P.b[b].beg + j];
				int32_t qg = c->q - a->q - a->ql; int64_t rg = c->r - a->r - a->rl;
				if (qg > 300 || (int32_t)rg > 300)
					if (qg > MAX_SEED_GAP || (int32_t)rg > MAX_SEED_GAP || !orc_gap_similarity(idx, seq, a->q + a->ql, c->q, a->r + a->rl, c->r)) brk[nbrk++] = j;
			}
			split_block(&P, b, brk, nbrk, prm);
			free(brk);
		}
		remove_bad_blocks(&P);
		nb0 = P.nb; /* CheckAlnBlockSpanMultiSeqs, :81-118,158-164 */
		for (int64_t b = 0; b < nb0; b++) {
			int64_t n = P.b[b].end - P.b[b].beg, nbrk = 0, last = -1;
			int64_t *brk = (int64_t *)malloc((size_t)(n ? n : 1) * sizeof(int64_t));
			for (int64_t i = 0, j = 1; j < n; j++) {
				if (last == -1) last = contig_end(idx, P.f[P.b[b].beg + i].r);
				if (P.f[P.b[b].beg + j].r > last) { brk[nbrk++] = j; i = j; last = contig_end(idx, P.f[P.b[b].beg + i].r); }
			}
			split_block(&P, b, brk, nbrk, prm);
			free(brk);
		}
		remove_bad_blocks(&P);
	}
	int64_t words = serialise(&P, out);
	free(P.f); free(P.b);
	return words;
}

int64_t orc_normal_pairs(int64_t nfrag, const int64_t *fr, int64_t **out)
{ /* IdentifyNormalPairs, src/ProcessCandidateAlignment.cpp:241-265: the inplace_merge by (qPos,rPos)
   * puts every new gap fragment right before the seed that follows it */
	int64_t *o = (int64_t *)malloc((size_t)(2 * nfrag + 1) * 5 * sizeof(int64_t)), n = 0;
	for (int64_t i = 0; i < nfrag; i++) {
		memcpy(o + 5 * n, fr + 5 * i, 5 * sizeof(int64_t)); n++;
		if (nfrag == 1 || i + 1 == nfrag) continue;
		int64_t qe = fr[5 * i + 1] + fr[5 * i + 3], re = fr[5 * i + 2] + fr[5 * i + 4];
		int64_t qg = fr[5 * (i + 1) + 1] - qe, rg = fr[5 * (i + 1) + 2] - re;
		if (qg < 0) qg = 0;
		if (rg < 0) rg = 0;
		if (qg > 0 || rg > 0) { o[5 * n] = 0; o[5 * n + 1] = qe; o[5 * n + 2] = re; o[5 * n + 3] = qg; o[5 * n + 4] = rg; n++; }
	}
	*out = o;
	return n;
}

/* ---------------------------------------------------------------------------------------------
 * gapped fill
 * ------------------------------------------------------------------------------------------- */
#define NEG_INF (-0x3fffffff)

int32_t orc_dp_align(const char *ref_frag, int32_t m, const char *qry_frag, int32_t n, char *out1, char *out2)
{ /* ksw2_alignment -> ksw_extz2_sse -> ksw_backtrack (src/ksw2_alignment.cpp:25-273), restated as
   * an absolute-score Gotoh recurrence: rows i over the query fragment, columns j over the reference
   * fragment; match +1, mismatch -1, any non-ACGT 0; gap of length L costs 2 + L.
   * Per cell we keep: dir (0 diag, 1 E, 2 F), xe (E of the next row extends E here), xf (likewise F). */
	size_t cells = (size_t)m * (size_t)n;
	uint8_t *fl = (uint8_t *)malloc(cells ? cells : 1);
	int32_t *H = (int32_t *)malloc((size_t)(m + 1) * sizeof(int32_t)); /* previous row, H[j+1] = H(i-1, j) */
	int32_t *E = (int32_t *)malloc((size_t)(m + 1) * sizeof(int32_t)); /* E(i, j) for the row being built */
	for (int j = 0; j < m; j++) { H[j + 1] = -(2 + (j + 1)); E[j + 1] = NEG_INF; }
	H[0] = 0;
	for (int i = 0; i < n; i++) {
		int qc = nt4(qry_frag[i]);
		int32_t hdiag = H[0];               /* H(i-1, -1) */
		int32_t hleft = -(2 + (i + 1));     /* H(i, -1) */
		int32_t f = NEG_INF;                /* F(i, -1) */
		H[0] = hleft;
		for (int j = 0; j < m; j++) {
			int rc = nt4(ref_frag[j]);
			int32_t s = (qc > 3 || rc > 3) ? 0 : (qc == rc ? 1 : -1);
			int32_t hup = H[j + 1];         /* H(i-1, j) */
			/* E(i,j) = max(H(i-1,j) - 3, E(i-1,j) - 1): consumes a query base, tie -> open */
			int32_t eo = hup - 3, ee = E[j + 1] - 1;
			int xe_prev = ee > eo;          /* belongs to cell (i-1, j) */
			int32_t e = xe_prev ? ee : eo;
			/* F(i,j) = max(H(i,j-1) - 3, F(i,j-1) - 1): consumes a reference base, tie -> open */
			int32_t fo = hleft - 3, fe = f - 1;
			int xf_prev = fe > fo;          /* belongs to cell (i, j-1) */
			f = xf_prev ? fe : fo;
			int32_t h = hdiag + s; int d = 0;
			if (e > h) { h = e; d = 1; }
			if (f > h) { h = f; d = 2; }
			fl[(size_t)i * m + j] = (uint8_t)d;
			if (i > 0 && xe_prev) fl[(size_t)(i - 1) * m + j] |= 8;
			if (j > 0 && xf_prev) fl[(size_t)i * m + j - 1] |= 16;
			E[j + 1] = e;
			hdiag = hup; H[j + 1] = h; hleft = h;
		}
	}
	/* traceback, ksw_backtrack src/ksw2_alignment.cpp:25-68 */
	int32_t L = 0, i = n - 1, j = m - 1, state = 0;
	char *t1 = (char *)malloc((size_t)(m + n + 1)), *t2 = (char *)malloc((size_t)(m + n + 1));
	while (i >= 0 && j >= 0) {
		uint8_t t = fl[(size_t)i * m + j];
		if (state == 0) state = t & 7;
		else if (!((t >> (state + 2)) & 1)) state = t & 7;
		if (state == 0) { t1[L] = ref_frag[j]; t2[L] = qry_frag[i]; L++; i--; j--; }
		else if (state == 1) { t1[L] = '-'; t2[L] = qry_frag[i]; L++; i--; }
		else { t1[L] = ref_frag[j]; t2[L] = '-'; L++; j--; }
	}
	while (i >= 0) { t1[L] = '-'; t2[L] = qry_frag[i]; L++; i--; }
	while (j >= 0) { t1[L] = ref_frag[j]; t2[L] = '-'; L++; j--; }
	for (int32_t k = 0; k < L; k++) { out1[k] = t1[L - 1 - k]; out2[k] = t2[L - 1 - k]; }
	out1[L] = out2[L] = 0;
	free(t1); free(t2); free(fl); free(H); free(E);
	return L;
}

int32_t orc_frag_align(const orc_index_t *idx, const char *seq, int32_t qPos, int64_t rPos, int32_t qLen, int32_t rLen,
                       char *out1, char *out2, int32_t *score_inc, int32_t *used_dp)
{ /* GenerateFragAlignment, non-seed branch, src/ProcessCandidateAlignment.cpp:308-342 */
	*used_dp = 0; *score_inc = 0;
	if (qLen == 0) {
		for (int i = 0; i < rLen; i++) { out1[i] = orc_text_char(idx, rPos + i); out2[i] = '-'; }
		out1[rLen] = out2[rLen] = 0;
		return rLen;
	}
	if (rLen == 0) {
		for (int i = 0; i < qLen; i++) { out1[i] = '-'; out2[i] = seq[qPos + i]; }
		out1[qLen] = out2[qLen] = 0;
		return qLen;
	}
	char *rf = (char *)malloc((size_t)rLen + 1);
	for (int i = 0; i < rLen; i++) rf[i] = orc_text_char(idx, rPos + i);
	int32_t L;
	if (qLen == rLen) { /* CheckFragPairMismatch, :49-61: query N positions are skipped */
		int mism = 0;
		for (int i = 0; i < qLen; i++) { int b = nt4(seq[qPos + i]); if (b != 4 && b != nt4(rf[i])) mism++; }
		if (mism <= 5) {
			memcpy(out1, rf, (size_t)rLen); memcpy(out2, seq + qPos, (size_t)qLen);
			out1[rLen] = out2[qLen] = 0; *score_inc = qLen - mism;
			free(rf);
			return qLen;
		}
	}
	*used_dp = 1;
	L = orc_dp_align(rf, rLen, seq + qPos, qLen, out1, out2);
	int same = 0; /* CountIdenticalPairs, :38-47: '-' and N share class 4 */
	for (int i = 0; i < L; i++) if (nt4(out1[i]) == nt4(out2[i])) same++;
	*score_inc = same;
	free(rf);
	return L;
}
