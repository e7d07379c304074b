// capi.cu -- the extern "C" entry points of include/gsalign_b200.h.
#include "gsa_internal.cuh"
#include "dpx.cuh"
#include <stdarg.h>
#include <string.h>
#include <chrono>
#include <algorithm>

int gsa_fail(gsa_ctx *ctx, int code, const char *fmt, ...)
{
	char buf[1024];
	va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
	if (ctx) ctx->err = buf;
	return code;
}

int gsa_ensure(gsa_ctx *ctx, DevBuf &b, size_t bytes)
{
	if (bytes <= b.cap && b.p) return GSA_OK;
	// grow-only, with slack so that similar contigs do not reallocate; the slack is capped: index structures are tens of GB
	size_t want = bytes + std::min<size_t>(bytes / 4, (size_t)256 << 20) + 256;
	if (b.p) { cudaError_t e = cudaFree(b.p); b.p = nullptr; b.cap = 0; if (e != cudaSuccess) return gsa_fail(ctx, GSA_ERR_CUDA, "cudaFree: %s", cudaGetErrorString(e)); }
	cudaError_t e = cudaMalloc(&b.p, want);
	if (e != cudaSuccess) { b.p = nullptr; return gsa_fail(ctx, GSA_ERR_NOMEM, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e)); }
	b.cap = want;
	return GSA_OK;
}

int gsa_ensure_host(gsa_ctx *ctx, HostBuf &b, size_t bytes)
{
	if (bytes <= b.cap && b.p) return GSA_OK;
	size_t want = bytes + bytes / 4 + 256;
	if (b.p) { cudaFreeHost(b.p); b.p = nullptr; b.cap = 0; }
	cudaError_t e = cudaMallocHost(&b.p, want);
	if (e != cudaSuccess) { b.p = nullptr; return gsa_fail(ctx, GSA_ERR_NOMEM, "cudaMallocHost(%zu): %s", want, cudaGetErrorString(e)); }
	b.cap = want;
	return GSA_OK;
}

// ---- small transfers that stay off the copy engines ---------------------------------------------------------------------
// A lane waits on a handful of counters per contig.  As cudaMemcpyAsync calls those 8-byte reads queue on the same copy
// engine as the bulk transfers of the OTHER lanes (a 125 MB contig upload, 100 MB of records coming back) and wait
// milliseconds behind them.  Pinned host memory is mapped into the device's address space, so a few warps store / load
// the words over PCIe directly instead; bulk transfers keep using the copy engines.
#define GSA_SMALL_BYTES ((size_t)256 << 10)
__global__ void k_copy_words(uint32_t *dst, const uint32_t *src, size_t n)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// Bulk transfers between pinned host memory and HBM.  GSA_COPY_CHUNK_MB=<n> queues them in pieces of n MB instead of one
// operation -- an experiment knob: measured at C4 (profiles/r2_summary.md, call N) pieces of 1-4 MB make the end-to-end step
// 6-15 % slower, so the default is one operation per transfer.  (Moving the bulk bytes with a kernel over the mapped host
// pointers instead of a copy engine was tried too: 151 vs 80 ms per step, call X.)
int gsa_bulk_copy(gsa_ctx *ctx, void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t stream)
{
	static const size_t piece = [] { const char *e = getenv("GSA_COPY_CHUNK_MB"); long v = e ? atol(e) : 0; return v <= 0 ? (size_t)0 : (size_t)v << 20; }();
	if (piece == 0 || bytes <= piece) { CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, kind, stream)); return GSA_OK; }
	for (size_t o = 0; o < bytes; o += piece)
		CUDA_TRY(ctx, cudaMemcpyAsync((char *)dst + o, (const char *)src + o, std::min(piece, bytes - o), kind, stream));
	return GSA_OK;
}

static int small_copy(gsa_ctx *ctx, void *dst, const void *src, size_t bytes, bool to_host)
{
	if (bytes == 0) return GSA_OK;
	static const bool use_dma = [] { const char *e = getenv("GSA_SMALL_COPY"); return e && strcmp(e, "dma") == 0; }();
	void *mapped = nullptr;
	void *host = to_host ? dst : const_cast<void *>(src);
	if (!use_dma && bytes <= GSA_SMALL_BYTES && (bytes & 3) == 0 && (((uintptr_t)dst | (uintptr_t)src) & 3) == 0 && cudaHostGetDevicePointer(&mapped, host, 0) == cudaSuccess && mapped) {
		const size_t n = bytes >> 2;
		const unsigned grid = (unsigned)std::min<size_t>(64, (n + 255) / 256);
		if (to_host) k_copy_words<<<grid, 256, 0, ctx->stream>>>((uint32_t *)mapped, (const uint32_t *)src, n);
		else k_copy_words<<<grid, 256, 0, ctx->stream>>>((uint32_t *)dst, (const uint32_t *)mapped, n);
		KERNEL_CHECK(ctx);
		return GSA_OK;
	}
	(void)cudaGetLastError();
	CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice, ctx->stream));
	return GSA_OK;
}
int gsa_small_d2h(gsa_ctx *ctx, void *host_pinned, const void *dev, size_t bytes) { return small_copy(ctx, host_pinned, dev, bytes, true); }

// the first min(*d_count, cap) elements of a device table whose length is only known on the device
__global__ void k_copy_counted(uint32_t *dst, const uint32_t *src, const int32_t *count, int words_per_elem, int cap)
{
	const size_t n = (size_t)max(0, min(*count, cap)) * (size_t)words_per_elem;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
int gsa_small_d2h_counted(gsa_ctx *ctx, void *host_pinned, const void *dev, size_t elem_bytes, const int32_t *d_count, int cap)
{
	void *mapped = nullptr;
	static const bool use_dma = [] { const char *e = getenv("GSA_SMALL_COPY"); return e && strcmp(e, "dma") == 0; }();
	if (!use_dma && (elem_bytes & 3) == 0 && cudaHostGetDevicePointer(&mapped, host_pinned, 0) == cudaSuccess && mapped) {
		k_copy_counted<<<16, 256, 0, ctx->stream>>>((uint32_t *)mapped, (const uint32_t *)dev, d_count, (int)(elem_bytes >> 2), cap);
		KERNEL_CHECK(ctx);
		return GSA_OK;
	}
	(void)cudaGetLastError();
	CUDA_TRY(ctx, cudaMemcpyAsync(host_pinned, dev, elem_bytes * (size_t)cap, cudaMemcpyDeviceToHost, ctx->stream));
	return GSA_OK;
}
int gsa_small_h2d(gsa_ctx *ctx, void *dev, const void *host_pinned, size_t bytes) { return small_copy(ctx, dev, host_pinned, bytes, false); }

extern "C" {

void gsa_default_params(gsa_params *p)
{ // reference src/main.cpp:203-215
	p->min_seed_len = 15; p->sensitive = 0; p->max_indel = 25; p->min_block_score = 200;
	p->min_aln_len = 200; p->min_idy = 70; p->one_on_one = 0;
}

int gsa_create(int device, gsa_ctx **out)
{
	if (!out) return GSA_ERR_ARG;
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
		fprintf(stderr, "gsalign_b200: no usable CUDA device %d (%s); there is no CPU fallback\n", device, e == cudaSuccess ? "bad ordinal" : cudaGetErrorString(e));
		return GSA_ERR_CUDA;
	}
	gsa_ctx *ctx = new gsa_ctx();
	ctx->device = device;
	gsa_default_params(&ctx->prm);
	memset(&ctx->tm, 0, sizeof(ctx->tm));
	memset(&ctx->ix, 0, sizeof(ctx->ix));
	if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return GSA_ERR_CUDA; }
	// K1 reads single random 32-byte sectors of a multi-GB index; the L2 otherwise fetches 64 bytes from DRAM per miss (the
	// neighbour sector is dead weight for a rank block or a suffix-array group).  GSA_L2_FETCH=32|64|128 sets the hint.
	if (const char *g = getenv("GSA_L2_FETCH")) { int v = atoi(g); if (v == 32 || v == 64 || v == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v); }
	for (int i = 0; i < 12; i++) cudaEventCreate(&ctx->ev[i]);
	if (cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
	    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return GSA_ERR_CUDA; }
	ctx->side[0] = ctx->stream2;
	for (int i = 0; i < GSA_NSIDE; i++) {
		if (i > 0 && cudaStreamCreateWithFlags(&ctx->side[i], cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return GSA_ERR_CUDA; }
		if (cudaEventCreateWithFlags(&ctx->ev_side[i], cudaEventDisableTiming) != cudaSuccess) { delete ctx; return GSA_ERR_CUDA; }
	}
	if (gsa_ensure_host(ctx, ctx->h_small, 1 << 20) != GSA_OK) { delete ctx; return GSA_ERR_NOMEM; }
	if (gsa_dpx_init_device(ctx) != GSA_OK) { fprintf(stderr, "gsalign_b200: %s\n", ctx->err.c_str()); delete ctx; return GSA_ERR_CUDA; }
	*out = ctx;
	return GSA_OK;
}

int gsa_create_shared(gsa_ctx *owner, gsa_ctx **out)
{
	if (!out) return GSA_ERR_ARG;
	*out = nullptr;
	if (!owner) return GSA_ERR_ARG;
	if (!owner->have_index) return gsa_fail(owner, GSA_ERR_ARG, "gsa_create_shared: the owner has no index yet");
	CUDA_TRY(owner, cudaSetDevice(owner->device));
	// the k-mer prefix table of the owner's current parameters is shared too (a lane with another seed length builds its own)
	int k = owner->prm.min_seed_len < GSA_KTAB_MAX_K ? owner->prm.min_seed_len : GSA_KTAB_MAX_K;
	GSA_TRY(gsa_impl_build_ktab(owner, k));
	GSA_TRY(gsa_impl_build_kbits(owner, owner->prm.min_seed_len));
	CUDA_TRY(owner, cudaStreamSynchronize(owner->stream));
	gsa_ctx *ctx = nullptr;
	int rc = gsa_create(owner->device, &ctx);
	if (rc != GSA_OK) return rc;
	ctx->prm = owner->prm; ctx->ix = owner->ix; ctx->N = owner->N; ctx->cend = owner->cend;
	ctx->contig_off = owner->contig_off; ctx->contig_len = owner->contig_len;
	ctx->shares_index = true;
	size_t cb = ctx->cend.size() * sizeof(ContigEnd);
	if ((rc = gsa_ensure(ctx, ctx->d_cend, cb ? cb : 16)) != GSA_OK || cudaMemcpy(ctx->d_cend.p, ctx->cend.data(), cb, cudaMemcpyHostToDevice) != cudaSuccess) {
		gsa_destroy(ctx);
		return gsa_fail(owner, GSA_ERR_CUDA, "gsa_create_shared: cannot copy the contig table");
	}
	ctx->have_index = true;
	*out = ctx;
	return GSA_OK;
}

void gsa_destroy(gsa_ctx *ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	DevBuf *bufs[] = {&ctx->d_occ, &ctx->d_txt, &ctx->d_sa, &ctx->d_ktab, &ctx->d_kbits, &ctx->d_cend, &ctx->d_seq, &ctx->d_qpk, &ctx->d_qinv,
	                  &ctx->d_counter, &ctx->d_chain, &ctx->d_sq, &ctx->d_sr, &ctx->d_sl, &ctx->d_cub, &ctx->d_cq, &ctx->d_cr, &ctx->d_cl, &ctx->d_cb,
	                  &ctx->d_frag, &ctx->d_fblk, &ctx->d_aln1, &ctx->d_aln2, &ctx->d_bsum, &ctx->d_var, &ctx->pf[0].buf, &ctx->pf[1].buf};
	for (DevBuf *b : bufs) if (b->p) cudaFree(b->p);
	for (DevBuf &b : ctx->d_tmp) if (b.p) cudaFree(b.p);
	gsa_comm_destroy(ctx);
	DevBuf *gb[] = {&ctx->d_outbox, &ctx->d_sizes, &ctx->d_cfrag, &ctx->d_anchor, &ctx->d_packst};
	for (DevBuf *b : gb) if (b->p) cudaFree(b->p);
	for (DevBuf &b : ctx->d_inbox) if (b.p) cudaFree(b.p);
	if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
	for (auto &p : ctx->pf) if (p.ev) cudaEventDestroy(p.ev);
	if (ctx->ev_gather) cudaEventDestroy(ctx->ev_gather);
	if (ctx->ev_outbox) cudaEventDestroy(ctx->ev_outbox);
	if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
	HostBuf *hb[] = {&ctx->h_small, &ctx->h_stage, &ctx->h_frag, &ctx->h_aln1, &ctx->h_aln2, &ctx->h_blocks, &ctx->h_inbox, &ctx->h_rec, &ctx->h_var, &ctx->h_vrange};
	for (HostBuf *b : hb) if (b->p) cudaFreeHost(b->p);
	for (int i = 0; i < 12; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
	if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
	if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
	for (int i = 0; i < GSA_NSIDE; i++) {
		if (ctx->ev_side[i]) cudaEventDestroy(ctx->ev_side[i]);
		if (i > 0 && ctx->side[i]) { cudaStreamSynchronize(ctx->side[i]); cudaStreamDestroy(ctx->side[i]); }
	}
	if (ctx->stream2) { cudaStreamSynchronize(ctx->stream2); cudaStreamDestroy(ctx->stream2); }
	if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
}

const char *gsa_last_error(const gsa_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int gsa_index_upload(gsa_ctx *ctx, const gsa_index_view *view)
{
	if (!ctx) return GSA_ERR_ARG;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	return gsa_impl_index_upload(ctx, view);
}

int gsa_set_wide_index(gsa_ctx *ctx, int enable)
{
	if (!ctx) return GSA_ERR_ARG;
	ctx->force_wide = enable != 0;
	return GSA_OK;
}

int gsa_index_clone(gsa_ctx *dst, gsa_ctx *src)
{
	if (!dst || !src || dst == src) return GSA_ERR_ARG;
	// the source's prefix table and presence bitmap travel with the copy
	CUDA_TRY(src, cudaSetDevice(src->device));
	int k = src->prm.min_seed_len < GSA_KTAB_MAX_K ? src->prm.min_seed_len : GSA_KTAB_MAX_K;
	GSA_TRY(gsa_impl_build_ktab(src, k));
	GSA_TRY(gsa_impl_build_kbits(src, src->prm.min_seed_len));
	CUDA_TRY(src, cudaStreamSynchronize(src->stream));
	return gsa_impl_index_clone(dst, src);
}

int64_t gsa_index_bytes(const gsa_ctx *ctx)
{
	if (!ctx || !ctx->have_index) return 0;
	return (int64_t)(ctx->d_occ.cap + ctx->d_txt.cap + ctx->d_sa.cap + ctx->d_ktab.cap + ctx->d_kbits.cap);
}

int gsa_set_stream(gsa_ctx *ctx, void *cuda_stream)
{
	if (!ctx) return GSA_ERR_ARG;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
	ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false;
	return GSA_OK;
}

int gsa_set_params(gsa_ctx *ctx, const gsa_params *p)
{
	if (!ctx || !p) return GSA_ERR_ARG;
	if (p->min_seed_len < 10 || p->min_seed_len > 30) return gsa_fail(ctx, GSA_ERR_ARG, "min_seed_len must be 10..30 (reference src/main.cpp:257)");
	if (p->max_indel < 10 || p->max_indel > 100) return gsa_fail(ctx, GSA_ERR_ARG, "max_indel must be 10..100 (reference src/main.cpp:266)");
	ctx->prm = *p;
	return GSA_OK;
}

static int contig_reset(gsa_ctx *ctx, uint32_t len)
{
	if (!ctx->have_index) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_contig_begin: no index uploaded");
	if (len >= 0x7FFFFF00u) return gsa_fail(ctx, GSA_ERR_LIMIT, "gsa_contig_begin: contig longer than 2^31 (positions are int in the reference)");
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	ctx->qlen = len; ctx->have_contig = false; ctx->have_seeds = false; ctx->have_cluster = false;
	ctx->n_seeds = 0; ctx->n_cseeds = 0; ctx->n_frags = 0; ctx->aln_bytes = 0; ctx->dp_timed = false; ctx->have_fill = false; ctx->split_hazard = 0;
	memset(&ctx->tm, 0, sizeof(ctx->tm));
	GSA_TRY(gsa_ensure(ctx, ctx->d_seq, (size_t)len + 64));
	return GSA_OK;
}

// Double buffering of the upload: starts the host-to-device copy of a contig the caller will pass to gsa_contig_begin /
// gsa_align_contig soon, on a copy stream of its own, so that it runs under the kernels of the contig being processed.
// seq must stay valid and unchanged until that call (pinned memory for a truly asynchronous copy).
int gsa_contig_prefetch(gsa_ctx *ctx, const char *seq, uint32_t len)
{
	if (!ctx || !seq || len == 0) return GSA_ERR_ARG;
	if (len >= 0x7FFFFF00u) return gsa_fail(ctx, GSA_ERR_LIMIT, "gsa_contig_prefetch: contig longer than 2^31");
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	if (!ctx->copy_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
	gsa_ctx::Prefetch *slot = nullptr;
	for (auto &p : ctx->pf) if (p.src == seq && p.len == len) return GSA_OK;   // already on its way
	for (auto &p : ctx->pf) if (!p.src && !slot) slot = &p;
	if (!slot) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_contig_prefetch: two uploads are already pending");
	if (!slot->ev) CUDA_TRY(ctx, cudaEventCreateWithFlags(&slot->ev, cudaEventDisableTiming));
	// the slot's buffer last held a contig whose processing has ended (the API is synchronous per context), or an upload
	// that was never claimed: that one may still be running
	CUDA_TRY(ctx, cudaEventSynchronize(slot->ev));
	GSA_TRY(gsa_ensure(ctx, slot->buf, (size_t)len + 64));
	GSA_TRY(gsa_bulk_copy(ctx, slot->buf.p, seq, len, cudaMemcpyHostToDevice, ctx->copy_stream));
	CUDA_TRY(ctx, cudaEventRecord(slot->ev, ctx->copy_stream));
	slot->src = seq; slot->len = len;
	return GSA_OK;
}

int gsa_contig_begin(gsa_ctx *ctx, const char *seq, uint32_t len)
{
	if (!ctx || (!seq && len)) return GSA_ERR_ARG;
	for (auto &p : ctx->pf) {
		if (!p.src || p.src != seq || p.len != len) continue;
		// the upload is already under way: its buffer becomes the contig buffer, the old contig buffer becomes the slot's
		CUDA_TRY(ctx, cudaSetDevice(ctx->device));
		p.src = nullptr;
		std::swap(ctx->d_seq, p.buf);
		GSA_TRY(contig_reset(ctx, len));
		ctx->h_seq = seq;
		CUDA_TRY(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
		CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, p.ev, 0));
		GSA_TRY(gsa_impl_pack_query(ctx));
		CUDA_TRY(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
		ctx->have_contig = true;
		return GSA_OK;
	}
	GSA_TRY(contig_reset(ctx, len));
	ctx->h_seq = seq;
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
	GSA_TRY(gsa_bulk_copy(ctx, ctx->d_seq.p, seq, len, cudaMemcpyHostToDevice, ctx->stream));
	GSA_TRY(gsa_impl_pack_query(ctx));
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
	ctx->have_contig = true;
	return GSA_OK;
}

int gsa_contig_begin_device(gsa_ctx *ctx, const void *dev_seq, uint32_t len)
{
	if (!ctx || (!dev_seq && len)) return GSA_ERR_ARG;
	GSA_TRY(contig_reset(ctx, len));
	ctx->h_seq = nullptr;
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_seq.p, dev_seq, len, cudaMemcpyDeviceToDevice, ctx->stream));
	GSA_TRY(gsa_impl_pack_query(ctx));
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
	ctx->have_contig = true;
	return GSA_OK;
}

int gsa_seed(gsa_ctx *ctx, int64_t *n_seeds)
{
	if (!ctx) return GSA_ERR_ARG;
	if (!ctx->have_contig) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_seed: call gsa_contig_begin first");
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
	GSA_TRY(gsa_impl_seed(ctx));
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
	ctx->tm.n_seeds = ctx->n_seeds;
	if (n_seeds) *n_seeds = ctx->n_seeds;
	return GSA_OK;
}

int gsa_fetch_seeds(gsa_ctx *ctx, int32_t *q, int64_t *r, int32_t *l)
{
	if (!ctx || !ctx->have_seeds) return GSA_ERR_ARG;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	size_t n = (size_t)ctx->n_seeds;
	if (n == 0) return GSA_OK;
	CUDA_TRY(ctx, cudaMemcpy(q, ctx->d_sq.p, n * 4, cudaMemcpyDeviceToHost));
	CUDA_TRY(ctx, cudaMemcpy(r, ctx->d_sr.p, n * 8, cudaMemcpyDeviceToHost));
	CUDA_TRY(ctx, cudaMemcpy(l, ctx->d_sl.p, n * 4, cudaMemcpyDeviceToHost));
	return GSA_OK;
}

int gsa_cluster(gsa_ctx *ctx, int32_t *n_blocks)
{
	if (!ctx) return GSA_ERR_ARG;
	if (!ctx->have_seeds) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_cluster: call gsa_seed first");
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
	GSA_TRY(gsa_impl_cluster(ctx));
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[5], ctx->stream));
	ctx->have_cluster = true;
	if (n_blocks) *n_blocks = (int32_t)ctx->final_blocks.size();
	return GSA_OK;
}

int gsa_fill(gsa_ctx *ctx, gsa_alignment *out)
{
	if (!ctx || !out) return GSA_ERR_ARG;
	if (!ctx->have_cluster) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_fill: call gsa_cluster first");
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[6], ctx->stream));
	ctx->have_fill = false;
	GSA_TRY(gsa_impl_fill(ctx, out));
	ctx->have_fill = true;
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[7], ctx->stream));
	CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev[7]));
	cudaEventElapsedTime(&ctx->tm.h2d_ms, ctx->ev[0], ctx->ev[1]);
	cudaEventElapsedTime(&ctx->tm.seed_ms, ctx->ev[2], ctx->ev[3]);
	cudaEventElapsedTime(&ctx->tm.cluster_ms, ctx->ev[4], ctx->ev[5]);
	cudaEventElapsedTime(&ctx->tm.fill_ms, ctx->ev[6], ctx->ev[7]);
	cudaEventElapsedTime(&ctx->tm.total_ms, ctx->ev[0], ctx->ev[7]);
	if (ctx->qlen > 0) cudaEventElapsedTime(&ctx->tm.k_seed_ms, ctx->ev[8], ctx->ev[9]);
	if (ctx->dp_timed) cudaEventElapsedTime(&ctx->tm.k_dp_ms, ctx->ev[10], ctx->ev[11]);
	return GSA_OK;
}

int gsa_variants(gsa_ctx *ctx, gsa_variant_list *out)
{
	if (!ctx || !out) return GSA_ERR_ARG;
	if (!ctx->have_fill) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_variants: call gsa_fill first");
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	return gsa_impl_variants(ctx, out);
}

int gsa_split_hazard(const gsa_ctx *ctx) { return ctx ? ctx->split_hazard : 0; }

int gsa_set_host_results(gsa_ctx *ctx, int enable)
{
	if (!ctx) return GSA_ERR_ARG;
	ctx->host_results = enable != 0;
	return GSA_OK;
}

int gsa_result_device(gsa_ctx *ctx, gsa_alignment *out)
{
	if (!ctx || !out) return GSA_ERR_ARG;
	if (!ctx->have_cluster) return gsa_fail(ctx, GSA_ERR_ARG, "gsa_result_device: call gsa_fill first");
	memset(out, 0, sizeof(*out));
	out->n_blocks = (int32_t)ctx->out_blocks.size(); out->blocks = ctx->out_blocks.data();
	if (out->n_blocks == 0) return GSA_OK;
	out->n_frags = ctx->n_frags; out->frags = (const gsa_frag *)ctx->d_frag.p;
	out->aln_bytes = ctx->aln_bytes; out->aln1 = (const char *)ctx->d_aln1.p; out->aln2 = (const char *)ctx->d_aln2.p;
	return GSA_OK;
}

int gsa_align_contig(gsa_ctx *ctx, const char *seq, uint32_t len, gsa_alignment *out)
{
	GSA_TRY(gsa_contig_begin(ctx, seq, len));
	GSA_TRY(gsa_seed(ctx, nullptr));
	GSA_TRY(gsa_cluster(ctx, nullptr));
	return gsa_fill(ctx, out);
}

int gsa_get_timing(const gsa_ctx *ctx, gsa_timing *out)
{
	if (!ctx || !out) return GSA_ERR_ARG;
	*out = ctx->tm;
	return GSA_OK;
}

int gsa_dp_batch(gsa_ctx *ctx, int32_t n_pairs, const char *ref, const int64_t *ref_off, const char *qry,
                 const int64_t *qry_off, char *out1, char *out2, int32_t *out_len, float *kernel_ms)
{
	if (!ctx || n_pairs < 0) return GSA_ERR_ARG;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	return gsa_impl_dp_batch(ctx, n_pairs, ref, ref_off, qry, qry_off, out1, out2, out_len, nullptr, kernel_ms);
}

int gsa_dp_batch_identity(gsa_ctx *ctx, int32_t n_pairs, const char *ref, const int64_t *ref_off, const char *qry,
                          const int64_t *qry_off, char *out1, char *out2, int32_t *out_len, int32_t *out_identical)
{
	if (!ctx || n_pairs < 0 || !out_identical) return GSA_ERR_ARG;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	return gsa_impl_dp_batch(ctx, n_pairs, ref, ref_off, qry, qry_off, out1, out2, out_len, out_identical, nullptr);
}

} // extern "C"
