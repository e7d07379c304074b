// gather.cu -- the one collective of the path: finished alignment records of every GPU are packed
// into a per-GPU outbox in HBM and gathered to the root GPU with grouped ncclSend / ncclRecv over NVLink; the root
// owns the emitters (reference: the append-only tail of the contig loop, src/GSAlign.cpp:523-548).  Nothing else
// crosses GPUs.  Two ways to build the communicator: one process per GPU (a 128-byte ncclUniqueId handed round by
// the host however it likes: gsa_comm_init_rank) or one process driving all GPUs (bin/GSAlign -gpus N:
// gsa_comm_init_all).
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the library loads -- and the single-GPU path runs -- where NCCL is
// absent, and inside a process that already holds a libnccl (PyTorch bundles its own) the same copy is used.
//
// Outbox image of one GPU (little-endian, sections padded to 16 bytes), one record per finished contig:
//     int64[4]   {contig index, n_blocks, n_frags, aln_bytes}
//     gsa_block[n_blocks]  gsa_frag[n_frags]  aln1[aln_bytes]  aln2[aln_bytes]
#include "gsa_internal.cuh"
#include "scan.cuh"
#include <dlfcn.h>
#include <thread>
#include <algorithm>
#include <vector>
#include <stdlib.h>
#include <string.h>
#include <mutex>

// ---- the few NCCL entry points used, bound by name -------------------------------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;   // ncclSuccess = 0
enum { NCCL_UINT8 = 1, NCCL_INT64 = 4 }; // ncclDataType_t values (nccl.h): ncclUint8 = 1, ncclInt64 = 4

struct NcclApi {
	void *h = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	std::string err;
};

static NcclApi *nccl_api()
{
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, [] {
		const char *names[] = {"libnccl.so.2", "libnccl.so"};
		for (const char *n : names) if ((api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
		if (!api.h) { api.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
#define BIND(field, sym) do { *(void **)(&api.field) = dlsym(api.h, sym); if (!api.field) { api.err = std::string("libnccl lacks ") + sym; return; } } while (0)
		BIND(GetUniqueId, "ncclGetUniqueId"); BIND(CommInitRank, "ncclCommInitRank"); BIND(CommInitAll, "ncclCommInitAll");
		BIND(CommDestroy, "ncclCommDestroy"); BIND(GroupStart, "ncclGroupStart"); BIND(GroupEnd, "ncclGroupEnd");
		BIND(Send, "ncclSend"); BIND(Recv, "ncclRecv"); BIND(AllGather, "ncclAllGather"); BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
	});
	return &api;
}

#define NCCL_TRY(ctx, api, call)                                                                                   \
	do {                                                                                                           \
		ncclResult_t _r = (call);                                                                                  \
		if (_r != 0) return gsa_fail((ctx), GSA_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, (api)->GetErrorString(_r)); \
	} while (0)

static inline int64_t pad16(int64_t n) { return (n + 15) & ~(int64_t)15; }
static const int64_t REC_HDR = 32;

// ---- the compact form of a record -------------------------------------------------------------------------------------------
// A fragment record is 40 bytes, and a 3 Gbp pair has 54 M of them: they are 85 % of what the gather moves.  Almost all of a
// record is redundant: inside a block every fragment starts where the previous one ends (both coordinates), a seed's other
// fields repeat its length, and a gap fragment's row offset is the running sum of the row slots before it (fill.cu lays the
// rows out by a prefix sum over the fragments; a gapped alignment sits right-aligned in a slot of qLen + rLen).  So a record
// travels as 8 bytes per fragment -- the three lengths and two flags -- plus an ANCHOR {fragment index, rPos, qPos, row base}
// wherever the chain of positions restarts (block starts, anything unexpected) and every 256 fragments, so that the host can
// expand any stretch on its own.  The pack kernel checks every rule it relies on against the real record; a record that breaks
// one (or needs more anchors than were provided for) travels in the plain form.
#define PK_STRIDE 256
#define PK_LEN_BITS 21
#define PK_ALN_BITS 20
struct GsaAnchor { int64_t first; int64_t rPos; int32_t qPos; int32_t pad; int64_t row_base; };
static_assert(sizeof(GsaAnchor) == 32, "anchor layout is part of the image format");
enum { PK_FT_DEL = 1, PK_FT_INS = 2, PK_FT_COPY = 3, PK_FT_DP = 4 };   // gsa_frag::reserved (fill.cu)

__host__ __device__ inline int64_t pk_row_slot(int32_t qLen, int32_t rLen, bool seed, bool gapped)
{
	return seed ? 0 : gapped ? (int64_t)qLen + rLen : qLen == 0 ? rLen : qLen;
}

struct FPack {
	const gsa_frag *frag; uint64_t *cfrag; GsaAnchor *anchor; int32_t *st; int anchor_cap;   // st: {n_anchors, broken rule, n_frags}
	struct Item { gsa_frag f; uint8_t flag; };
	__device__ Item load(int64_t i) const
	{
		Item it; it.f = frag[i];
		bool restart = true;
		if (i > 0) { const gsa_frag p = frag[i - 1]; restart = it.f.qPos != p.qPos + p.qLen || it.f.rPos != p.rPos + p.rLen; }
		it.flag = restart || (i % PK_STRIDE) == 0;
		return it;
	}
	__device__ unsigned long long value(const Item &it) const
	{
		return CH_PACK2(it.flag, pk_row_slot(it.f.qLen, it.f.rLen, it.f.bSeed != 0, it.f.reserved == PK_FT_DP));
	}
	__device__ void emit(int64_t i, unsigned long long excl, const Item &it, bool valid) const
	{
		if (!valid) return;
		const gsa_frag &f = it.f;
		const bool seed = f.bSeed != 0, gapped = f.reserved == PK_FT_DP;
		const int64_t row_base = CH_HI(excl), slot = pk_row_slot(f.qLen, f.rLen, seed, gapped);
		bool ok = f.qLen >= 0 && f.rLen >= 0 && f.aln_len >= 0 && f.qLen < (1 << PK_LEN_BITS) && f.rLen < (1 << PK_LEN_BITS) && f.aln_len < (1 << PK_ALN_BITS) &&
		          (f.bSeed == 0 || f.bSeed == 1) && row_base + slot < 0x7FFFFFFFll;
		if (seed) ok = ok && f.rLen == f.qLen && f.aln_len == f.qLen && f.aln_off == 0 && f.reserved == 0;
		else {
			const int expect = f.qLen == 0 ? PK_FT_DEL : f.rLen == 0 ? PK_FT_INS : gapped ? PK_FT_DP : PK_FT_COPY;
			ok = ok && f.reserved == expect && f.aln_len <= slot && f.aln_off == row_base + (gapped ? slot - f.aln_len : 0);
		}
		if (!ok) atomicOr(st + 1, 1);
		cfrag[i] = (uint64_t)(uint32_t)f.qLen | ((uint64_t)(uint32_t)f.rLen << PK_LEN_BITS) | ((uint64_t)(uint32_t)f.aln_len << (2 * PK_LEN_BITS)) |
		           ((uint64_t)seed << 62) | ((uint64_t)gapped << 63);
		if (it.flag) {
			const int64_t a = CH_LO(excl);
			if (a < anchor_cap) { GsaAnchor x; x.first = i; x.rPos = f.rPos; x.qPos = f.qPos; x.pad = 0; x.row_base = row_base; anchor[a] = x; }
		}
	}
	__device__ void finish(unsigned long long total) const { st[0] = (int32_t)CH_LO(total); }
};

__global__ void k_pack_init(int32_t *st, int32_t n) { if (threadIdx.x == 0) { st[0] = 0; st[1] = 0; st[2] = n; st[3] = 0; } }

// expands fragments [a.first, end) of a compact record from their anchor (host)
static void pk_expand(const uint64_t *cfrag, const GsaAnchor &a, int64_t end, gsa_frag *dst)
{
	int64_t r = a.rPos, row = a.row_base; int32_t q = a.qPos;
	for (int64_t i = a.first; i < end; i++) {
		const uint64_t c = cfrag[i];
		const int32_t qLen = (int32_t)(c & ((1u << PK_LEN_BITS) - 1)), rLen = (int32_t)((c >> PK_LEN_BITS) & ((1u << PK_LEN_BITS) - 1));
		const int32_t alen = (int32_t)((c >> (2 * PK_LEN_BITS)) & ((1u << PK_ALN_BITS) - 1));
		const bool seed = (c >> 62) & 1, gapped = (c >> 63) & 1;
		const int64_t slot = pk_row_slot(qLen, rLen, seed, gapped);
		gsa_frag f;
		f.rPos = r; f.qPos = q; f.qLen = qLen; f.rLen = rLen; f.bSeed = seed ? 1 : 0; f.aln_len = alen;
		f.aln_off = seed ? 0 : row + (gapped ? slot - alen : 0);
		f.reserved = seed ? 0 : qLen == 0 ? PK_FT_DEL : rLen == 0 ? PK_FT_INS : gapped ? PK_FT_DP : PK_FT_COPY;
		dst[i] = f;
		r += rLen; q += qLen; row += slot;
	}
}

static int comm_common_init(gsa_ctx *ctx)
{
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	if (!ctx->comm_stream) {
		int lo = 0, hi = 0;
		cudaDeviceGetStreamPriorityRange(&lo, &hi); // the gather must not queue behind the lanes' compute kernels
		CUDA_TRY(ctx, cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
	}
	if (!ctx->ev_gather) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_gather, cudaEventDisableTiming));
	GSA_TRY(gsa_ensure(ctx, ctx->d_sizes, 16 * 1024));
	return GSA_OK;
}

extern "C" {

int gsa_comm_unique_id(void *id, int32_t id_bytes)
{
	NcclApi *api = nccl_api();
	if (!api->err.empty()) { fprintf(stderr, "gsalign_b200: %s\n", api->err.c_str()); return GSA_ERR_CUDA; }
	if (!id || id_bytes < (int32_t)sizeof(ncclUniqueId)) return GSA_ERR_ARG;
	ncclUniqueId u;
	if (api->GetUniqueId(&u) != 0) return GSA_ERR_CUDA;
	memcpy(id, &u, sizeof(u));
	return GSA_OK;
}

int gsa_comm_init_rank(gsa_ctx *ctx, const void *id, int32_t rank, int32_t n_ranks)
{
	if (!ctx || !id || rank < 0 || rank >= n_ranks || n_ranks > 1024) return GSA_ERR_ARG;
	NcclApi *api = nccl_api();
	if (!api->err.empty()) return gsa_fail(ctx, GSA_ERR_CUDA, "%s", api->err.c_str());
	GSA_TRY(comm_common_init(ctx));
	ncclUniqueId u; memcpy(&u, id, sizeof(u));
	ncclComm_t c = nullptr;
	NCCL_TRY(ctx, api, api->CommInitRank(&c, n_ranks, u, rank));
	ctx->nccl_comm = c; ctx->comm_rank = rank; ctx->comm_size = n_ranks;
	return GSA_OK;
}

int gsa_comm_init_all(gsa_ctx *const *ctxs, int32_t n)
{
	if (!ctxs || n < 1 || n > 64) return GSA_ERR_ARG;
	NcclApi *api = nccl_api();
	if (!api->err.empty()) return gsa_fail(ctxs[0], GSA_ERR_CUDA, "%s", api->err.c_str());
	int devs[64]; ncclComm_t comms[64];
	for (int i = 0; i < n; i++) { if (!ctxs[i]) return GSA_ERR_ARG; devs[i] = ctxs[i]->device; GSA_TRY(comm_common_init(ctxs[i])); }
	NCCL_TRY(ctxs[0], api, api->CommInitAll(comms, n, devs));
	for (int i = 0; i < n; i++) { ctxs[i]->nccl_comm = comms[i]; ctxs[i]->comm_rank = i; ctxs[i]->comm_size = n; }
	return GSA_OK;
}

int gsa_comm_destroy(gsa_ctx *ctx)
{
	if (!ctx) return GSA_ERR_ARG;
	if (ctx->nccl_comm) { cudaSetDevice(ctx->device); nccl_api()->CommDestroy((ncclComm_t)ctx->nccl_comm); ctx->nccl_comm = nullptr; }
	return GSA_OK;
}

int gsa_outbox_reset(gsa_ctx *owner)
{
	if (!owner) return GSA_ERR_ARG;
	std::lock_guard<std::mutex> lk(owner->outbox_mu);
	owner->outbox_used = 0;
	return GSA_OK;
}

int64_t gsa_outbox_bytes(gsa_ctx *owner)
{
	if (!owner) return 0;
	std::lock_guard<std::mutex> lk(owner->outbox_mu);
	return owner->outbox_used;
}

int gsa_outbox_reserve(gsa_ctx *owner, int64_t bytes)
{
	if (!owner || bytes < 0) return GSA_ERR_ARG;
	std::lock_guard<std::mutex> lk(owner->outbox_mu);
	if (owner->outbox_used != 0) return gsa_fail(owner, GSA_ERR_ARG, "gsa_outbox_reserve: outbox in use");
	CUDA_TRY(owner, cudaSetDevice(owner->device));
	return gsa_ensure(owner, owner->d_outbox, (size_t)bytes + 64);
}

// Packs the result of the last gsa_fill() of `lane` (a context on the owner's GPU, possibly the owner itself) into the
// owner's outbox, asynchronously on the lane's stream: the fragment list goes through the pack kernel (compact form) or is
// copied as it is (plain form: GSA_GATHER_RAW=1, or a record the compact form cannot carry), rows are copied device to
// device, the O(#blocks) headers come from the host.
int gsa_outbox_append(gsa_ctx *owner, gsa_ctx *lane, int64_t contig)
{
	if (!owner || !lane || owner->device != lane->device) return GSA_ERR_ARG;
	if (!lane->have_cluster) return gsa_fail(lane, GSA_ERR_ARG, "gsa_outbox_append: call gsa_fill first");
	if (contig < 0) return gsa_fail(lane, GSA_ERR_ARG, "gsa_outbox_append: negative contig index");
	CUDA_TRY(lane, cudaSetDevice(lane->device));
	const int64_t nb = (int64_t)lane->out_blocks.size(), nf = nb ? lane->n_frags : 0, ab = nb ? lane->aln_bytes : 0;
	const char *rawenv = getenv("GSA_GATHER_RAW");
	bool compact = !(rawenv && rawenv[0] == '1') && nf > 0 && ab < (1ll << 30) && lane->qlen < (1u << 29);
	int64_t n_anchor = 0;
	if (compact) { // the fragment list in its compact form, in scratch of the lane; the host learns how many anchors it took
		const int64_t cap = nf / PK_STRIDE + 2 + 4 * (int64_t)lane->final_blocks.size() + 1024;
		const int64_t tiles = chain_tiles(nf);
		GSA_TRY(gsa_ensure(lane, lane->d_cfrag, (size_t)nf * 8));
		GSA_TRY(gsa_ensure(lane, lane->d_anchor, (size_t)cap * sizeof(GsaAnchor)));
		GSA_TRY(gsa_ensure(lane, lane->d_packst, 64 + (size_t)(tiles + 2) * 8));
		int32_t *st = (int32_t *)lane->d_packst.p;
		unsigned long long *chain = (unsigned long long *)((char *)lane->d_packst.p + 64);
		CUDA_TRY(lane, cudaMemsetAsync(chain, 0, (size_t)(tiles + 2) * 8, lane->stream));
		k_pack_init<<<1, 32, 0, lane->stream>>>(st, (int32_t)nf);
		FPack f; f.frag = (const gsa_frag *)lane->d_frag.p; f.cfrag = (uint64_t *)lane->d_cfrag.p; f.anchor = (GsaAnchor *)lane->d_anchor.p; f.st = st; f.anchor_cap = (int)cap;
		ChainState cs; cs.ticket = (unsigned int *)chain; cs.total = nullptr; cs.status = chain + 2;
		k_chain<FPack><<<(unsigned)tiles, CH_THREADS, 0, lane->stream>>>(f, st + 2, cs);
		lane->tm.launches += 2;
		CUDA_TRY(lane, cudaGetLastError());
		GSA_TRY(gsa_ensure_host(lane, lane->h_rec, 64));
		GSA_TRY(gsa_small_d2h(lane, lane->h_rec.p, st, 16));
		CUDA_TRY(lane, cudaStreamSynchronize(lane->stream));
		const int32_t *hs = (const int32_t *)lane->h_rec.p;
		n_anchor = hs[0];
		if (hs[1] != 0 || n_anchor > cap) compact = false;   // a rule of the compact form does not hold for this record: send it as it is
	}
	const int64_t hdr_bytes = compact ? 2 * REC_HDR : REC_HDR;
	const int64_t frag_bytes = compact ? pad16(nf * 8) + n_anchor * (int64_t)sizeof(GsaAnchor) : pad16(nf * (int64_t)sizeof(gsa_frag));
	const int64_t need = hdr_bytes + pad16(nb * (int64_t)sizeof(gsa_block)) + frag_bytes + 2 * pad16(ab);
	// header(s) + block headers travel in one small pinned staging block of the lane
	const size_t hb = (size_t)(hdr_bytes + nb * (int64_t)sizeof(gsa_block));
	GSA_TRY(gsa_ensure_host(lane, lane->h_rec, hb));
	int64_t *hdr = (int64_t *)lane->h_rec.p;
	hdr[0] = compact ? -1 - contig : contig; hdr[1] = nb; hdr[2] = nf; hdr[3] = ab;
	if (compact) { hdr[4] = n_anchor; hdr[5] = hdr[6] = hdr[7] = 0; }
	if (nb) memcpy((char *)lane->h_rec.p + hdr_bytes, lane->out_blocks.data(), (size_t)nb * sizeof(gsa_block));
	{
		// Reserving the range and queueing the copies into it happen under the outbox lock: a lane that has to grow the outbox
		// moves the image after waiting for every lane's last RECORDED append, so no append may sit between its reservation
		// and its event (its copies would land in the buffer that is being freed).
		std::lock_guard<std::mutex> lk(owner->outbox_mu);
		if ((size_t)(owner->outbox_used + need) > owner->d_outbox.cap) {
			// grow: wait for the copies in flight, move the image (rare: capacities are remembered across steps)
			for (cudaEvent_t e : owner->outbox_pending) CUDA_TRY(owner, cudaEventSynchronize(e));
			DevBuf bigger;
			static const size_t slack = getenv("GSA_OUTBOX_SLACK") ? (size_t)atoll(getenv("GSA_OUTBOX_SLACK")) : ((size_t)64 << 20); // tests shrink it to grow often
			GSA_TRY(gsa_ensure(owner, bigger, (size_t)(2 * (owner->outbox_used + need)) + slack));
			if (owner->outbox_used) CUDA_TRY(owner, cudaMemcpy(bigger.p, owner->d_outbox.p, (size_t)owner->outbox_used, cudaMemcpyDeviceToDevice));
			if (owner->d_outbox.p) CUDA_TRY(owner, cudaFree(owner->d_outbox.p));
			owner->d_outbox = bigger;
		}
		const int64_t off = owner->outbox_used;
		owner->outbox_used += need;
		if (!lane->ev_outbox) CUDA_TRY(lane, cudaEventCreateWithFlags(&lane->ev_outbox, cudaEventDisableTiming));
		bool known = false;
		for (cudaEvent_t e : owner->outbox_pending) known |= e == lane->ev_outbox;
		if (!known) owner->outbox_pending.push_back(lane->ev_outbox);
		char *dst = (char *)owner->d_outbox.p + off;
		if (owner->ev_gather) CUDA_TRY(lane, cudaStreamWaitEvent(lane->stream, owner->ev_gather, 0)); // the previous gather has left the outbox
		CUDA_TRY(lane, cudaMemcpyAsync(dst, lane->h_rec.p, hb, cudaMemcpyHostToDevice, lane->stream));
		char *p = dst + hdr_bytes + pad16(nb * (int64_t)sizeof(gsa_block));
		if (compact) {
			CUDA_TRY(lane, cudaMemcpyAsync(p, lane->d_cfrag.p, (size_t)nf * 8, cudaMemcpyDeviceToDevice, lane->stream));
			p += pad16(nf * 8);
			if (n_anchor) CUDA_TRY(lane, cudaMemcpyAsync(p, lane->d_anchor.p, (size_t)n_anchor * sizeof(GsaAnchor), cudaMemcpyDeviceToDevice, lane->stream));
			p += n_anchor * (int64_t)sizeof(GsaAnchor);
		} else {
			if (nf) CUDA_TRY(lane, cudaMemcpyAsync(p, lane->d_frag.p, (size_t)nf * sizeof(gsa_frag), cudaMemcpyDeviceToDevice, lane->stream));
			p += pad16(nf * (int64_t)sizeof(gsa_frag));
		}
		if (ab) CUDA_TRY(lane, cudaMemcpyAsync(p, lane->d_aln1.p, (size_t)ab, cudaMemcpyDeviceToDevice, lane->stream));
		p += pad16(ab);
		if (ab) CUDA_TRY(lane, cudaMemcpyAsync(p, lane->d_aln2.p, (size_t)ab, cudaMemcpyDeviceToDevice, lane->stream));
		CUDA_TRY(lane, cudaEventRecord(lane->ev_outbox, lane->stream));
	}
	// the staging block is reused by the next append of this lane: the H2D above must have left it
	CUDA_TRY(lane, cudaEventSynchronize(lane->ev_outbox));
	return GSA_OK;
}

// the communication stream waits for every lane's last append
static int outbox_join(gsa_ctx *owner)
{
	std::lock_guard<std::mutex> lk(owner->outbox_mu);
	for (cudaEvent_t e : owner->outbox_pending) CUDA_TRY(owner, cudaStreamWaitEvent(owner->comm_stream, e, 0));
	return GSA_OK;
}

static int ensure_inbox(gsa_ctx *root, int n)
{
	if ((int)root->d_inbox.size() < n) { root->d_inbox.resize((size_t)n); root->inbox_bytes.resize((size_t)n, 0); }
	return GSA_OK;
}

// One process per GPU: every rank calls this once per job.  The outbox fill levels travel first (one 8-byte all-gather,
// read back on the communication stream only: the lanes' streams are never drained), then one grouped send/recv batch
// moves the images to the root.
int gsa_gather_records(gsa_ctx *owner, int32_t root)
{
	if (!owner) return GSA_ERR_ARG;
	if (!owner->nccl_comm) return gsa_fail(owner, GSA_ERR_ARG, "gsa_gather_records: call gsa_comm_init_rank first");
	NcclApi *api = nccl_api();
	CUDA_TRY(owner, cudaSetDevice(owner->device));
	const int n = owner->comm_size, me = owner->comm_rank;
	if (root < 0 || root >= n) return GSA_ERR_ARG;
	GSA_TRY(outbox_join(owner));
	int64_t *d_sz = (int64_t *)owner->d_sizes.p;     // [0] = mine, [8 ..] = everybody's
	int64_t *h_sz = (int64_t *)owner->h_small.p + 4096;
	h_sz[0] = gsa_outbox_bytes(owner);
	CUDA_TRY(owner, cudaMemcpyAsync(d_sz, h_sz, 8, cudaMemcpyHostToDevice, owner->comm_stream));
	NCCL_TRY(owner, api, api->AllGather(d_sz, d_sz + 8, 1, NCCL_INT64, (ncclComm_t)owner->nccl_comm, owner->comm_stream));
	CUDA_TRY(owner, cudaMemcpyAsync(h_sz + 8, d_sz + 8, (size_t)n * 8, cudaMemcpyDeviceToHost, owner->comm_stream));
	CUDA_TRY(owner, cudaStreamSynchronize(owner->comm_stream));
	if (me == root) {
		GSA_TRY(ensure_inbox(owner, n));
		for (int r = 0; r < n; r++) {
			owner->inbox_bytes[(size_t)r] = h_sz[8 + r];
			if (r != root && h_sz[8 + r] > 0) GSA_TRY(gsa_ensure(owner, owner->d_inbox[(size_t)r], (size_t)h_sz[8 + r]));
		}
		NCCL_TRY(owner, api, api->GroupStart());
		for (int r = 0; r < n; r++)
			if (r != root && h_sz[8 + r] > 0) NCCL_TRY(owner, api, api->Recv(owner->d_inbox[(size_t)r].p, (size_t)h_sz[8 + r], NCCL_UINT8, r, (ncclComm_t)owner->nccl_comm, owner->comm_stream));
		NCCL_TRY(owner, api, api->GroupEnd());
	} else if (h_sz[8 + me] > 0) {
		NCCL_TRY(owner, api, api->Send(owner->d_outbox.p, (size_t)h_sz[8 + me], NCCL_UINT8, root, (ncclComm_t)owner->nccl_comm, owner->comm_stream));
	}
	CUDA_TRY(owner, cudaEventRecord(owner->ev_gather, owner->comm_stream)); // gsa_gather_wait / the next step's appends order after it
	return GSA_OK;
}

// One process driving all GPUs: the same exchange issued for every rank inside one NCCL group (sizes are known on the host).
int gsa_gather_records_all(gsa_ctx *const *ctxs, int32_t n, int32_t root)
{
	if (!ctxs || n < 1 || root < 0 || root >= n) return GSA_ERR_ARG;
	NcclApi *api = nccl_api();
	gsa_ctx *R = ctxs[root];
	if (!R->nccl_comm) return gsa_fail(R, GSA_ERR_ARG, "gsa_gather_records_all: call gsa_comm_init_all first");
	GSA_TRY(ensure_inbox(R, n));
	for (int r = 0; r < n; r++) {
		CUDA_TRY(ctxs[r], cudaSetDevice(ctxs[r]->device));
		GSA_TRY(outbox_join(ctxs[r]));
		R->inbox_bytes[(size_t)r] = gsa_outbox_bytes(ctxs[r]);
	}
	CUDA_TRY(R, cudaSetDevice(R->device));
	for (int r = 0; r < n; r++) if (r != root && R->inbox_bytes[(size_t)r] > 0) GSA_TRY(gsa_ensure(R, R->d_inbox[(size_t)r], (size_t)R->inbox_bytes[(size_t)r]));
	NCCL_TRY(R, api, api->GroupStart());
	for (int r = 0; r < n; r++) {
		const size_t b = (size_t)R->inbox_bytes[(size_t)r];
		if (r == root || b == 0) continue;
		NCCL_TRY(R, api, api->Send(ctxs[r]->d_outbox.p, b, NCCL_UINT8, root, (ncclComm_t)ctxs[r]->nccl_comm, ctxs[r]->comm_stream));
		NCCL_TRY(R, api, api->Recv(R->d_inbox[(size_t)r].p, b, NCCL_UINT8, r, (ncclComm_t)R->nccl_comm, R->comm_stream));
	}
	NCCL_TRY(R, api, api->GroupEnd());
	for (int r = 0; r < n; r++) { CUDA_TRY(ctxs[r], cudaSetDevice(ctxs[r]->device)); CUDA_TRY(ctxs[r], cudaEventRecord(ctxs[r]->ev_gather, ctxs[r]->comm_stream)); }
	return GSA_OK;
}

// blocks the host until this rank's part of the last gather has run
int gsa_gather_wait(gsa_ctx *owner)
{
	if (!owner || !owner->comm_stream) return GSA_ERR_ARG;
	CUDA_TRY(owner, cudaSetDevice(owner->device));
	CUDA_TRY(owner, cudaStreamSynchronize(owner->comm_stream));
	return GSA_OK;
}

// On the root after a gather: the image of `rank`'s outbox where it arrived (device pointer; the root's own image is its outbox).
int gsa_inbox_device(gsa_ctx *root, int32_t rank, const void **dev_ptr, int64_t *bytes)
{
	if (!root || !dev_ptr || !bytes || rank < 0 || rank >= (int)root->inbox_bytes.size()) return GSA_ERR_ARG;
	*bytes = root->inbox_bytes[(size_t)rank];
	*dev_ptr = rank == root->comm_rank ? root->d_outbox.p : root->d_inbox[(size_t)rank].p;
	return GSA_OK;
}

// ... copied to pinned host memory for the emitters (valid until the next call on this context)
int gsa_inbox_host(gsa_ctx *root, int32_t rank, const void **host_ptr, int64_t *bytes)
{
	const void *d = nullptr;
	GSA_TRY(gsa_inbox_device(root, rank, &d, bytes));
	CUDA_TRY(root, cudaSetDevice(root->device));
	GSA_TRY(gsa_ensure_host(root, root->h_inbox, (size_t)*bytes + 16));
	if (*bytes) CUDA_TRY(root, cudaMemcpyAsync(root->h_inbox.p, d, (size_t)*bytes, cudaMemcpyDeviceToHost, root->comm_stream));
	CUDA_TRY(root, cudaStreamSynchronize(root->comm_stream));
	*host_ptr = root->h_inbox.p;
	return GSA_OK;
}

// Host-side walk over an outbox image: fills *out with pointers into the image for the record at *offset and advances it.
// Returns 1 while there is a record, 0 at the end, < 0 on a malformed image.  For a record in the compact form out->frags
// stays NULL: gsa_record_frags expands its fragment list.
struct RecView { bool compact; int64_t contig, nb, nf, ab, n_anchor; const char *blocks, *frags, *anchors, *aln1, *aln2; int64_t end; };
static int rec_view(const void *image, int64_t bytes, int64_t o, RecView &v)
{
	if (o + REC_HDR > bytes) return GSA_ERR_ARG;
	const char *base = (const char *)image;
	int64_t h[4]; memcpy(h, base + o, sizeof(h)); o += REC_HDR;
	v.compact = h[0] < 0; v.contig = v.compact ? -1 - h[0] : h[0];
	v.nb = h[1]; v.nf = h[2]; v.ab = h[3]; v.n_anchor = 0;
	if (v.nb < 0 || v.nf < 0 || v.ab < 0 || v.nb > bytes || v.nf > bytes || v.ab > bytes) return GSA_ERR_ARG;
	if (v.compact) {
		if (o + REC_HDR > bytes) return GSA_ERR_ARG;
		memcpy(h, base + o, sizeof(h)); o += REC_HDR;
		v.n_anchor = h[0];
		if (v.n_anchor < 0 || v.n_anchor > v.nf || (v.nf > 0 && v.n_anchor == 0)) return GSA_ERR_ARG;
	}
	const int64_t fb = v.compact ? pad16(v.nf * 8) + v.n_anchor * (int64_t)sizeof(GsaAnchor) : pad16(v.nf * (int64_t)sizeof(gsa_frag));
	v.end = o + pad16(v.nb * (int64_t)sizeof(gsa_block)) + fb + 2 * pad16(v.ab);
	if (v.end > bytes) return GSA_ERR_ARG;
	v.blocks = base + o; o += pad16(v.nb * (int64_t)sizeof(gsa_block));
	v.frags = base + o; v.anchors = v.compact ? base + o + pad16(v.nf * 8) : nullptr; o += fb;
	v.aln1 = base + o; o += pad16(v.ab);
	v.aln2 = base + o;
	return GSA_OK;
}

int gsa_record_next(const void *image, int64_t bytes, int64_t *offset, int64_t *contig, gsa_alignment *out)
{
	if (!image || !offset || !contig || !out) return GSA_ERR_ARG;
	if (*offset >= bytes) return 0;
	RecView v;
	if (rec_view(image, bytes, *offset, v) != GSA_OK) return GSA_ERR_ARG;
	memset(out, 0, sizeof(*out));
	*contig = v.contig;
	out->n_blocks = (int32_t)v.nb; out->blocks = (const gsa_block *)v.blocks;
	out->n_frags = v.nf; out->frags = v.compact ? nullptr : (const gsa_frag *)v.frags;
	out->aln_bytes = v.ab; out->aln1 = v.aln1; out->aln2 = v.aln2;
	*offset = v.end;
	return 1;
}

int gsa_record_frags(const void *image, int64_t bytes, int64_t record_offset, gsa_frag *dst, int32_t n_threads)
{
	if (!image || record_offset < 0 || record_offset >= bytes) return GSA_ERR_ARG;
	RecView v;
	if (rec_view(image, bytes, record_offset, v) != GSA_OK) return GSA_ERR_ARG;
	if (v.nf == 0) return GSA_OK;
	if (!dst) return GSA_ERR_ARG;
	if (!v.compact) { memcpy(dst, v.frags, (size_t)v.nf * sizeof(gsa_frag)); return GSA_OK; }
	const uint64_t *cfrag = (const uint64_t *)v.frags;
	const GsaAnchor *an = (const GsaAnchor *)v.anchors;
	// anchors are in fragment order and the first one sits on fragment 0: every stretch between two anchors is independent
	if (an[0].first != 0) return GSA_ERR_ARG;
	for (int64_t k = 0; k < v.n_anchor; k++) {
		const int64_t end = k + 1 < v.n_anchor ? an[k + 1].first : v.nf;
		if (an[k].first < 0 || an[k].first >= end || end > v.nf) return GSA_ERR_ARG;
	}
	const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, v.n_anchor / 64));
	auto work = [&](int t) {
		for (int64_t k = v.n_anchor * t / nt; k < v.n_anchor * (t + 1) / nt; k++) pk_expand(cfrag, an[k], k + 1 < v.n_anchor ? an[k + 1].first : v.nf, dst);
	};
	if (nt == 1) work(0);
	else { std::vector<std::thread> th; for (int t = 0; t < nt; t++) th.emplace_back(work, t); for (auto &x : th) x.join(); }
	return GSA_OK;
}

} // extern "C"
