"""The one collective of the path: finished alignment records of every rank are packed into a
per-rank outbox (device memory) and gathered to rank 0, which owns the emitters.  Nothing else crosses GPUs.

Record layout inside an outbox (all little-endian, 16-byte aligned sections):
    header  int64[4]  = {contig index, n_blocks, n_frags, aln_bytes}
    blocks  gsa_block[n_blocks]   (24 bytes each, include/gsalign_b200.h)
    frags   gsa_frag[n_frags]     (40 bytes each)
    aln1    char[aln_bytes]       reference rows of the gap fragments
    aln2    char[aln_bytes]       query rows
The gather is ncclSend/ncclRecv grouped into one batch (torch.distributed.batch_isend_irecv) with exact sizes, after a
tiny all_gather of the outbox fill levels; the same code runs on gloo/CPU tensors in the tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

HEADER_WORDS = 4
BLOCK_BYTES = 24
FRAG_BYTES = 40


def _pad16(n: int) -> int:
    return (n + 15) & ~15


def record_bytes(n_blocks: int, n_frags: int, aln_bytes: int) -> int:
    return 8 * HEADER_WORDS + _pad16(n_blocks * BLOCK_BYTES) + _pad16(n_frags * FRAG_BYTES) + 2 * _pad16(aln_bytes)


class Outbox:
    """Append-only byte buffer on one device; lanes reserve space under a lock and fill it on their own stream."""

    def __init__(self, capacity: int, device):
        self.buf = torch.empty(capacity, dtype=torch.uint8, device=device)
        self.used = 0

    def reset(self):
        self.used = 0

    def reserve(self, nbytes: int) -> int:
        off = self.used
        if off + nbytes > self.buf.numel():
            raise RuntimeError(f"outbox overflow: {off + nbytes} > {self.buf.numel()} bytes")
        self.used = off + nbytes
        return off

    def put(self, off: int, contig: int, blocks_u8: torch.Tensor, frags_u8: torch.Tensor, aln1_u8: torch.Tensor, aln2_u8: torch.Tensor):
        """copies one contig's record to [off, off + record_bytes) (asynchronously on the current stream)"""
        n_blocks, n_frags, aln_bytes = blocks_u8.numel() // BLOCK_BYTES, frags_u8.numel() // FRAG_BYTES, aln1_u8.numel()
        hdr = torch.tensor([contig, n_blocks, n_frags, aln_bytes], dtype=torch.int64)
        o = off
        self.buf[o:o + 8 * HEADER_WORDS].copy_(hdr.view(torch.uint8), non_blocking=True); o += 8 * HEADER_WORDS
        self.buf[o:o + blocks_u8.numel()].copy_(blocks_u8, non_blocking=True); o += _pad16(blocks_u8.numel())
        self.buf[o:o + frags_u8.numel()].copy_(frags_u8, non_blocking=True); o += _pad16(frags_u8.numel())
        self.buf[o:o + aln_bytes].copy_(aln1_u8, non_blocking=True); o += _pad16(aln_bytes)
        self.buf[o:o + aln_bytes].copy_(aln2_u8, non_blocking=True)


def gather_to_root(outbox: torch.Tensor, used: int, root: int = 0, group=None):
    """All ranks call this once per job.  Returns on root a list (one entry per rank) of uint8 tensors holding that rank's
    outbox[:used]; elsewhere None.  One size all_gather + one grouped send/recv batch."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes_t = torch.zeros(world, dtype=torch.int64, device=outbox.device)
    dist.all_gather_into_tensor(sizes_t, torch.tensor([used], dtype=torch.int64, device=outbox.device), group=group)
    sizes = [int(x) for x in sizes_t.cpu().tolist()]   # one host sync
    if rank == root:
        inbox = [outbox[:used] if r == root else torch.empty(sizes[r], dtype=torch.uint8, device=outbox.device) for r in range(world)]
        ops = [dist.P2POp(dist.irecv, inbox[r], r, group=group) for r in range(world) if r != root and sizes[r] > 0]
    else:
        inbox = None
        ops = [dist.P2POp(dist.isend, outbox[:used], root, group=group)] if used > 0 else []
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return inbox


def unpack(buf: torch.Tensor):
    """Splits one rank's outbox into records: list of (contig, blocks u8 array, frags u8 array, aln1 bytes, aln2 bytes) (host)"""
    a = buf.cpu().numpy()
    out, o = [], 0
    while o < a.shape[0]:
        contig, n_blocks, n_frags, aln_bytes = (int(x) for x in a[o:o + 8 * HEADER_WORDS].view(np.int64)); o += 8 * HEADER_WORDS
        blocks = a[o:o + n_blocks * BLOCK_BYTES]; o += _pad16(n_blocks * BLOCK_BYTES)
        frags = a[o:o + n_frags * FRAG_BYTES]; o += _pad16(n_frags * FRAG_BYTES)
        aln1 = a[o:o + aln_bytes]; o += _pad16(aln_bytes)
        aln2 = a[o:o + aln_bytes]; o += _pad16(aln_bytes)
        out.append((contig, blocks, frags, aln1, aln2))
    return out
