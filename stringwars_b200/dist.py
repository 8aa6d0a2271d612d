"""Multi-GPU sharding of a pair batch: one process per GPU, `torch.distributed` for the plumbing.

BASELINE.json's north_star: "the batch of pairs shards trivially across the 8 GPUs of one box with a single
NCCL scatter/gather over NVLink". Pairs are independent, so there is no data-path collective inside the compute; the
only exchange is (i) the root scattering each rank's slice of the two tapes and (ii) gathering the results back.

`scatter_compute_gather` is that exchange for one process per GPU (the single-process form lives inside the C ABI,
csrc/multi_gpu.cu, and follows the same plan):
  * the root cuts the pair range into world x C byte-balanced parts on the device (prefix sums + searchsorted) and
    broadcasts ONE fixed-size header (the cut table); that header is the only host round trip of a step — there is no
    per-rank size exchange and no all_gather of bounds;
  * every shard travels in C chunks, each chunk one grouped send/recv (`batch_isend_irecv` = one ncclGroupStart/End)
    of the tape bytes and of u32 offsets rebuilt relative to the shard's first byte (4 B per string on the wire, not 8);
  * receivers post every chunk's receives up front and start a chunk's kernels as soon as it has landed (stream-ordered
    wait), so compute on chunk c overlaps the transfer of chunk c+1; the root's own shard never moves;
  * results come back in ONE grouped send/recv, in the narrow dtype the caller chose (u16 distances: 2 B per pair).
The functions take device tensors and a `compute` callable so the host logic is exercised on CPU with the gloo backend
(tests/test_dist_gloo.py).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

MAX_CHUNKS = 16
TARGET_CHUNK_BYTES = 64 << 20


def shard_bounds(offsets_a, offsets_b, world: int) -> List[int]:
    """Contiguous pair ranges with balanced BYTES (bytes track cells for similar-length pairs).
    Returns world+1 cut points; shards may be empty. Accepts numpy arrays or torch tensors; with device tensors
    the prefix sums and the search stay on the device and only the world-1 cut points cross to the host."""
    if not torch.is_tensor(offsets_a):
        offsets_a = torch.from_numpy(np.ascontiguousarray(offsets_a).astype(np.int64))
        offsets_b = torch.from_numpy(np.ascontiguousarray(offsets_b).astype(np.int64))
    n = offsets_a.numel() - 1
    total = (offsets_a - offsets_a[0]) + (offsets_b - offsets_b[0])
    targets = (total[-1] * torch.arange(1, world, device=total.device, dtype=torch.int64)) // world
    found = torch.searchsorted(total, targets, right=False).tolist() if world > 1 else []
    cuts = [0]
    for c in found:
        cuts.append(max(cuts[-1], min(int(c), n)))
    cuts.append(n)
    return cuts


def cut_table(offsets_a: torch.Tensor, offsets_b: torch.Tensor, parts: int) -> torch.Tensor:
    """Device tensor of 3 x (parts + 1) int64: first pair of every part (equal BYTES per part), then the A and B offsets
    at those pairs. Everything stays on the device."""
    n = offsets_a.numel() - 1
    total = (offsets_a - offsets_a[0]) + (offsets_b - offsets_b[0])
    targets = (total[-1] // parts) * torch.arange(0, parts, device=total.device, dtype=torch.int64)
    cuts = torch.searchsorted(total, targets, right=False).clamp_(max=n)
    cuts = torch.cummax(cuts, dim=0).values
    cuts = torch.cat([cuts, torch.tensor([n], device=total.device, dtype=torch.int64)])
    cuts[0] = 0
    return torch.stack([cuts, offsets_a[cuts], offsets_b[cuts]])


def chunks_for(total_bytes: int, world: int) -> int:
    return int(min(MAX_CHUNKS, max(1, total_bytes // max(world, 1) // TARGET_CHUNK_BYTES)))


def _wire(t: torch.Tensor) -> torch.Tensor:
    """What travels: the tensor's bytes (the NCCL process group has no int16, and a byte count is what a scatter is)."""
    return t.view(torch.uint8) if t.dtype != torch.uint8 else t


ComputeFn = Callable[[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor], None]


def scatter_compute_gather(a: Optional[Tuple[torch.Tensor, torch.Tensor]], b: Optional[Tuple[torch.Tensor, torch.Tensor]],
                           device: torch.device, compute: ComputeFn, total_pairs: int, result_dtype: torch.dtype = torch.int64,
                           chunks: Optional[int] = None, src: int = 0, group=None, stats: Optional[dict] = None) -> Optional[torch.Tensor]:
    """scatter -> per-rank compute -> gather.

    Rank `src` passes the tapes `(data u8, offsets int64)` as tensors on `device`, everyone else None.
    `compute(a_data, a_offsets, b_data, b_offsets, out)` fills `out[i]` for the len(a_offsets) - 1 pairs it is given
    (offsets are int64 on the root's own shard and int32 — relative to a_data / b_data — on received chunks).
    Returns the full result tensor (dtype `result_dtype`, input order) on `src`, None elsewhere."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    max_parts = world * MAX_CHUNKS
    header = torch.zeros(1 + 3 * (max_parts + 1), dtype=torch.int64, device=device)
    if rank == src:
        a_data, a_offs = a
        b_data, b_offs = b
        if chunks is None:
            ends = torch.stack([a_offs[0], a_offs[-1], b_offs[0], b_offs[-1]]).tolist()
            chunks = chunks_for((ends[1] - ends[0]) + (ends[3] - ends[2]), world)
        chunks = int(max(1, min(MAX_CHUNKS, chunks)))
        parts = world * chunks
        table = cut_table(a_offs, b_offs, parts)
        header[0] = chunks
        for row in range(3):
            header[1 + row * (max_parts + 1): 1 + row * (max_parts + 1) + parts + 1] = table[row]
    if world > 1:
        dist.broadcast(header, src=src, group=group)
    host = header.tolist()  # the step's one host round trip
    C = int(host[0])
    parts = world * C
    cut = host[1: 1 + parts + 1]
    a_at = host[1 + (max_parts + 1): 1 + (max_parts + 1) + parts + 1]
    b_at = host[1 + 2 * (max_parts + 1): 1 + 2 * (max_parts + 1) + parts + 1]
    lo, hi = cut[rank * C], cut[(rank + 1) * C]
    sent = 0

    if rank == src:
        out = torch.empty(total_pairs, dtype=result_dtype, device=device)
        keep, send_reqs = [], []
        for c in range(C):
            ops = []
            for r in range(world):
                if r == src:
                    continue
                k, k0 = r * C + c, r * C
                if cut[k + 1] == cut[k]:
                    continue
                pieces = [a_data[a_at[k]: a_at[k + 1]], b_data[b_at[k]: b_at[k + 1]],
                          (a_offs[cut[k]: cut[k + 1] + 1] - a_at[k0]).to(torch.int32), (b_offs[cut[k]: cut[k + 1] + 1] - b_at[k0]).to(torch.int32)]
                for t in pieces:
                    if t.numel():
                        ops.append(dist.P2POp(dist.isend, _wire(t), r, group))
                        sent += t.numel() * t.element_size()
                keep.append(pieces)
            if ops:
                send_reqs += dist.batch_isend_irecv(ops)
        if hi > lo:
            compute(a_data, a_offs[lo: hi + 1], b_data, b_offs[lo: hi + 1], out[lo:hi])
        for req in send_reqs:
            req.wait()
        ops = []
        for r in range(world):
            rlo, rhi = cut[r * C], cut[(r + 1) * C]
            if r != src and rhi > rlo:
                ops.append(dist.P2POp(dist.irecv, _wire(out[rlo:rhi]), r, group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if stats is not None:
            stats.update(sent_bytes=sent, gathered_bytes=(total_pairs - (hi - lo)) * out.element_size(), chunks=C)
        del keep
        return out

    # receivers: post every chunk's receives, then compute chunk by chunk as the data lands
    k0 = rank * C
    a_buf = torch.empty(a_at[k0 + C] - a_at[k0] + 64, dtype=torch.uint8, device=device)
    b_buf = torch.empty(b_at[k0 + C] - b_at[k0] + 64, dtype=torch.uint8, device=device)
    local = torch.empty(hi - lo, dtype=result_dtype, device=device)
    chunk_reqs, chunk_offs = [], []
    for c in range(C):
        k = k0 + c
        cnt = cut[k + 1] - cut[k]
        if cnt == 0:
            chunk_reqs.append([]), chunk_offs.append(None)
            continue
        ao = torch.empty(cnt + 1, dtype=torch.int32, device=device)
        bo = torch.empty(cnt + 1, dtype=torch.int32, device=device)
        pieces = [a_buf[a_at[k] - a_at[k0]: a_at[k + 1] - a_at[k0]], b_buf[b_at[k] - b_at[k0]: b_at[k + 1] - b_at[k0]], ao, bo]
        ops = [dist.P2POp(dist.irecv, _wire(t), src, group) for t in pieces if t.numel()]
        chunk_reqs.append(dist.batch_isend_irecv(ops) if ops else [])
        chunk_offs.append((ao, bo))
    for c in range(C):
        k = k0 + c
        if cut[k + 1] == cut[k]:
            continue
        for req in chunk_reqs[c]:
            req.wait()
        ao, bo = chunk_offs[c]
        compute(a_buf, ao, b_buf, bo, local[cut[k] - lo: cut[k + 1] - lo])
    if hi > lo:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, _wire(local), src, group)]):
            req.wait()
    return None


def sharded_pairs(a, b, device: torch.device, compute: Callable[[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor], torch.Tensor],
                  total_pairs: int, src: int = 0, group=None) -> Optional[torch.Tensor]:
    """Round-1 call shape kept for callers that return a fresh int64 tensor per shard:
    `compute(a_data, a_offs, b_data, b_offs) -> int64[n_local]`."""
    def fill(a_data, a_offs, b_data, b_offs, out):
        out.copy_(compute(a_data, a_offs, b_data, b_offs))
    return scatter_compute_gather(a, b, device, fill, total_pairs, torch.int64, None, src, group)
