"""Multi-GPU sharding of a pair batch: one process per GPU, `torch.distributed` for the plumbing.

BASELINE.json's north_star: "the batch of pairs shards trivially across the 8 GPUs of one box with a single
NCCL scatter/gather over NVLink". Pairs are independent, so there is no data-path collective inside the
compute; the only exchange is (i) rank 0 scattering each rank's slice of the two tapes and (ii) gathering the
u64 results back, both as single grouped NCCL operations (`batch_isend_irecv` = one ncclGroupStart/End).
The functions take the process group's device tensors and a `compute` callable so the host logic can be
exercised on CPU with the gloo backend (tests/test_dist_gloo.py).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(offsets_a: np.ndarray, offsets_b: np.ndarray, world: int) -> List[int]:
    """Contiguous pair ranges with balanced BYTES (bytes track cells for similar-length pairs).
    Returns world+1 cut points; shards may be empty."""
    n = len(offsets_a) - 1
    total = (offsets_a - offsets_a[0]).astype(np.uint64) + (offsets_b - offsets_b[0]).astype(np.uint64)
    cuts = [0]
    for r in range(1, world):
        target = int(total[-1]) * r // world
        cuts.append(max(cuts[-1], int(np.searchsorted(total, target, side="left"))))
    cuts.append(n)
    return cuts


def _slice_tape(data: torch.Tensor, offsets: torch.Tensor, lo: int, hi: int) -> Tuple[torch.Tensor, torch.Tensor]:
    o = offsets[lo: hi + 1]
    base = int(o[0]) if hi >= lo else 0
    end = int(o[-1]) if hi >= lo else 0
    return data[base:end], (o - o[0]) if o.numel() else o


def scatter_pairs(a: Optional[Tuple[torch.Tensor, torch.Tensor]], b: Optional[Tuple[torch.Tensor, torch.Tensor]],
                  device: torch.device, src: int = 0, group=None):
    """Rank `src` holds tapes (data u8, offsets int64 as torch tensors on `device`); every rank returns its
    shard `(a_data, a_offsets, b_data, b_offsets, lo, hi)` with offsets rebased to 0. One metadata broadcast,
    then ONE grouped send/recv carrying the four buffers of every shard."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    meta = torch.zeros(world * 4 + world + 1, dtype=torch.int64, device=device)
    shards = None
    if rank == src:
        a_data, a_offs = a
        b_data, b_offs = b
        cuts = shard_bounds(a_offs.cpu().numpy().astype(np.uint64), b_offs.cpu().numpy().astype(np.uint64), world)
        shards = []
        for r in range(world):
            ad, ao = _slice_tape(a_data, a_offs, cuts[r], cuts[r + 1])
            bd, bo = _slice_tape(b_data, b_offs, cuts[r], cuts[r + 1])
            shards.append((ad.contiguous(), ao.contiguous(), bd.contiguous(), bo.contiguous()))
            meta[4 * r: 4 * r + 4] = torch.tensor([ad.numel(), ao.numel(), bd.numel(), bo.numel()])
        meta[4 * world:] = torch.tensor(cuts)
    dist.broadcast(meta, src=src, group=group)
    sizes = meta[4 * rank: 4 * rank + 4].tolist()
    cuts = meta[4 * world:].tolist()
    lo, hi = cuts[rank], cuts[rank + 1]
    if rank == src:
        mine = shards[src]
        ops = []
        for r in range(world):
            if r == src:
                continue
            for t in shards[r]:
                if t.numel():
                    ops.append(dist.P2POp(dist.isend, t, r, group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return (*mine, lo, hi)
    bufs = [torch.empty(sizes[0], dtype=torch.uint8, device=device), torch.empty(sizes[1], dtype=torch.int64, device=device),
            torch.empty(sizes[2], dtype=torch.uint8, device=device), torch.empty(sizes[3], dtype=torch.int64, device=device)]
    ops = [dist.P2POp(dist.irecv, t, src, group) for t in bufs if t.numel()]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return (*bufs, lo, hi)


def gather_results(local: torch.Tensor, cuts_lo_hi: Tuple[int, int], total: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Gather variable-size result shards (int64/u64 viewed as int64) to `dst` in input order with one grouped
    send/recv. Returns the full tensor on `dst`, None elsewhere."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = cuts_lo_hi
    bounds = torch.tensor([lo, hi], dtype=torch.int64, device=local.device)
    all_bounds = [torch.zeros_like(bounds) for _ in range(world)]
    dist.all_gather(all_bounds, bounds, group=group)
    if rank == dst:
        out = torch.empty(total, dtype=local.dtype, device=local.device)
        out[lo:hi] = local
        ops = []
        for r in range(world):
            rlo, rhi = all_bounds[r].tolist()
            if r != dst and rhi > rlo:
                ops.append(dist.P2POp(dist.irecv, out[rlo:rhi], r, group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out
    if hi > lo:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), dst, group)]):
            req.wait()
    return None


def sharded_pairs(a, b, device: torch.device, compute: Callable[[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor], torch.Tensor],
                  total_pairs: int, src: int = 0, group=None) -> Optional[torch.Tensor]:
    """scatter -> per-rank compute -> gather. `compute(a_data, a_offs, b_data, b_offs) -> int64[n_local]`."""
    a_data, a_offs, b_data, b_offs, lo, hi = scatter_pairs(a, b, device, src, group)
    local = compute(a_data, a_offs, b_data, b_offs) if hi > lo else torch.empty(0, dtype=torch.int64, device=device)
    return gather_results(local, (lo, hi), total_pairs, src, group)
