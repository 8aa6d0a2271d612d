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
        cuts = shard_bounds(a_offs, b_offs, world)
        idx = torch.tensor(cuts, device=a_offs.device, dtype=torch.int64)
        a_bytes, b_bytes = a_offs[idx].tolist(), b_offs[idx].tolist()  # one small readback for all cut points
        shards, sizes = [], []
        for r in range(world):
            lo, hi = cuts[r], cuts[r + 1]
            ad, bd = a_data[a_bytes[r]: a_bytes[r + 1]], b_data[b_bytes[r]: b_bytes[r + 1]]
            ao, bo = a_offs[lo: hi + 1] - a_bytes[r], b_offs[lo: hi + 1] - b_bytes[r]
            shards.append((ad, ao, bd, bo))
            sizes += [ad.numel(), ao.numel(), bd.numel(), bo.numel()]
        meta.copy_(torch.tensor(sizes + cuts, dtype=torch.int64))
    dist.broadcast(meta, src=src, group=group)
    meta_host = meta.tolist()
    sizes = meta_host[4 * rank: 4 * rank + 4]
    cuts = meta_host[4 * world:]
    lo, hi = cuts[rank], cuts[rank + 1]
    if rank == src:
        mine = shards[src]
        ops = []
        for r in range(world):
            if r == src:
                continue
            for t in shards[r]:
                if t.numel():
                    ops.append(dist.P2POp(dist.isend, t.contiguous(), r, group))
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
