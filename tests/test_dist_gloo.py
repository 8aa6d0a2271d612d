"""N>1 host logic on CPU: the NCCL scatter/compute/gather module run over gloo with world_size 2 and 3.
The per-rank compute is injected (here: the CPU oracle, which is allowed inside tests) so that the sharding,
rebasing and ordering logic of stringwars_b200/dist.py is what is under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stringwars_b200 import datasets
from stringwars_b200 import dist as swdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_pairs, ragged, result_path, chunks=None, narrow=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    device = torch.device("cpu")
    a = b = None
    if rank == 0:
        ta, tb = (datasets.words_pairs(n_pairs, seed=5) if ragged else datasets.ascii_pairs(n_pairs, 64, seed=5))
        a = (torch.from_numpy(ta.data.copy()), torch.from_numpy(ta.offsets.astype(np.int64)))
        b = (torch.from_numpy(tb.data.copy()), torch.from_numpy(tb.offsets.astype(np.int64)))

    def compute(a_data, a_offs, b_data, b_offs):
        out = orc.levenshtein_pairs(a_data.numpy(), a_offs.numpy().astype(np.uint64), b_data.numpy(), b_offs.numpy().astype(np.uint64))
        return torch.from_numpy(out.astype(np.int64))

    if chunks is None and not narrow:
        got = swdist.sharded_pairs(a, b, device, compute, n_pairs)
    else:
        # the chunked call shape: several chunks per shard, u32 offsets on received chunks, narrow results
        def fill(a_data, a_offs, b_data, b_offs, out):
            assert a_offs.dtype in (torch.int64, torch.int32)
            out.copy_(compute(a_data, a_offs, b_data, b_offs).to(out.dtype))
        stats = {}
        got = swdist.scatter_compute_gather(a, b, device, fill, n_pairs, torch.int16 if narrow else torch.int64, chunks, stats=stats)
        if rank == 0:
            assert stats["chunks"] == (chunks or 1) and stats["sent_bytes"] >= 0
    if rank == 0:
        want = orc.levenshtein_pairs(ta.data, ta.offsets, tb.data, tb.offsets)
        ok = got is not None and np.array_equal(got.numpy().astype(np.uint64), want)
        open(result_path, "w").write("ok" if ok else "mismatch")
    else:
        assert got is None
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_pairs,ragged", [(2, 1001, False), (2, 257, True), (3, 2, True), (2, 1, False)])
def test_scatter_compute_gather(tmp_path, world, n_pairs, ragged):
    """Non-divisible pair counts, ragged lengths, and shards that end up empty (2 pairs on 3 ranks)."""
    result = tmp_path / "result.txt"
    mp.spawn(_worker, args=(world, _free_port(), n_pairs, ragged, str(result)), nprocs=world, join=True)
    assert result.read_text() == "ok"


@pytest.mark.parametrize("world,n_pairs,ragged,chunks", [(2, 1001, True, 3), (3, 500, False, 4), (2, 5, True, 16), (3, 1, False, 2)])
def test_chunked_scatter_narrow_results(tmp_path, world, n_pairs, ragged, chunks):
    """C chunks per shard (more chunks than pairs included), int32 relative offsets, int16 results."""
    result = tmp_path / "result.txt"
    mp.spawn(_worker, args=(world, _free_port(), n_pairs, ragged, str(result), chunks, True), nprocs=world, join=True)
    assert result.read_text() == "ok"


def test_cut_table_is_monotonic_and_byte_balanced():
    a, b = datasets.utf8_line_pairs(800, 30, 300, seed=4)
    ao, bo = torch.from_numpy(a.offsets.astype(np.int64)), torch.from_numpy(b.offsets.astype(np.int64))
    table = swdist.cut_table(ao, bo, 12).numpy()
    cuts = table[0]
    assert cuts[0] == 0 and cuts[-1] == 800 and (np.diff(cuts) >= 0).all()
    assert np.array_equal(table[1], a.offsets[cuts].astype(np.int64)) and np.array_equal(table[2], b.offsets[cuts].astype(np.int64))
    shares = np.diff(table[1] + table[2])
    assert shares.max() - shares.min() <= 2 * 600


def test_shard_bounds_balance_bytes():
    a, b = datasets.utf8_line_pairs(500, 30, 300, seed=3)
    cuts = swdist.shard_bounds(a.offsets, b.offsets, 4)
    assert cuts[0] == 0 and cuts[-1] == 500 and all(x <= y for x, y in zip(cuts, cuts[1:]))
    total = (a.offsets + b.offsets).astype(np.int64)
    shares = [int(total[cuts[r + 1]] - total[cuts[r]]) for r in range(4)]
    assert max(shares) - min(shares) <= 2 * 600  # within one pair of each other
