"""On-GPU probe: thread-per-pair bit-vector kernels on strings that are NOT 16-byte aligned (odd fixed lengths and
ragged lengths), kernel-only, parity-checked against the oracle. Not a bench."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import stringwars_b200 as sw
from stringwars_b200 import datasets
from oracle import oracle as orc

scope = sw.DeviceScope.gpu_device(0)
eng = sw.LevenshteinDistances(scope)
dev = torch.device("cuda", 0)

def to_dev(t):
    pad = torch.cat([torch.from_numpy(np.array(t.data)), torch.zeros(64, dtype=torch.uint8)]).to(dev)
    return pad, torch.from_numpy(np.array(t.offsets).view(np.int64)).to(dev)

def run(name, a, b, reps=5):
    ad, ao = to_dev(a); bd, bo = to_dev(b)
    out = torch.zeros(len(a), dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        eng.compute_pairs_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), len(a), out.data_ptr(), stream=st)
    torch.cuda.synchronize()
    eng.set_profiling(True)
    for _ in range(reps):
        eng.compute_pairs_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), len(a), out.data_ptr(), stream=st)
    ms, n = eng.kernel_time()
    eng.set_profiling(False)
    cells, _ = datasets.cells_and_bytes(a, b)
    sample = min(len(a), 20000)
    want = orc.levenshtein_pairs(a.data, a.offsets[:sample + 1], b.data, b.offsets[:sample + 1], orc.RAPIDFUZZ_CLASS, threads=16)
    ok = bool(np.array_equal(out.cpu().numpy().view(np.uint64)[:sample], want))
    print(f"{name:<22} {eng.last_kernel:<22} {cells / (ms / reps) / 1e6:9.0f} GCUPS  launches/call {n / reps:.0f}  parity {ok}", flush=True)

rng = np.random.default_rng(7)
alphabet = np.arange(0x20, 0x7F, dtype=np.uint8)
run("ascii 63x63", *datasets.ascii_pairs(1 << 19, 63))
run("ascii 61x64", datasets.ascii_pairs(1 << 19, 61)[0], datasets.ascii_pairs(1 << 19, 64)[1])
run("ascii 250x250", *datasets.ascii_pairs(1 << 17, 250))
run("ragged 33..64", datasets._ragged(rng.integers(33, 65, 1 << 19), alphabet, rng, False), datasets._ragged(rng.integers(33, 65, 1 << 19), alphabet, rng, False))
run("ragged 129..256", datasets._ragged(rng.integers(129, 257, 1 << 17), alphabet, rng, False), datasets._ragged(rng.integers(129, 257, 1 << 17), alphabet, rng, False))
