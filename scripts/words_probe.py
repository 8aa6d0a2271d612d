"""Config 1 (words <= 16 bytes): the register kernel on 4M element-wise pairs and the cross kernel at the harness's side 1557.
    python scripts/words_probe.py"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stringwars_b200 as sw
from stringwars_b200 import datasets

dev = torch.device("cuda", 0)
scope = sw.DeviceScope.gpu_device(0)
eng = sw.LevenshteinDistances(scope)
n = 1 << 22
a, b = datasets.words_pairs(n, seed=7)
pad = lambda x: torch.cat([torch.from_numpy(x.copy()), torch.zeros(64, dtype=torch.uint8)]).to(dev)
ad, bd = pad(a.data), pad(b.data)
ao, bo = torch.from_numpy(a.offsets.astype(np.int64)).to(dev), torch.from_numpy(b.offsets.astype(np.int64)).to(dev)
out = torch.zeros(n, dtype=torch.int64, device=dev)
eng.set_profiling(True)
for i in range(8):
    eng.compute_pairs_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), n, out.data_ptr())
    if i == 2:
        eng.kernel_time()
ms, k = eng.kernel_time()
print(f"pairs: {eng.last_kernel}: {ms / k:.4f} ms per launch, {n / (ms / k) / 1e6:.1f} G pairs/s")
eng.set_profiling(False)
side = 1557
q, c = datasets.words_pairs(side, seed=13)
tq, tc = sw.BytesTape.from_arrays(q.data, q.offsets, memory="managed"), sw.BytesTape.from_arrays(c.data, c.offsets, memory="managed")
m = sw.unified_mat(side, side, np.uint64, memory="managed")
for _ in range(5):
    eng.compute_into(scope, tq, tc, m)
t0 = time.perf_counter()
for _ in range(200):
    eng.compute_into(scope, tq, tc, m)
dt = (time.perf_counter() - t0) / 200
cells = float(np.outer(q.lengths().astype(np.float64), c.lengths().astype(np.float64)).sum())
print(f"cross side {side}: {eng.last_kernel}: {dt * 1e6:.1f} us per call, {cells / dt / 1e9:.0f} GCUPS")
