"""Code-point engine on config 3 (100k UTF-8 line pairs): per-call wall time and, under ncu, the launch list.
    python scripts/utf8_probe.py [pairs]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stringwars_b200 as sw
from stringwars_b200 import datasets

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
a, b = datasets.utf8_line_pairs(n, seed=3)
dev = torch.device("cuda", 0)
pad = lambda x: torch.cat([torch.from_numpy(x.copy()), torch.zeros(64, dtype=torch.uint8)]).to(dev)
ad, bd = pad(a.data), pad(b.data)
ao, bo = torch.from_numpy(a.offsets.astype(np.int64)).to(dev), torch.from_numpy(b.offsets.astype(np.int64)).to(dev)
out = torch.zeros(n, dtype=torch.int64, device=dev)
scope = sw.DeviceScope.gpu_device(0)
for name, eng in (("bytes", sw.LevenshteinDistances(scope)), ("codepoints", sw.LevenshteinDistancesUtf8(scope))):
    for _ in range(3):
        eng.compute_pairs_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), n, out.data_ptr())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        eng.compute_pairs_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), n, out.data_ptr())
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms per call, kernel {eng.last_kernel}")
