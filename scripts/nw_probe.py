"""Config 5 (10k ACGT pairs x 1024, NW scores) kernel timing, linear and affine gaps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import stringwars_b200 as sw
from stringwars_b200 import datasets
scope = sw.DeviceScope.gpu_device(0)
btc, costs = sw.unary_class_costs(2, -1)
a, b = datasets.acgt_pairs(10_000, 1024)
dev = torch.device("cuda", 0)
pad = lambda t: torch.cat([torch.from_numpy(t.data.copy()), torch.zeros(64, dtype=torch.uint8)]).to(dev)
ad, bd = pad(a), pad(b)
ao, bo = torch.from_numpy(a.offsets.view(np.int64).copy()).to(dev), torch.from_numpy(b.offsets.view(np.int64).copy()).to(dev)
out = torch.zeros(len(a), dtype=torch.int64, device=dev)
cells = 10_000 * 1024 * 1024
for name, open_, ext, cls in (("nw_linear", -2, -2, sw.NeedlemanWunschScores), ("nw_affine", -5, -1, sw.NeedlemanWunschScores),
                              ("sw_linear", -2, -2, sw.SmithWatermanScores), ("sw_affine", -5, -1, sw.SmithWatermanScores)):
    eng = cls(scope, btc, costs, open_, ext)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        eng.compute_pairs_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), len(a), out.data_ptr(), stream=st)
    eng.set_profiling(True)
    for _ in range(3):
        eng.compute_pairs_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), len(a), out.data_ptr(), stream=st)
    ms, n = eng.kernel_time()
    print(f"{name}: {eng.last_kernel} {ms / n:.3f} ms  {cells / (ms / n) / 1e6:.0f} GCUPS  checksum {int(out.sum().item())}", flush=True)
