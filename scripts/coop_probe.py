"""Lane-group bit-vector kernel only: 10k ACGT pairs of 1024 (or argv[1]) symbols, kernel-only timing. Not a bench."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import stringwars_b200 as sw
from stringwars_b200 import datasets
L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
scope = sw.DeviceScope.gpu_device(0)
eng = sw.LevenshteinDistances(scope)
dev = torch.device("cuda", 0)
a, b = datasets.acgt_pairs(N, L)
pad = lambda t: torch.cat([torch.from_numpy(np.array(t.data)), torch.zeros(64, dtype=torch.uint8)]).to(dev)
ad, bd = pad(a), pad(b)
ao, bo = torch.from_numpy(np.array(a.offsets).view(np.int64)).to(dev), torch.from_numpy(np.array(b.offsets).view(np.int64)).to(dev)
out = torch.zeros(N, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    eng.compute_pairs_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), N, out.data_ptr(), stream=st)
torch.cuda.synchronize()
eng.set_profiling(True)
for _ in range(5):
    eng.compute_pairs_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), N, out.data_ptr(), stream=st)
ms, n = eng.kernel_time()
print(f"{L}x{L} x {N}: {eng.last_kernel} {ms / n:.4f} ms {N * L * L / (ms / n) / 1e6:.0f} GCUPS checksum {int(out.sum().item())}")
