"""Load-phase kernel timing (CUDA events around the kernel) for the current SWB_LOAD_VARIANT / SWB_LOAD_TILE."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import stringwars_b200 as sw
from stringwars_b200 import datasets
L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = 1 << 20
scope = sw.DeviceScope.gpu_device(0); eng = sw.LevenshteinDistances(scope)
dev = torch.device("cuda", 0)
sets = []
for s in (1, 2):
    a, b = datasets.ascii_pairs(n, L, seed=s)
    pad = lambda t: torch.cat([torch.from_numpy(t.data.copy()), torch.zeros(64, dtype=torch.uint8)]).to(dev)
    sets.append((pad(a), torch.from_numpy(a.offsets.view(np.int64).copy()).to(dev), pad(b), torch.from_numpy(b.offsets.view(np.int64).copy()).to(dev)))
out = torch.zeros(n, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
eng.set_profiling(True)
for i in range(13):
    ad, ao, bd, bo = sets[i & 1]
    eng.load_phase_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), n, out.data_ptr(), stream=st)
    if i == 2: eng.kernel_time()
ms, k = eng.kernel_time()
ms /= k
print(f"L={L} variant={os.environ.get('SWB_LOAD_VARIANT','0')} tile={os.environ.get('SWB_LOAD_TILE','256')}: {ms*1e3:.1f} us, algorithmic {2*L*n/ms/1e6:.0f} GB/s ({2*L*n/ms/1e6/6574.5*100:.1f}% of measured peak), total {(2*L*n+24*n)/ms/1e6:.0f} GB/s", flush=True)
