"""Host-tape (e2e) call timing with the native pipeline trace, for a few chunk sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stringwars_b200 as sw
from stringwars_b200 import datasets
L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = 1 << 20
scope = sw.DeviceScope.gpu_device(0)
eng = sw.LevenshteinDistances(scope)
sets = [datasets.ascii_pairs(n, L, seed=s, pinned=True) for s in (1, 2)]
out = sw.PinnedBuffer(8 * n)
o = out.view(np.uint64)
for i in range(3):
    eng.compute_pairs(scope, *sets[i & 1], out=o)
t0 = time.perf_counter()
for i in range(10):
    eng.compute_pairs(scope, *sets[i & 1], out=o)
dt = (time.perf_counter() - t0) / 10
nbytes = 2 * L * n + 16 * (n + 1)
print(f"L={L} chunk_mb={os.environ.get('SWB_CHUNK_MB', '8')}: {dt * 1e3:.3f} ms/call, {n * L * L / dt / 1e9:.0f} GCUPS, H2D {nbytes / dt / 1e9:.1f} GB/s", flush=True)
