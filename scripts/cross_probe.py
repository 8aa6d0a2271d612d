"""Cross-product timing in the harness's shape: side x side over disjoint halves of one tape (bench.rs:123-148)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stringwars_b200 as sw
from stringwars_b200 import datasets
scope = sw.DeviceScope.gpu_device(0)
lev = sw.LevenshteinDistances(scope)
for name, tape, side in (("words side=195 (harness default batch on B200)", datasets.words_pairs(4000)[0], 195), ("words side=1557", datasets.words_pairs(4000)[0], 1557),
                         ("ascii64 side=1557", datasets.ascii_pairs(4000, 64)[0], 1557), ("lines~320B side=195", datasets.utf8_line_pairs(500)[0], 195),
                         ("acgt1024 side=195", datasets.acgt_pairs(400, 1024)[0], 195)):
    q, c = tape.subview(0, side), tape.subview(side, 2 * side)
    cells = int(q.lengths().sum()) * int(c.lengths().sum())
    m = sw.unified_mat(side, side, np.uint64)  # the harness's UnifiedMat: pinned here
    for _ in range(3): lev.compute_into(scope, q, c, m)
    t0 = time.perf_counter(); reps = 10
    for _ in range(reps): lev.compute_into(scope, q, c, m)
    dt = (time.perf_counter() - t0) / reps
    lev.set_profiling(True); lev.compute_into(scope, q, c, m); kms, kn = lev.kernel_time(); lev.set_profiling(False)
    print(f"{name}: {dt*1e3:.3f} ms/call (host tapes, e2e) {cells/dt/1e9:.0f} GCUPS  kernel={lev.last_kernel} kernel-only {kms:.3f} ms ({kn} launches) checksum={int(m.sum())}", flush=True)
