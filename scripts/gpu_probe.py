"""Quick on-GPU exploration: parity spot checks + kernel-only timings per kernel family. Not a bench."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import stringwars_b200 as sw
from stringwars_b200 import datasets
from oracle import oracle as orc

scope = sw.DeviceScope.gpu_device(0)
eng = sw.LevenshteinDistances(scope)
dev = torch.device("cuda", 0)
res = {}

def to_dev(t):
    pad = torch.cat([torch.from_numpy(t.data), torch.zeros(64, dtype=torch.uint8)]).to(dev)
    return pad, torch.from_numpy(t.offsets.view(np.int64)).to(dev)

def time_kernel(a, b, algo, reps=5):
    ad, ao = to_dev(a); bd, bo = to_dev(b)
    out = torch.zeros(len(a), dtype=torch.int64, device=dev)
    eng.set_algorithm(algo)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        eng.compute_pairs_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), len(a), out.data_ptr(), stream=st)
    torch.cuda.synchronize()
    eng.set_profiling(True)
    for _ in range(reps):
        eng.compute_pairs_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), len(a), out.data_ptr(), stream=st)
    ms, n = eng.kernel_time()
    eng.set_profiling(False)
    return ms / reps, eng.last_kernel, out.cpu().numpy().view(np.uint64)

for name, gen, npairs in (("ascii64", lambda n: datasets.ascii_pairs(n, 64), 1 << 20), ("ascii256", lambda n: datasets.ascii_pairs(n, 256), 1 << 18),
                          ("words", lambda n: datasets.words_pairs(n), 1 << 20), ("utf8", lambda n: datasets.utf8_line_pairs(n), 50000),
                          ("acgt1024", lambda n: datasets.acgt_pairs(n, 1024), 10000)):
    a, b = gen(npairs)
    cells, nbytes = datasets.cells_and_bytes(a, b)
    sample = min(len(a), 20000)
    want = orc.levenshtein_pairs(a.data, a.offsets[:sample + 1], b.data, b.offsets[:sample + 1], orc.RAPIDFUZZ_CLASS, threads=16)
    for algo, an in ((sw.ALGO_AUTO, "auto"), (sw.ALGO_WAVEFRONT, "wavefront")):
        try:
            ms, kern, got = time_kernel(a, b, algo)
            ok = bool(np.array_equal(got[:sample], want))
            res[f"{name}/{an}"] = {"kernel": kern, "ms": ms, "gcups": cells / ms / 1e6, "gbs": nbytes / ms / 1e6, "parity": ok}
        except Exception as e:
            res[f"{name}/{an}"] = {"error": str(e)}
        print(name, an, res[f"{name}/{an}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe.json", "w"), indent=1)
