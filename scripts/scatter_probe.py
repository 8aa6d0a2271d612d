"""Single-process multi-GPU probe: config 4's shard size per GPU, batch resident on GPU 0, NCCL scatter vs peer loads.
    python scripts/scatter_probe.py [n_gpus] [pairs_per_gpu]      (NCCL_* environment is read by NCCL at communicator creation)"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stringwars_b200 as sw
from stringwars_b200 import _lib

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
per_gpu = int(sys.argv[2]) if len(sys.argv) > 2 else 1_250_000
L, total = 256, per_gpu * n_gpus
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(4)
ad = torch.randint(0x20, 0x7F, (total * L + 64,), dtype=torch.uint8, device=dev, generator=gen)
bd = torch.randint(0x20, 0x7F, (total * L + 64,), dtype=torch.uint8, device=dev, generator=gen)
offs = torch.arange(total + 1, dtype=torch.int64, device=dev) * L
out = torch.zeros(total, dtype=torch.int16, device=dev)
ref = torch.zeros(total, dtype=torch.int16, device=dev)
torch.cuda.synchronize()
scope = sw.DeviceScope.gpu_devices(0, n_gpus)
lev = sw.LevenshteinDistances(scope)
va, vb = _lib.TapeView(ad.data_ptr(), offs.data_ptr(), total, 64), _lib.TapeView(bd.data_ptr(), offs.data_ptr(), total, 64)
single = sw.DeviceScope.gpu_device(0)
sl = sw.LevenshteinDistances(single)
sl.compute_pairs_device_view(va, vb, ref.data_ptr(), _lib.RESULT_16)
torch.cuda.synchronize()
cells = float(L) * L * total
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("NCCL_")) or "default NCCL settings"
ce_scope = sw.DeviceScope.gpu_devices(0, n_gpus).set_transport(sw.TRANSPORT_COPY_ENGINE)
ce_lev = sw.LevenshteinDistances(ce_scope)
for name, fn in (("nccl scatter", lev.compute_pairs_scatter), ("copy-engine scatter", ce_lev.compute_pairs_scatter), ("peer loads", lev.compute_pairs_peer)):
    for _ in range(2):
        st = fn(va, vb, out.data_ptr(), _lib.RESULT_16)
    t0 = time.perf_counter()
    for _ in range(5):
        st = fn(va, vb, out.data_ptr(), _lib.RESULT_16)
    ms = (time.perf_counter() - t0) / 5 * 1e3
    moved = st.get("sent_bytes", st.get("remote_read_bytes", 0))
    ok = bool(torch.equal(out, ref))
    print(f"[{tag}] {n_gpus} GPUs, {total} pairs: {name}: {ms:.3f} ms, {cells / ms / 1e6:.0f} GCUPS, {moved / ms / 1e6:.0f} GB/s over NVLink, parity {ok}", flush=True)
