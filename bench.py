#!/usr/bin/env python
"""bench.py — GCUPS of batched Levenshtein on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ascii64|ascii256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: every rank computes the distances of its own
`--pairs` (default 2^20) synthetic pairs, so per-GPU work is fixed ("scaling": "weak").
  value     whole-job GCUPS with the tapes already resident in HBM; the timed region is the device-resident
            C-ABI call (schedule kernel + compute kernel) — CUDA events, max over ranks.
  e2e       the same metric through the host-pointer C-ABI call a harness makes (`swb_compute_pairs`):
            pinned host tapes in, pinned host results out, H2D/D2H inside the timed region.
  roofline  HBM roofline of the dominant compute kernel: algorithmic bytes (len(a)+len(b) per pair,
            bench.rs:414) / kernel time from CUDA events recorded around that kernel on its launch stream.
  cpu_baseline  the oracle's rapidfuzz-class restatement timed on this box's host cores (bounded sample).
`--impl reference` times that CPU restatement alone (the reference's own crates cannot be built here).
torch is used for device memory, streams and torch.distributed only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (string length, default pairs per GPU, BASELINE.json config it restates)
    "ascii64": (64, 1 << 20, "configs[1]: 1M random ASCII pairs 64x64"),
    "ascii256": (256, 1 << 20, "configs[3] per-GPU shard: 256x256 ASCII pairs (north_star parity gate: 1M length-256 pairs)"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        """Starts the poller and waits for its first row: nvidia-smi's NVML start-up takes driver locks for about a
        second, which would otherwise land inside (and distort) a timed region that is only milliseconds long."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            deadline = time.perf_counter() + 10.0
            while not self.rows and time.perf_counter() < deadline:
                time.sleep(0.05)
            self.mark()
        except Exception:
            self.proc = None

    def mark(self):
        """Forget the rows seen so far: only rows polled after this call count as 'during the timed region'."""
        self.first = len(self.rows)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        rows = self.rows[getattr(self, "first", 0):] or self.rows[-1:]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference(a, b, cells_per_pair: float, budget_s: float, threads: int):
    """Times the oracle's rapidfuzz-class restatement (oracle/lev_oracle.c: orc_rf_lev_bytes) over a bounded
    sample of the workload with `threads` host threads, timed inside C over enough repeats to fill
    `budget_s`. Returns (gcups, pairs_processed, seconds)."""
    from oracle import oracle as orc
    orc.build()
    n = len(a)
    secs, _ = orc.levenshtein_pairs_timed(a.data, a.offsets, b.data, b.offsets, orc.RAPIDFUZZ_CLASS, threads, 1)
    repeats = int(max(1, min(10_000, budget_s / max(secs, 1e-6))))
    secs, _ = orc.levenshtein_pairs_timed(a.data, a.offsets, b.data, b.offsets, orc.RAPIDFUZZ_CLASS, threads, repeats)
    return n * repeats * cells_per_pair / secs / 1e9, n * repeats, secs


def main():
    # The contract is ONE JSON line on stdout. Native libraries (NCCL prints its version banner on stdout) do not
    # know that: point fd 1 at stderr for the whole run and keep the real stdout for the final line only.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ascii64", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU (default: the workload's)")
    ap.add_argument("--algorithm", default="auto", choices=["auto", "wavefront", "bitparallel"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    length, default_pairs, config_name = WORKLOADS[args.workload]
    n_pairs = args.pairs or default_pairs
    cells_per_pair = float(length * length)
    host_threads = os.cpu_count() or 1

    from stringwars_b200 import datasets

    # ------------------------------------------------------------------ reference arm (CPU restatement)
    if args.impl == "reference":
        if rank != 0:
            return 0
        a, b = datasets.ascii_pairs(min(n_pairs, 1 << 18), length, seed=42)
        per_step = max(1.0, min(args.cpu_seconds, 120.0 / max(args.steps + args.warmup, 1)))
        for _ in range(args.warmup):
            cpu_reference(a, b, cells_per_pair, per_step / 4, host_threads)
        cells = secs = 0.0
        sample = 0
        for _ in range(args.steps):
            g, sample, dt = cpu_reference(a, b, cells_per_pair, per_step, host_threads)
            cells += g * 1e9 * dt
            secs += dt
        value = cells / secs / 1e9
        line = {
            "impl": "reference", "metric": "GCUPS (DP cells/s) on batched Levenshtein", "value": value, "unit": "GCUPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 bit-vectors (CPU)", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {config_name}", "pairs_per_step": sample, "length": length},
            "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": host_threads, "kind": "port",
                             "sample": f"{sample} pair evaluations of {length}x{length} per step ({len(a)} distinct pairs repeated), rapidfuzz-class restatement (oracle/lev_oracle.c orc_rf_lev_bytes), {host_threads} threads"},
            "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        emit(line)
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    import stringwars_b200 as sw

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    scope = sw.DeviceScope.gpu_device(local_rank)
    engine = sw.LevenshteinDistances(scope, 0, 1, 1, 1)
    engine.set_algorithm({"auto": sw.ALGO_AUTO, "wavefront": sw.ALGO_WAVEFRONT, "bitparallel": sw.ALGO_BITPARALLEL}[args.algorithm])

    # Two distinct input sets per rank, rotated every step: 2 x (data + offsets) > 126 MB L2 for both workloads.
    host_sets, dev_sets = [], []
    for s in range(2):
        a, b = datasets.ascii_pairs(n_pairs, length, seed=42 + 1000 * rank + s, pinned=True)
        host_sets.append((a, b))
        pad = lambda x: torch.cat([torch.from_numpy(x), torch.zeros(64, dtype=torch.uint8)]).to(device)
        dev_sets.append((pad(a.data), torch.from_numpy(a.offsets.view(np.int64)).to(device),
                         pad(b.data), torch.from_numpy(b.offsets.view(np.int64)).to(device)))
    d_out = torch.zeros(n_pairs, dtype=torch.int64, device=device)
    h_out = sw.PinnedBuffer(8 * n_pairs)
    h_out_arr = h_out.view(np.uint64)
    cells_step = cells_per_pair * n_pairs
    bytes_step_alg = 2.0 * length * n_pairs  # algorithmic bytes: len(a)+len(b) per pair
    stream = torch.cuda.current_stream(device)

    def device_step(i):
        ad, ao, bd, bo = dev_sets[i & 1]
        engine.compute_pairs_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), n_pairs, d_out.data_ptr(),
                                    stream=stream.cuda_stream)

    def host_step(i):
        a, b = host_sets[i & 1]
        engine.compute_pairs(scope, a, b, out=h_out_arr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def timed(step_fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for i in range(steps):
            step_fn(i)
        e1.record(stream)
        torch.cuda.synchronize(device)
        wall = time.perf_counter() - t0
        barrier()
        dev_ms = e0.elapsed_time(e1)
        return dev_ms, wall * 1e3

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident value
    for i in range(args.warmup):
        device_step(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if rank == 0:
        sampler.mark()
    launches0 = sw.launch_count()
    engine.set_profiling(True)
    dev_ms, wall_ms = timed(device_step, args.steps)
    kernel_ms, kernel_launches = engine.kernel_time()
    engine.set_profiling(False)
    launches = sw.launch_count() - launches0
    # The timed region is K x ~0.2 ms — shorter than one nvidia-smi poll. Keep the identical step running (untimed)
    # for ~0.4 s so the clock/throttle record covers the same work under sustained load.
    t_obs = time.perf_counter()
    i_obs = 0
    while time.perf_counter() - t_obs < 0.4:
        device_step(i_obs)
        i_obs += 1
    torch.cuda.synchronize(device)
    kernel_name = engine.last_kernel
    # the device call synchronises once inside (schedule readback), so events and wall agree; report the larger
    step_ms = max_over_ranks(max(dev_ms, wall_ms) / args.steps)
    value = cells_step * world / (step_ms * 1e-3) / 1e9

    # ---- e2e through the host C-ABI
    for i in range(args.warmup):
        host_step(i)
    _, e2e_wall = timed(host_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms = max_over_ranks(e2e_wall / args.steps)
    e2e_value = cells_step * world / (e2e_ms * 1e-3) / 1e9
    checksum = int(h_out_arr.sum())

    # ---- pair-load phase alone (HBM read rate of the staging loads)
    engine.set_profiling(True)
    for i in range(args.warmup + 5):
        ad, ao, bd, bo = dev_sets[i & 1]
        engine.load_phase_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), n_pairs, d_out.data_ptr(), stream=stream.cuda_stream)
        if i == args.warmup - 1:
            engine.kernel_time()
    load_ms, load_n = engine.kernel_time()
    engine.set_profiling(False)

    # ---- N>1: the exchange the north_star names — ONE NCCL scatter of tape slices from rank 0, per-rank compute,
    # ONE gather of results (stringwars_b200/dist.py). Whole batch (world x n_pairs) starts resident on rank 0.
    scatter_gather = None
    if world > 1:
        from stringwars_b200 import dist as swdist
        total_pairs = n_pairs * world
        full_a = full_b = None
        if rank == 0:
            fa, fb = datasets.ascii_pairs(total_pairs, length, seed=4242)
            full_a = (torch.from_numpy(fa.data.copy()).to(device), torch.from_numpy(fa.offsets.view(np.int64).copy()).to(device))
            full_b = (torch.from_numpy(fb.data.copy()).to(device), torch.from_numpy(fb.offsets.view(np.int64).copy()).to(device))

        def sg_compute(ad, ao, bd, bo):
            n_local = ao.numel() - 1
            res = torch.empty(n_local, dtype=torch.int64, device=device)
            if ad.numel() == 0:
                ad = torch.zeros(16, dtype=torch.uint8, device=device)
            if bd.numel() == 0:
                bd = torch.zeros(16, dtype=torch.uint8, device=device)
            engine.compute_pairs_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), n_local, res.data_ptr(), stream=stream.cuda_stream)
            return res

        sg_checksum = None

        def sg_step(i):
            nonlocal sg_checksum
            res = swdist.sharded_pairs(full_a, full_b, device, sg_compute, total_pairs)
            if res is not None and sg_checksum is None:
                sg_checksum = int(res.sum().item())

        for i in range(2):
            sg_step(i)
        sg_steps = max(3, min(args.steps, 10))
        sg_dev, sg_wall = timed(sg_step, sg_steps)
        sg_ms = max_over_ranks(max(sg_dev, sg_wall) / sg_steps)
        scatter_gather = {"value": cells_per_pair * total_pairs / (sg_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": sg_ms,
                          "pairs": total_pairs, "what": "batch resident on rank 0 -> one grouped NCCL send/recv scatter -> per-rank kernels -> one grouped gather",
                          "checksum": sg_checksum}

    peak, peak_src = measured_peaks()
    traffic, traffic_src = None, None
    try:  # DRAM bytes per launch of this kernel on this workload, from the committed ncu --set full capture
        entry = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{kernel_name}|{args.workload}|{n_pairs}")
        if entry:
            traffic, traffic_src = entry["traffic"], entry["source"]
    except Exception:
        pass
    achieved = bytes_step_alg / (kernel_ms / max(kernel_launches, 1) * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    load_gbs = bytes_step_alg / (load_ms / max(load_n, 1) * 1e-3) / 1e9 if load_ms > 0 else 0.0

    line = None
    if rank == 0:
        a0, b0 = host_sets[0]
        cpu_gcups, cpu_sample, cpu_dt = (cpu_reference(a0, b0, cells_per_pair, args.cpu_seconds, host_threads) if world == 1 else (None, 0, 0))
        line = {
            "metric": "GCUPS (DP cells/s) on batched Levenshtein", "value": value, "unit": "GCUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 bit-vectors" if "bitpar" in kernel_name else "int32",
            "data": "synthetic (seed 42, iid printable ASCII)",
            "config": {"workload": f"{args.workload}: {config_name}", "pairs_per_gpu": n_pairs, "length": length,
                       "algorithm": args.algorithm, "kernel": kernel_name,
                       "l2": "two input sets (2 x %.0f MB) rotated every step: larger than the 126 MB L2" % ((bytes_step_alg + 16.0 * n_pairs) / 1e6),
                       "timing": "max(CUDA events on the launch stream, host wall) / steps, max over ranks"},
            "e2e": {"value": e2e_value, "unit": "GCUPS", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(bytes_step_alg + 2 * 8 * (n_pairs + 1)), "d2h_bytes_per_step": 8 * n_pairs,
                    "api": "swb_compute_pairs (host tapes, pinned)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": kernel_name, "kernel_ms": kernel_ms / max(kernel_launches, 1), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_step_alg,
                         "kernel_gcups": cells_step / (kernel_ms / max(kernel_launches, 1) * 1e-3) / 1e9 if kernel_ms else None,
                         "load_phase": {"kernel": "load_phase_kernel", "achieved": load_gbs, "frac": load_gbs / peak, "ms": load_ms / max(load_n, 1)}},
            "clocks": dict(clocks or {}, note="nvidia-smi polled every 100 ms from the start of the timed region through an untimed 0.4 s continuation of the same step and the e2e region"),
            "checksum": checksum,
        }
        if scatter_gather is not None:
            line["scatter_gather"] = scatter_gather
        if cpu_gcups is not None:
            line["cpu_baseline"] = {"value": cpu_gcups, "unit": "GCUPS", "cores": host_threads, "kind": "port",
                                    "sample": f"{cpu_sample} pair evaluations of {length}x{length} in {cpu_dt:.1f} s ({len(a0)} distinct pairs repeated), rapidfuzz-class restatement (oracle/lev_oracle.c orc_rf_lev_bytes), {host_threads} threads"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
