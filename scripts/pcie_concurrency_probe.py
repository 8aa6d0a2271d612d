"""Host<->device copy bandwidth with 1, 2, 4, 8 GPUs copying AT THE SAME TIME (one process, one stream per GPU,
page-locked buffers): the platform ceiling behind the e2e scaling numbers. Prints one JSON object.

    python scripts/pcie_concurrency_probe.py > gpurun_out/pcie_concurrency.json
"""
import glob
import json
import os
import subprocess
import sys
import time

import torch

MB = 256


def topo():
    info = {"cpus": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    info["numa_nodes"] = {p.split("node")[-1].split("/")[0]: open(p).read().strip() for p in sorted(glob.glob("/sys/devices/system/node/node*/cpulist"))}
    gpus = []
    for i in range(torch.cuda.device_count()):
        bus = torch.cuda.get_device_properties(i).pci_bus_id if hasattr(torch.cuda.get_device_properties(i), "pci_bus_id") else None
        entry = {"index": i, "pci_bus_id": bus}
        try:
            q = subprocess.run(["nvidia-smi", "-i", str(i), "--query-gpu=pci.bus_id,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max", "--format=csv,noheader"],
                               capture_output=True, text=True, timeout=20).stdout.strip()
            entry["nvidia_smi"] = q
            bdf = q.split(",")[0].strip().lower()
            bdf = bdf[4:] if len(bdf) > 12 else bdf  # 00000000:1B:00.0 -> 0000:1b:00.0
            node = f"/sys/bus/pci/devices/{bdf}/numa_node"
            entry["numa_node"] = open(node).read().strip() if os.path.exists(node) else None
        except Exception as exc:
            entry["error"] = str(exc)
        gpus.append(entry)
    info["gpus"] = gpus
    try:
        info["lspci_tree_head"] = subprocess.run(["lspci", "-tv"], capture_output=True, text=True, timeout=20).stdout.splitlines()[:40]
    except Exception:
        info["lspci_tree_head"] = None
    return info


def measure(n_gpus, direction):
    n = MB << 20
    host, dev, streams = [], [], []
    for g in range(n_gpus):
        torch.cuda.set_device(g)
        host.append(torch.empty(n, dtype=torch.uint8).pin_memory())
        host[-1].fill_(g + 1)
        dev.append(torch.empty(n, dtype=torch.uint8, device=f"cuda:{g}"))
        streams.append(torch.cuda.Stream(device=g))

    def round_():
        for g in range(n_gpus):
            with torch.cuda.stream(streams[g]):
                if direction == "h2d":
                    dev[g].copy_(host[g], non_blocking=True)
                else:
                    host[g].copy_(dev[g], non_blocking=True)

    for _ in range(2):
        round_()
    for g in range(n_gpus):
        torch.cuda.synchronize(g)
    reps = 8
    t0 = time.perf_counter()
    for _ in range(reps):
        round_()
    for g in range(n_gpus):
        torch.cuda.synchronize(g)
    dt = time.perf_counter() - t0
    total = n * reps * n_gpus / dt / 1e9
    return {"gpus": n_gpus, "direction": direction, "aggregate_gbs": total, "per_gpu_gbs": total / n_gpus}


def main():
    out = {"topology": topo(), "buffer_mb": MB, "runs": []}
    have = torch.cuda.device_count()
    for n in (1, 2, 4, 8):
        if n <= have:
            for direction in ("h2d", "d2h"):
                out["runs"].append(measure(n, direction))
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
