"""H2D bandwidth from pinned memory allocated under different CPU affinities (NUMA placement check)."""
import os, sys, time, glob
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def numa_nodes():
    nodes = {}
    for p in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
        nodes[int(p.split("node")[-1].split("/")[0])] = open(p).read().strip()
    return nodes

def parse(cpulist):
    out = []
    for part in cpulist.split(","):
        if "-" in part:
            a, b = part.split("-"); out += list(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out

print("numa nodes:", numa_nodes())
try:
    bus = torch.cuda.get_device_properties(0).pci_bus_id
except Exception:
    bus = None
for p in glob.glob("/sys/bus/pci/devices/*/numa_node"):
    pass
os.system("nvidia-smi topo -m 2>/dev/null | head -12")
dev = torch.device("cuda", 0)
n = 128 << 20
d = torch.empty(n, dtype=torch.uint8, device=dev)
all_cpus = sorted(os.sched_getaffinity(0))
for node, cl in numa_nodes().items():
    cpus = [c for c in parse(cl) if c in all_cpus]
    if not cpus:
        continue
    os.sched_setaffinity(0, cpus)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    e0.record()
    for _ in range(10):
        h2.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 10
    print(f"node {node}: H2D {n / ms / 1e6:.1f} GB/s, D2H {n / ms2 / 1e6:.1f} GB/s")
    del h, h2
os.sched_setaffinity(0, all_cpus)
