"""bench.py honours the driver's contract: exactly one JSON line on stdout with the required keys."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(args, timeout):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"stdout must hold exactly one line, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_line():
    """`--impl reference`: the CPU restatement alone, all host threads, same metric/unit/config keys, no GPU needed."""
    line = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu-seconds", "0.3", "--pairs", "4096"], 120)
    assert BASE_KEYS <= set(line)
    assert line["impl"] == "reference" and line["unit"] == "GCUPS" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]
    # both arms describe the workload with the same `config` object
    sys.path.insert(0, ROOT)
    import bench
    length, _, name = bench.WORKLOADS["ascii64"]
    assert line["config"] == bench.shared_config("ascii64", name, 4096, length)


@pytest.mark.gpu
def test_our_arm_line():
    line = _run(["--steps", "3", "--warmup", "3", "--cpu-seconds", "0.5", "--pairs", "131072", "--no-gates"], 600)
    assert BASE_KEYS | {"roofline", "clocks"} <= set(line)
    assert line["n_gpus"] == 1 and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["gpu_launches"] >= line["steps"] and line["value"] > line["cpu_baseline"]["value"]
    roof = line["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    assert roof["kernel"].startswith("lev_bitpar") and roof["load_phase"]["frac"] > 0
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] == 2 * 64 * 131072 + 16 * (131072 + 1) and e2e["d2h_bytes_per_step"] == 8 * 131072
    assert 0 < e2e["value"] < line["value"]
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
