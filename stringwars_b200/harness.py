"""Python twin of the reference's measurement loop, so our contender prints the same line as theirs.

Restates `measure_throughput` / `WorkUnits` / `ReportAs::Cups` / `BenchStats::report` (StringWars
utils.rs:529-561, 652-692, 721-799; Python twin utils.py:291-336): uncounted warm-up, wall-time budget, one
closure call per iteration returning its work units, CUPS = sum(cells) / elapsed.
"""
from __future__ import annotations

import os
import re
import time
from dataclasses import dataclass
from typing import Callable, Optional


@dataclass
class WorkUnits:
    """`WorkUnits::new(elements, bytes)` (utils.rs:529-545): cells and bytes one closure call processed."""
    elements: int
    bytes: int


@dataclass
class BenchBudget:
    """`BenchBudget::from_env(default_warmup, default_time)` (utils.rs:566-581)."""
    warmup_seconds: float = 5.0
    time_seconds: float = 30.0

    @classmethod
    def from_env(cls, default_warmup: float = 5.0, default_time: float = 30.0) -> "BenchBudget":
        def _f(name, default):
            try:
                return float(os.environ.get(name, default))
            except ValueError:
                return default
        return cls(_f("STRINGWARS_WARMUP", default_warmup), _f("STRINGWARS_TIME", default_time))


def should_run(name: str) -> bool:
    """`should_run` (utils.rs:457-483): `STRINGWARS_FILTER` is a regex matched against the variant name."""
    pattern = os.environ.get("STRINGWARS_FILTER")
    if not pattern:
        return True
    try:
        return re.search(pattern, name) is not None
    except re.error:
        return True


def scale_si(value: float) -> str:
    """SI scaling G/M/k used by the report line (utils.rs:487-500, 705-714)."""
    for factor, suffix in ((1e12, "T"), (1e9, "G"), (1e6, "M"), (1e3, "k")):
        if value >= factor:
            return f"{value / factor:.2f} {suffix}"
    return f"{value:.2f} "


@dataclass
class BenchStats:
    name: str
    elapsed_seconds: float
    elements: int
    bytes: int
    calls: int
    latencies_ns: list

    @property
    def cups(self) -> float:
        return self.elements / self.elapsed_seconds if self.elapsed_seconds > 0 else 0.0

    def report(self) -> str:
        """Canonical line: name | CUPS | bytes/s | p50 | p99 (perf-counter columns of utils.rs:672-680 are
        CPU-only and omitted)."""
        lat = sorted(self.latencies_ns)
        q = lambda f: lat[min(len(lat) - 1, int(f * len(lat)))] / 1e3 if lat else 0.0
        return (f"{self.name:<46} {scale_si(self.cups)}CUPS | {scale_si(self.bytes / max(self.elapsed_seconds, 1e-12))}B/s"
                f" | p50 {q(0.5):.1f} us | p99 {q(0.99):.1f} us | {self.calls} calls")


def measure_throughput(name: str, budget: BenchBudget, closure: Callable[[], WorkUnits],
                       printer: Optional[Callable[[str], None]] = print) -> Optional[BenchStats]:
    """`measure_throughput(name, ReportAs::Cups, budget, closure)` (utils.rs:721-799)."""
    if not should_run(name):
        return None
    deadline = time.perf_counter() + budget.warmup_seconds
    while budget.warmup_seconds > 0 and time.perf_counter() < deadline:
        closure()
    elements = nbytes = calls = 0
    latencies = []
    start = time.perf_counter()
    deadline = start + budget.time_seconds
    while True:
        t0 = time.perf_counter_ns()
        units = closure()
        latencies.append(time.perf_counter_ns() - t0)
        elements += units.elements
        nbytes += units.bytes
        calls += 1
        if time.perf_counter() >= deadline:  # at least one measured call, like the reference
            break
    stats = BenchStats(name, time.perf_counter() - start, elements, nbytes, calls, latencies)
    if printer:
        printer(stats.report())
    return stats
