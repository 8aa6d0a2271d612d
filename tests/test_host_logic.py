"""Host-side logic that needs no GPU: tapes, synthetic config generators, harness accounting."""
import numpy as np

from stringwars_b200 import datasets, harness
from stringwars_b200.tape import BytesTape


def test_generators_are_deterministic_and_shaped():
    a1, b1 = datasets.ascii_pairs(1000, 64)
    a2, b2 = datasets.ascii_pairs(1000, 64)
    assert np.array_equal(a1.data, a2.data) and np.array_equal(b1.data, b2.data)
    assert a1.data.min() >= 0x20 and a1.data.max() <= 0x7E and (a1.lengths() == 64).all()
    assert not np.array_equal(a1.data, b1.data)
    wa, wb = datasets.words_pairs(1000)
    assert wa.lengths().min() >= 1 and wa.lengths().max() <= 16 and len(wa) == len(wb) == 1000
    ga, _ = datasets.acgt_pairs(10, 1024)
    assert set(np.unique(ga.data)) <= set(b"ACGT") and (ga.lengths() == 1024).all()


def test_utf8_lines_are_valid_utf8_in_range():
    a, b = datasets.utf8_line_pairs(300)
    for t in (a, b):
        lens = t.lengths()
        assert lens.min() >= 120 and lens.max() <= 512
        for i in range(len(t)):
            t[i].decode("utf-8")
    multibyte = sum(len(a[i]) != len(a[i].decode()) for i in range(len(a)))
    assert multibyte > 100  # mixed scripts really are there


def test_cells_accounting_matches_harness_definition():
    a = BytesTape.from_strings([b"abc", b"", b"xy"])
    b = BytesTape.from_strings([b"abcd", b"q", b""])
    cells, nbytes = datasets.cells_and_bytes(a, b)
    assert cells == 3 * 4 + 0 * 1 + 2 * 0 and nbytes == 3 + 4 + 0 + 1 + 2 + 0


def test_measure_throughput_loop(monkeypatch):
    calls = []
    stats = harness.measure_throughput("uniform/test<1gpu>", harness.BenchBudget(0.0, 0.05),
                                       lambda: calls.append(1) or harness.WorkUnits(1000, 10), printer=None)
    assert stats.calls == len(calls) >= 1 and stats.elements == 1000 * stats.calls
    assert "CUPS" in stats.report() and "uniform/test<1gpu>" in stats.report()
    monkeypatch.setenv("STRINGWARS_FILTER", "rapidfuzz")
    assert harness.measure_throughput("uniform/test<1gpu>", harness.BenchBudget(0, 0.01), lambda: harness.WorkUnits(1, 1), None) is None
    assert harness.scale_si(2.5e12).startswith("2.50 T") and harness.scale_si(3.2e9).startswith("3.20 G")


def test_budget_from_env(monkeypatch):
    monkeypatch.setenv("STRINGWARS_TIME", "1.5")
    monkeypatch.setenv("STRINGWARS_WARMUP", "0")
    b = harness.BenchBudget.from_env(5.0, 30.0)
    assert b.time_seconds == 1.5 and b.warmup_seconds == 0.0
