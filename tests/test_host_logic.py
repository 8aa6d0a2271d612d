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


def test_alphabetic_words_fit_one_byte_alphabet():
    """The code-point cross-product workload: valid UTF-8 words whose multi-byte code points (Latin-1, Cyrillic) number at
    most 127 over the whole tape — what the engine's batch-wide byte alphabet needs (utf8.cuh BATCH_DICT_CAP)."""
    words = datasets.alphabetic_words(3000, seed=2)
    text = [words[i].decode("utf-8") for i in range(len(words))]
    assert all(1 <= len(w) <= 12 for w in text)
    wide = {ch for w in text for ch in w if ord(ch) >= 0x80}
    assert 40 < len(wide) <= 127
    assert any(ord(ch) < 0x80 for w in text for ch in w)
    again = datasets.alphabetic_words(3000, seed=2)
    assert np.array_equal(words.data, again.data) and np.array_equal(words.offsets, again.offsets)


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


def test_tapes_with_u32_offsets_and_views():
    """`BytesTapeView<u32>` / `AnyBytesTape::View32`: offsets keep their width through subviews and into the C view."""
    import pytest
    t = BytesTape.from_strings([b"kitten", b"", b"sitting", "naïve"], offset_dtype=np.uint32)
    assert t.offsets.dtype == np.uint32 and t.offset_bits == 32 and len(t) == 4
    assert t[2] == b"sitting" and t[3] == "naïve".encode() and t.lengths().tolist() == [6, 0, 7, 6]
    sub = t.subview(1, 3)
    assert sub.offsets.dtype == np.uint32 and len(sub) == 2 and sub[1] == b"sitting" and int(sub.offsets[0]) == 6  # offsets stay absolute
    v = sub.view()
    assert v.n == 2 and v.offset_bits == 32 and v.data == t.data.ctypes.data and v.offsets == sub.offsets.ctypes.data
    wide = t.with_offsets(np.uint64)
    assert wide.offset_bits == 64 and wide[3] == t[3] and wide.view().offset_bits == 64
    assert BytesTape(np.zeros(4, np.uint8), [0, 2, 4]).offsets.dtype == np.uint64  # anything else is widened to u64
    with pytest.raises(IndexError):
        t.subview(2, 9)
    empty = BytesTape.from_strings([])
    assert len(empty) == 0 and empty.view().n == 0 and empty.view().data != 0  # never a NULL data pointer


def test_result_width_checks_need_no_gpu():
    """`swb_result_type_t`: the Python mirror accepts 64 / 32 / 16-bit results of the engine's signedness only."""
    import pytest
    from stringwars_b200 import _lib, szs
    assert szs._check_out(np.zeros(5, np.uint16), np.uint64, 5) == _lib.RESULT_16
    assert szs._check_out(np.zeros(5, np.uint32), np.uint64, 5) == _lib.RESULT_32
    assert szs._check_out(np.zeros(5, np.int64), np.int64, 5) == _lib.RESULT_64
    for bad in (np.zeros(5, np.int16), np.zeros(4, np.uint16), np.zeros(5, np.uint8), np.zeros((5, 2), np.uint64)[:, 0]):
        with pytest.raises(ValueError):
            szs._check_out(bad, np.uint64, 5)


def test_bench_config_object_is_shared_by_both_arms():
    import bench
    length, pairs, name = bench.WORKLOADS["ascii64"]
    cfg = bench.shared_config("ascii64", name, pairs, length)
    assert cfg["pairs_per_gpu"] == 1 << 20 and cfg["length"] == 64 and "configs[1]" in cfg["workload"] and "L2" in cfg["l2"]
