"""BASELINE.json's full sizes on the GPU: every pair of every full-size batch is checked bit-exact against the oracle
(the rapidfuzz-class family finishes 1M x 256x256 in about a second on the box's cores), next to size-independent
properties (symmetry, identity, both kernel families agreeing, checksum of chunked calls)."""
import numpy as np
import pytest

import stringwars_b200 as sw
from stringwars_b200 import datasets
from stringwars_b200.tape import BytesTape

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    scope = sw.DeviceScope.gpu_device(0)
    yield scope, sw.LevenshteinDistances(scope)
    scope.close()


def _full_check(oracle, a, b, got):
    want = oracle.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, oracle.RAPIDFUZZ_CLASS, threads=32)
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, f"{bad.size} of {len(a)} pairs differ from the oracle, first {bad[:5]}"


@pytest.mark.parametrize("length", [64, 256])
def test_one_million_pairs(ctx, oracle, length):
    """configs[1] (1M x 64x64) and the north-star parity gate (1M length-256 pairs)."""
    scope, lev = ctx
    n = 1 << 20
    a, b = datasets.ascii_pairs(n, length)
    d_ab = lev.compute_pairs(scope, a, b)
    # properties that hold for every pair
    assert d_ab.max() <= length and d_ab.min() >= 0
    d_ba = lev.compute_pairs(scope, b, a)
    assert np.array_equal(d_ab, d_ba)                                     # symmetry
    assert (lev.compute_pairs(scope, a, a) == 0).all()                    # identity
    # both kernel families agree on the whole batch
    wf = lev.set_algorithm(sw.ALGO_WAVEFRONT).compute_pairs(scope, a, b)
    lev.set_algorithm(sw.ALGO_AUTO)
    assert np.array_equal(d_ab, wf)
    # the oracle on ALL 2^20 pairs
    _full_check(oracle, a, b, d_ab)
    # checksum-of-checksums: chunked calls give the same total as one call
    total = int(d_ab.sum())
    parts = sum(int(lev.compute_pairs(scope, a.subview(lo, lo + n // 8), b.subview(lo, lo + n // 8)).sum()) for lo in range(0, n, n // 8))
    assert total == parts


def test_config3_full_100k_utf8_lines(ctx, oracle):
    scope, lev = ctx
    a, b = datasets.utf8_line_pairs(100_000)
    d = lev.compute_pairs(scope, a, b)
    la, lb = a.lengths().astype(np.int64), b.lengths().astype(np.int64)
    assert (d.astype(np.int64) >= np.abs(la - lb)).all() and (d.astype(np.int64) <= np.maximum(la, lb)).all()
    assert np.array_equal(d, lev.compute_pairs(scope, b, a))
    _full_check(oracle, a, b, d)


def test_config5_full_nw_10k_acgt(ctx, oracle):
    scope, _ = ctx
    btc, costs = sw.unary_class_costs(2, -1)
    a, b = datasets.acgt_pairs(10_000, 1024)
    for open_, extend in ((-2, -2), (-5, -1)):
        nw = sw.NeedlemanWunschScores(scope, btc, costs, open_, extend)
        got = nw.compute_pairs(scope, a, b)
        assert np.array_equal(got, nw.compute_pairs(scope, b, a))            # symmetric costs => symmetric score
        assert (got <= 2 * 1024).all() and (got >= -3 * 2048).all()
        want = oracle.align_pairs(a.data, a.offsets, b.data, b.offsets, btc, costs, open_, extend, threads=32)  # all 10k pairs
        assert np.array_equal(got, want)
