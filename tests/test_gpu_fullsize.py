"""BASELINE.json's full sizes on the GPU. The oracle cannot finish 1M x 256x256 in seconds on a test budget, so
full-size parity is established by (i) size-independent properties and (ii) bit-exact oracle checks on large
random samples of the full batch."""
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


def _sample_check(oracle, a, b, got, k=40_000, seed=0):
    idx = np.sort(np.random.default_rng(seed).choice(len(a), size=min(k, len(a)), replace=False))
    sa = BytesTape.from_strings([a[int(i)] for i in idx[:4000]])
    sb = BytesTape.from_strings([b[int(i)] for i in idx[:4000]])
    want = oracle.levenshtein_pairs(sa.data, sa.offsets, sb.data, sb.offsets, oracle.RAPIDFUZZ_CLASS, threads=16)
    assert np.array_equal(got[idx[:4000]], want)


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
    # oracle on a random sample + on the first 20k pairs
    _sample_check(oracle, a, b, d_ab)
    head = oracle.levenshtein_pairs(a.data, a.offsets[:20_001], b.data, b.offsets[:20_001], oracle.RAPIDFUZZ_CLASS, threads=16)
    assert np.array_equal(d_ab[:20_000], head)
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
    _sample_check(oracle, a, b, d, seed=1)


def test_config5_full_nw_10k_acgt(ctx, oracle):
    scope, _ = ctx
    btc, costs = sw.unary_class_costs(2, -1)
    a, b = datasets.acgt_pairs(10_000, 1024)
    for open_, extend in ((-2, -2), (-5, -1)):
        nw = sw.NeedlemanWunschScores(scope, btc, costs, open_, extend)
        got = nw.compute_pairs(scope, a, b)
        assert np.array_equal(got, nw.compute_pairs(scope, b, a))            # symmetric costs => symmetric score
        assert (got <= 2 * 1024).all() and (got >= -3 * 2048).all()
        want = oracle.align_pairs(a.data, a.offsets[:601], b.data, b.offsets[:601], btc, costs, open_, extend, threads=16)
        assert np.array_equal(got[:600], want)
