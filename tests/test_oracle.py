"""Pins the CPU oracle (oracle/lev_oracle.c). The reference holds no golden vectors for this path (parity is
"unpinned by the reference"), so the pins are: literature known answers, an independent
pure-Python implementation, committed vectors from a third-party implementation (torchaudio), agreement of the
two algorithm families, and algebraic properties."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import pyref
from stringwars_b200 import datasets
from stringwars_b200.tape import BytesTape

KNOWN = [(b"kitten", b"sitting", 3), (b"flaw", b"lawn", 2), (b"", b"abc", 3), (b"abc", b"", 3), (b"", b"", 0),
         (b"saturday", b"sunday", 3), (b"intention", b"execution", 5), (b"abc", b"abc", 0), (b"a" * 70, b"a" * 70, 0),
         (b"a" * 70, b"b" * 70, 70), (b"a" * 130, b"", 130)]


@pytest.mark.parametrize("a,b,d", KNOWN)
def test_known_answers(oracle, a, b, d):
    assert oracle.levenshtein(a, b, oracle.WAGNER_FISCHER) == d
    assert oracle.levenshtein(a, b, oracle.RAPIDFUZZ_CLASS) == d
    assert oracle.levenshtein(a, b, oracle.UTF8) == d
    assert pyref.levenshtein(a, b) == d


def test_golden_vectors(oracle, golden):
    for case in golden:
        a, b = bytes.fromhex(case["a"]), bytes.fromhex(case["b"])
        assert oracle.levenshtein(a, b, oracle.WAGNER_FISCHER) == case["bytes"], case["src"]
        assert oracle.levenshtein(a, b, oracle.RAPIDFUZZ_CLASS) == case["bytes"], case["src"]
        if "chars" in case:
            assert oracle.levenshtein(a, b, oracle.UTF8) == case["chars"], case["src"]


def test_families_agree_on_every_config(oracle):
    gens = [datasets.words_pairs(1000), datasets.ascii_pairs(2000, 64), datasets.ascii_pairs(300, 256),
            datasets.utf8_line_pairs(200), datasets.acgt_pairs(20, 1024), datasets.mutated_pairs(200, 64)]
    for a, b in gens:
        wf = oracle.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, oracle.WAGNER_FISCHER, threads=4)
        rf = oracle.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, oracle.RAPIDFUZZ_CLASS, threads=4)
        assert np.array_equal(wf, rf)


def test_against_pure_python(oracle):
    rng = np.random.default_rng(3)
    for _ in range(150):
        la, lb = int(rng.integers(0, 40)), int(rng.integers(0, 40))
        a = bytes(rng.integers(97, 100, size=la, dtype=np.uint8))
        b = bytes(rng.integers(97, 100, size=lb, dtype=np.uint8))
        assert oracle.levenshtein(a, b) == pyref.levenshtein(a, b)
        assert oracle.levenshtein(a, b, oracle.RAPIDFUZZ_CLASS) == pyref.levenshtein(a, b)


def test_block_boundaries(oracle):
    """Pattern lengths around the 64-symbol word boundary of the bit-parallel family."""
    rng = np.random.default_rng(5)
    for m in (1, 31, 32, 33, 63, 64, 65, 127, 128, 129, 191, 192, 193, 300):
        for n in (m, m + 1, 2 * m + 3, max(1, m // 2)):
            a = bytes(rng.integers(97, 101, size=m, dtype=np.uint8))
            b = bytes(rng.integers(97, 101, size=n, dtype=np.uint8))
            assert oracle.levenshtein(a, b, oracle.RAPIDFUZZ_CLASS) == oracle.levenshtein(a, b), (m, n)


def test_utf8_matches_python_decoder(oracle):
    a, b = datasets.utf8_line_pairs(40, 30, 90, seed=5)
    for i in range(len(a)):
        a[i].decode("utf-8"), b[i].decode("utf-8")  # generator emits valid UTF-8
        assert oracle.levenshtein(a[i], b[i], oracle.UTF8) == pyref.levenshtein_utf8(a[i], b[i])
    assert oracle.levenshtein("naïve".encode(), "naive".encode(), oracle.UTF8) == 1
    assert oracle.levenshtein("naïve".encode(), "naive".encode(), oracle.WAGNER_FISCHER) == 2


text = st.binary(min_size=0, max_size=48)


@settings(max_examples=200, deadline=None)
@given(text, text, text)
def test_properties(a, b, c):
    from oracle import oracle as orc
    d = lambda x, y: orc.levenshtein(x, y, orc.RAPIDFUZZ_CLASS)
    assert d(a, b) == d(b, a)
    assert d(a, a) == 0
    assert abs(len(a) - len(b)) <= d(a, b) <= max(len(a), len(b))
    assert d(a, c) <= d(a, b) + d(b, c)
    assert d(c + a, c + b) == d(a, b) == d(a + c, b + c)  # common affixes do not matter
    assert d(a, b) == orc.levenshtein(a, b, orc.WAGNER_FISCHER)


def test_cross_matches_pairs(oracle):
    q, c = datasets.words_pairs(12, seed=1)
    m = oracle.levenshtein_cross(q.data, q.offsets, c.data, c.offsets)
    for i in range(12):
        for j in range(12):
            assert m[i, j] == oracle.levenshtein(q[i], c[j])


@pytest.mark.parametrize("open_,extend,conv", [(-2, -2, 0), (-5, -1, 0), (-2, -2, 1), (-5, -1, 1)])
@pytest.mark.parametrize("local", [False, True])
def test_align_scores_against_pure_python(oracle, open_, extend, conv, local):
    btc, costs = oracle.unary_class_costs(2, -1)
    first = open_ if conv == 0 else open_ + extend
    rng = np.random.default_rng(9)
    for _ in range(40):
        la, lb = int(rng.integers(0, 30)), int(rng.integers(0, 30))
        a = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=la))
        b = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=lb))
        want = pyref.align_score(a, b, lambda x, y: 2 if x % 32 == y % 32 else -1, first, extend, local)
        assert oracle.align_score(a, b, btc, costs, open_, extend, conv, local) == want


def test_nw_known_answer(oracle):
    # Classic textbook pair, match +1 / mismatch -1 / gap -1 (Needleman-Wunsch on Wikipedia): score 0
    btc, costs = oracle.unary_class_costs(1, -1)
    assert oracle.align_score(b"GATTACA", b"GCATGCU", btc, costs, -1, -1) == 0
    # Levenshtein is NW with (0, -1, -1), negated
    assert -oracle.align_score(b"kitten", b"sitting", *oracle.unary_class_costs(0, -1), -1, -1) == 3


def test_tape_roundtrip():
    t = BytesTape.from_strings([b"ab", b"", b"cde"])
    assert len(t) == 3 and t[0] == b"ab" and t[1] == b"" and t[2] == b"cde"
    v = t.subview(1, 3)
    assert len(v) == 2 and v[1] == b"cde" and v.data is t.data
