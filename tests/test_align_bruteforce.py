"""Pins the alignment oracle (and, on the GPU box, the CUDA engines) independently of any Gotoh-style recurrence:
oracle/align_bruteforce.c enumerates EVERY alignment of two short strings and scores it by the definition of the gap
model, under both gap conventions of the reference's contenders. Exhaustive over a 3-letter
alphabet, plus literature known answers with citations (tests/golden/alignment_literature.json)."""
import itertools
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (name, match, mismatch or None for a general table, open, extend)
MODELS = [("linear", 2, -1, -2, -2), ("affine", 2, -1, -5, -1), ("steep-open", 1, -2, -7, -1), ("table-affine", None, None, -3, -1),
          ("table-linear", None, None, -2, -2)]


def _tables(oracle, match, mismatch):
    if match is not None:
        return oracle.unary_class_costs(match, mismatch)
    rng = np.random.default_rng(7)
    byte_to_class = (np.arange(256) % 32).astype(np.uint8)
    costs = rng.integers(-4, 6, size=(32, 32)).astype(np.int8)  # asymmetric on purpose: S(a_i, b_j) != S(b_j, a_i)
    return byte_to_class, costs


def _all_strings(alphabet: bytes, max_len: int):
    out = [b""]
    for n in range(1, max_len + 1):
        out += [bytes(t) for t in itertools.product(alphabet, repeat=n)]
    return out


def _literature():
    doc = json.load(open(os.path.join(ROOT, "tests", "golden", "alignment_literature.json")))
    for case in doc["cases"]:
        alphabet = case["alphabet"]
        byte_to_class = np.zeros(256, dtype=np.uint8)
        for i, ch in enumerate(alphabet):
            byte_to_class[ord(ch)] = i
        costs = np.zeros((32, 32), dtype=np.int8)
        if "substitution" in case:
            for pair, v in case["substitution"].items():
                i, j = alphabet.index(pair[0]), alphabet.index(pair[1])
                costs[i, j] = costs[j, i] = v
        else:
            costs[:, :] = case["mismatch"]
            np.fill_diagonal(costs, case["match"])
        yield case, byte_to_class, costs


def test_literature_known_answers(oracle):
    """Durbin et al. ch. 2 (BLOSUM50 sub-table: the general 32x32 class table), Wikipedia NW and SW: oracle and brute
    force both give the published score, under BOTH conventions' spelling of a linear gap."""
    n = 0
    for case, btc, costs in _literature():
        a, b, d, local = case["a"].encode(), case["b"].encode(), case["gap"], case["local"]
        for conv, (op, ex) in ((oracle.GAP_STRINGZILLA, (-d, -d)), (oracle.GAP_BIO, (0, -d))):
            assert oracle.align_score(a, b, btc, costs, op, ex, conv, local) == case["score"], case["source"]
            assert oracle.align_bruteforce(a, b, btc, costs, op, ex, conv, local) == case["score"], case["source"]
        n += 1
    assert n == 4


@pytest.mark.parametrize("model", MODELS, ids=[m[0] for m in MODELS])
def test_oracle_equals_bruteforce_exhaustively(oracle, model):
    """Every ordered pair of strings of length <= 4 (global) / <= 3 (local) over {A, C, G}, both conventions."""
    _, match, mismatch, op, ex = model
    btc, costs = _tables(oracle, match, mismatch)
    for local, max_len in ((False, 4), (True, 3)):
        strings = _all_strings(b"ACG", max_len)
        for conv in (oracle.GAP_STRINGZILLA, oracle.GAP_BIO):
            for a in strings:
                for b in strings:
                    want = oracle.align_bruteforce(a, b, btc, costs, op, ex, conv, local)
                    got = oracle.align_score(a, b, btc, costs, op, ex, conv, local)
                    assert got == want, (a, b, conv, local, got, want)


def test_oracle_equals_bruteforce_on_longer_random_pairs(oracle):
    rng = np.random.default_rng(11)
    btc, costs = _tables(oracle, None, None)
    for _ in range(400):
        a = bytes(rng.integers(65, 85, size=int(rng.integers(0, 8)), dtype=np.uint8))
        b = bytes(rng.integers(65, 85, size=int(rng.integers(0, 8)), dtype=np.uint8))
        # extend no dearer than open: with |extend| > |open| a recurrence may re-open inside a run (two adjacent "runs" in
        # the same string), which the maximal-run definition does not allow — a degenerate model no contender uses
        op = int(rng.integers(-9, 0))
        ex = int(rng.integers(op, 0))
        for conv in (oracle.GAP_STRINGZILLA, oracle.GAP_BIO):
            assert oracle.align_score(a, b, btc, costs, op, ex, conv, False) == oracle.align_bruteforce(a, b, btc, costs, op, ex, conv, False)
        if len(a) <= 5 and len(b) <= 5:
            assert oracle.align_score(a, b, btc, costs, op, ex, 0, True) == oracle.align_bruteforce(a, b, btc, costs, op, ex, 0, True)


def test_conventions_differ_by_one_extend_per_gap_run(oracle):
    """open + (k-1)*extend vs open + k*extend: on a pair whose best alignment has exactly one gap run the scores differ
    by exactly `extend` (this is the difference SURVEY.md §8a-NW flags between the two contenders)."""
    btc, costs = oracle.unary_class_costs(2, -9)
    a, b = b"ACGTACGT", b"ACGTTTACGT"  # one insertion of length 2, everything else matches
    sz = oracle.align_bruteforce(a, b, btc, costs, -5, -1, oracle.GAP_STRINGZILLA, False)
    bio = oracle.align_bruteforce(a, b, btc, costs, -5, -1, oracle.GAP_BIO, False)
    assert sz == 8 * 2 - 5 - 1 and bio == sz - 1


@pytest.mark.gpu
@pytest.mark.parametrize("model", MODELS, ids=[m[0] for m in MODELS])
def test_gpu_engines_equal_bruteforce_exhaustively(oracle, model):
    """The CUDA engines (through the C ABI) against the enumeration itself — no recurrence in between."""
    import stringwars_b200 as sw
    from stringwars_b200.tape import BytesTape
    _, match, mismatch, op, ex = model
    btc, costs = _tables(oracle, match, mismatch)
    scope = sw.DeviceScope.gpu_device(0)
    try:
        for local, max_len, cls in ((False, 4, sw.NeedlemanWunschScores), (True, 3, sw.SmithWatermanScores)):
            strings = _all_strings(b"ACG", max_len)
            pairs = [(a, b) for a in strings for b in strings]
            ta, tb = BytesTape.from_strings([p[0] for p in pairs]), BytesTape.from_strings([p[1] for p in pairs])
            for conv, sw_conv in ((oracle.GAP_STRINGZILLA, sw.GAP_STRINGZILLA), (oracle.GAP_BIO, sw.GAP_BIO)):
                want = np.array([oracle.align_bruteforce(a, b, btc, costs, op, ex, conv, local) for a, b in pairs], dtype=np.int64)
                eng = cls(scope, btc, costs, op, ex).set_gap_convention(sw_conv)
                got = eng.compute_pairs(scope, ta, tb)
                bad = np.flatnonzero(got != want)
                assert bad.size == 0, (pairs[bad[0]], int(got[bad[0]]), int(want[bad[0]]), eng.last_kernel)
        # same-shape batches reach the packed s16x2 kernels: every pair of 4-symbol strings
        fours = [s for s in _all_strings(b"ACG", 4) if len(s) == 4]
        pairs = [(a, b) for a in fours for b in fours]
        ta, tb = BytesTape.from_strings([p[0] for p in pairs]), BytesTape.from_strings([p[1] for p in pairs])
        for local, cls in ((False, sw.NeedlemanWunschScores), (True, sw.SmithWatermanScores)):
            want = np.array([oracle.align_bruteforce(a, b, btc, costs, op, ex, 0, local) for a, b in pairs], dtype=np.int64)
            eng = cls(scope, btc, costs, op, ex)
            assert np.array_equal(eng.compute_pairs(scope, ta, tb), want), eng.last_kernel
        for case, lbtc, lcosts in _literature():
            d = case["gap"]
            cls = sw.SmithWatermanScores if case["local"] else sw.NeedlemanWunschScores
            eng = cls(scope, lbtc, lcosts, -d, -d)
            assert int(eng.compute_pairs(scope, [case["a"]], [case["b"]])[0]) == case["score"], case["source"]
    finally:
        scope.close()
