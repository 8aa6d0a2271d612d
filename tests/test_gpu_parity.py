"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle — bit-exact.

Sizes here are what the oracle finishes in seconds; BASELINE.json's full sizes are covered by
tests/test_gpu_fullsize.py through size-independent properties plus sampled oracle checks."""
import numpy as np
import pytest

import stringwars_b200 as sw
from stringwars_b200 import datasets
from stringwars_b200.tape import BytesTape

pytestmark = pytest.mark.gpu

ALGOS = [sw.ALGO_AUTO, sw.ALGO_WAVEFRONT, sw.ALGO_BITPARALLEL]


@pytest.fixture(scope="module")
def scope():
    s = sw.DeviceScope.gpu_device(0)
    yield s
    s.close()


@pytest.fixture(scope="module")
def lev(scope):
    return sw.LevenshteinDistances(scope, 0, 1, 1, 1)


def _check_pairs(engine, scope, oracle, a, b, algo=sw.ALGO_AUTO):
    engine.set_algorithm(algo)
    got = engine.compute_pairs(scope, a, b)
    want = oracle.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, oracle.RAPIDFUZZ_CLASS, threads=8)
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, f"{bad.size} mismatches, first at {bad[:5]}: got {got[bad[:5]]} want {want[bad[:5]]} kernel {engine.last_kernel}"
    engine.set_algorithm(sw.ALGO_AUTO)
    return got


@pytest.mark.parametrize("algo", ALGOS)
def test_golden_vectors(scope, lev, golden, algo):
    lev.set_algorithm(algo)
    a = BytesTape.from_strings([bytes.fromhex(c["a"]) for c in golden])
    b = BytesTape.from_strings([bytes.fromhex(c["b"]) for c in golden])
    got = lev.compute_pairs(scope, a, b)
    want = np.array([c["bytes"] for c in golden], dtype=np.uint64)
    assert np.array_equal(got, want), np.flatnonzero(got != want)[:10]
    lev.set_algorithm(sw.ALGO_AUTO)


@pytest.mark.parametrize("algo", ALGOS)
def test_config1_words(scope, lev, oracle, algo):
    a, b = datasets.words_pairs(1000)
    _check_pairs(lev, scope, oracle, a, b, algo)


@pytest.mark.parametrize("algo", ALGOS)
def test_config2_ascii64(scope, lev, oracle, algo):
    a, b = datasets.ascii_pairs(50_000, 64)
    _check_pairs(lev, scope, oracle, a, b, algo)
    if algo == sw.ALGO_AUTO:
        assert "bitpar" in lev.last_kernel
    if algo == sw.ALGO_WAVEFRONT:
        assert "wavefront" in lev.last_kernel


@pytest.mark.parametrize("algo", ALGOS)
def test_config2_mutated(scope, lev, oracle, algo):
    a, b = datasets.mutated_pairs(3000, 64, 0.1)
    _check_pairs(lev, scope, oracle, a, b, algo)


@pytest.mark.parametrize("algo", ALGOS)
def test_config3_utf8_lines_bytes(scope, lev, oracle, algo):
    a, b = datasets.utf8_line_pairs(4000)
    _check_pairs(lev, scope, oracle, a, b, algo)


@pytest.mark.parametrize("algo", ALGOS)
def test_config4_ascii256(scope, lev, oracle, algo):
    a, b = datasets.ascii_pairs(20_000, 256)
    _check_pairs(lev, scope, oracle, a, b, algo)


@pytest.mark.parametrize("algo", [sw.ALGO_AUTO, sw.ALGO_WAVEFRONT])
def test_long_strings_row_strips(scope, lev, oracle, algo):
    a, b = datasets.acgt_pairs(300, 1024)
    _check_pairs(lev, scope, oracle, a, b, algo)
    a, b = datasets.ascii_pairs(64, 1500, seed=3)
    _check_pairs(lev, scope, oracle, a, b, algo)


@pytest.mark.parametrize("algo", ALGOS)
def test_edge_cases(scope, lev, oracle, algo):
    strings_a = [b"", b"", b"a", b"abc", b"a" * 31, b"a" * 32, b"a" * 33, b"b" * 64, b"c" * 65, b"x" * 129, b"q" * 257,
                 b"kitten", b"\x00\xff\x00", b"same", bytes(range(256)), b"z" * 300]
    strings_b = [b"", b"abc", b"", b"abc", b"a" * 32, b"a" * 31, b"b" * 33, b"b" * 64, b"c" * 64, b"y" * 130, b"q" * 255,
                 b"sitting", b"\xff\x00\xff", b"same", bytes(range(255, -1, -1)), b"z" * 10]
    a, b = BytesTape.from_strings(strings_a), BytesTape.from_strings(strings_b)
    got = _check_pairs(lev, scope, oracle, a, b, algo)
    assert got[0] == 0 and got[1] == 3 and got[2] == 1 and got[3] == 0 and got[11] == 3


@pytest.mark.parametrize("algo", ALGOS)
def test_ragged_lengths_and_misalignment(scope, lev, oracle, algo):
    rng = np.random.default_rng(17)
    for hi in (20, 70, 140, 300):
        la, lb = rng.integers(0, hi, size=3000), rng.integers(0, hi, size=3000)
        alphabet = np.arange(97, 101, dtype=np.uint8)
        a = datasets._ragged(la, alphabet, rng, False)
        b = datasets._ragged(lb, alphabet, rng, False)
        _check_pairs(lev, scope, oracle, a, b, algo)


def test_empty_batch_and_single_pair(scope, lev):
    empty = BytesTape.from_strings([])
    assert lev.compute_pairs(scope, empty, empty).size == 0
    assert lev.distance(b"kitten", b"sitting") == 3
    assert lev.distance(b"", b"") == 0
    assert lev.distance(b"", b"abc") == 3


def test_subview_offsets_need_not_start_at_zero(scope, lev, oracle):
    a, b = datasets.words_pairs(500)
    va, vb = a.subview(100, 400), b.subview(100, 400)
    got = lev.compute_pairs(scope, va, vb)
    want = oracle.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets)[100:400]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("algo", [sw.ALGO_AUTO, sw.ALGO_WAVEFRONT])
def test_cross_product_compute_into(scope, lev, oracle, algo):
    """`compute` / `compute_into` with disjoint query/candidate halves of one tape (bench.rs:123-148)."""
    lev.set_algorithm(algo)
    tape = datasets.words_pairs(400)[0]
    side = 150
    q, c = tape.subview(0, side), tape.subview(side, 2 * side)
    matrix = lev.compute(scope, q, c)
    want = oracle.levenshtein_cross(q.data, q.offsets, c.data, c.offsets, threads=8)
    assert np.array_equal(matrix, want)
    wide = np.full((side, side + 7), 12345, dtype=np.uint64)  # row stride > n_candidates
    lev.compute_into(scope, q, c, wide)
    assert np.array_equal(wide[:, :side], want) and (wide[:, side:] == 12345).all()
    lev.set_algorithm(sw.ALGO_AUTO)


def test_pinned_tapes(scope, lev, oracle):
    a, b = datasets.ascii_pairs(10_000, 64, pinned=True)
    out = sw.PinnedBuffer(8 * 10_000)
    got = lev.compute_pairs(scope, a, b, out=out.view(np.uint64))
    want = oracle.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, threads=8)
    assert np.array_equal(got, want)


def test_unsupported_cost_model_is_an_error(scope):
    with pytest.raises(sw.Error) as e:
        sw.LevenshteinDistances(scope, 0, 2, 1, 1)
    assert e.value.status_name == "INVALID_ARG"


# ---------------------------------------------------------------- NW / SW scores

@pytest.mark.parametrize("open_,extend", [(-2, -2), (-5, -1)])
@pytest.mark.parametrize("conv", [sw.GAP_STRINGZILLA, sw.GAP_BIO])
@pytest.mark.parametrize("local", [False, True])
def test_alignment_scores(scope, oracle, open_, extend, conv, local):
    btc, costs = sw.unary_class_costs(2, -1)
    cls = sw.SmithWatermanScores if local else sw.NeedlemanWunschScores
    engine = cls(scope, btc, costs, open_, extend).set_gap_convention(conv)
    # same-shape batches take the packed s16x2 kernels (an odd count, lopsided shapes, two row strips at 1024, single pair)
    lopsided = (datasets.acgt_pairs(301, 100, seed=5)[0], datasets.acgt_pairs(301, 257, seed=6)[1])
    tall = (datasets.acgt_pairs(33, 700, seed=7)[0], datasets.acgt_pairs(33, 40, seed=8)[1])
    for a, b in (datasets.acgt_pairs(300, 100, seed=2), datasets.words_pairs(600, seed=2), datasets.acgt_pairs(40, 1024, seed=3), lopsided, tall,
                 datasets.acgt_pairs(1, 64, seed=9)):
        got = engine.compute_pairs(scope, a, b)
        want = oracle.align_pairs(a.data, a.offsets, b.data, b.offsets, btc, costs, open_, extend, conv, local, threads=8)
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, (bad[:5], got[bad[:5]], want[bad[:5]], engine.last_kernel)


def test_alignment_general_substitution_table(scope, oracle):
    rng = np.random.default_rng(4)
    btc = rng.integers(0, 32, size=256).astype(np.uint8)
    costs = rng.integers(-6, 7, size=(32, 32)).astype(np.int8)  # asymmetric on purpose
    a, b = datasets.utf8_line_pairs(300, 20, 200, seed=4)
    for cls, local in ((sw.NeedlemanWunschScores, False), (sw.SmithWatermanScores, True)):
        engine = cls(scope, btc, costs, -3, -1)
        got = engine.compute_pairs(scope, a, b)
        want = oracle.align_pairs(a.data, a.offsets, b.data, b.offsets, btc, costs, -3, -1, 0, local, threads=8)
        assert np.array_equal(got, want)


def test_alignment_cross(scope, oracle):
    btc, costs = sw.unary_class_costs(2, -1)
    engine = sw.NeedlemanWunschScores(scope, btc, costs, -2, -2)
    tape = datasets.acgt_pairs(120, 90, seed=6)[0]
    q, c = tape.subview(0, 60), tape.subview(60, 120)
    got = engine.compute(scope, q, c)
    want = oracle.align_cross(q.data, q.offsets, c.data, c.offsets, btc, costs, -2, -2, threads=8)
    assert got.dtype == np.int64 and np.array_equal(got, want)


# ---------------------------------------------------------------- multi-device scope (single process)

def test_multi_device_scope_matches_single(scope, lev, oracle):
    if sw.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    a, b = datasets.ascii_pairs(30_001, 64, seed=8)  # not divisible by the device count
    single = lev.compute_pairs(scope, a, b)
    multi_scope = sw.DeviceScope.gpu_devices(0, sw.device_count())
    multi = sw.LevenshteinDistances(multi_scope).compute_pairs(multi_scope, a, b)
    assert np.array_equal(single, multi)


# ---------------------------------------------------------------- codepoint-level (UTF-8) engine

@pytest.fixture(scope="module")
def lev_utf8(scope):
    return sw.LevenshteinDistancesUtf8(scope, 0, 1, 1, 1)


def test_utf8_engine_config3_lines(scope, lev_utf8, oracle):
    """`LevenshteinDistancesUtf8` / `rapidfuzz ... distance(a.chars(), b.chars())` (bench.rs:386-399, 437)."""
    a, b = datasets.utf8_line_pairs(3000)
    got = lev_utf8.compute_pairs(scope, a, b)
    want = oracle.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, oracle.UTF8, threads=8)
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, (bad[:5], got[bad[:5]], want[bad[:5]], lev_utf8.last_kernel)
    assert "wavefront" in lev_utf8.last_kernel or "bitpar" in lev_utf8.last_kernel  # renamed to a byte alphabet: bit-vector kernels
    # codepoint-level differs from byte-level on multi-byte text, so this is not accidentally the byte path
    assert (got != oracle.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, threads=8)).any()


def test_utf8_engine_known_answers_and_edges(scope, lev_utf8, oracle):
    sa = ["naïve", "", "日本語テキスト", "😀😃😄", "abc", "Привет, мир", "é" * 40, "a" * 70, "ß", ""]
    sb = ["naive", "añb", "日本語のテキスト", "😀😄", "abc", "Привет мир!", "e" * 41, "ä" * 70, "ss", ""]
    a, b = BytesTape.from_strings(sa), BytesTape.from_strings(sb)
    got = lev_utf8.compute_pairs(scope, a, b)
    want = oracle.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, oracle.UTF8)
    assert np.array_equal(got, want)
    assert list(got[:5]) == [1, 3, 1, 1, 0]


def test_utf8_engine_ascii_shortcut_and_golden(scope, lev_utf8, oracle, golden):
    a, b = datasets.ascii_pairs(20_000, 64, seed=12)
    got = lev_utf8.compute_pairs(scope, a, b)
    assert np.array_equal(got, oracle.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, threads=8))
    assert "bitpar" in lev_utf8.last_kernel  # pure ASCII: bytes are code points, the byte kernels run
    cases = [c for c in golden if "chars" in c]
    ga = BytesTape.from_strings([bytes.fromhex(c["a"]) for c in cases])
    gb = BytesTape.from_strings([bytes.fromhex(c["b"]) for c in cases])
    assert np.array_equal(lev_utf8.compute_pairs(scope, ga, gb), np.array([c["chars"] for c in cases], dtype=np.uint64))


def test_utf8_engine_cross(scope, lev_utf8, oracle):
    tape = datasets.utf8_line_pairs(120, 20, 90, seed=9)[0]
    q, c = tape.subview(0, 60), tape.subview(60, 120)
    got = lev_utf8.compute(scope, q, c)
    want = oracle.levenshtein_cross(q.data, q.offsets, c.data, c.offsets, oracle.UTF8, threads=8)
    assert np.array_equal(got, want)


def test_utf8_engine_cross_batch_dictionary(scope, lev_utf8, oracle):
    """Cross products whose multi-byte symbols fit one batch-wide byte alphabet (at most 127 distinct) are translated to
    byte tapes and run the byte kernels; wider batches keep the 32-bit wavefront. Both bit-exact with the oracle."""
    rng = np.random.default_rng(31)
    words = datasets.alphabetic_words(700, seed=31)
    q, c = words.subview(0, 300), words.subview(300, 700)
    got = lev_utf8.compute(scope, q, c)
    assert np.array_equal(got, oracle.levenshtein_cross(q.data, q.offsets, c.data, c.offsets, oracle.UTF8, threads=8))
    assert "wavefront" not in lev_utf8.last_kernel, lev_utf8.last_kernel
    assert (got != oracle.levenshtein_cross(q.data, q.offsets, c.data, c.offsets, threads=8)).any()  # not the byte distance by accident
    # the tape against itself, longer strings (lines of words), empty strings, u32 offsets
    lines = [" ".join(bytes(words[int(i)]).decode() for i in rng.integers(0, 700, size=int(rng.integers(0, 30)))) for _ in range(90)]
    lines[5] = lines[40] = ""
    tape = BytesTape.from_strings(lines, offset_dtype=np.uint32)
    got = lev_utf8.compute(scope, tape, tape)
    assert np.array_equal(got, oracle.levenshtein_cross(tape.data, tape.offsets, tape.data, tape.offsets, oracle.UTF8, threads=8))
    assert "wavefront" not in lev_utf8.last_kernel, lev_utf8.last_kernel
    # bytes that are not UTF-8: the oracle's lenient split, 58 distinct multi-byte "symbols"
    junk = [bytes(rng.choice([0x41, 0x00, 0x80, 0xBF, 0xC3, 0xF0], size=int(rng.integers(0, 40))).astype(np.uint8)) for _ in range(160)]
    jq, jc = BytesTape.from_strings(junk[:70]), BytesTape.from_strings(junk[70:])
    got = lev_utf8.compute(scope, jq, jc)
    assert np.array_equal(got, oracle.levenshtein_cross(jq.data, jq.offsets, jc.data, jc.offsets, oracle.UTF8, threads=8))
    # one string with 200 distinct ideographs: too wide, the batch runs the wavefront kernel
    wide = BytesTape.from_strings([bytes(words[i]) for i in range(60)] + ["".join(chr(0x4E00 + i) for i in range(200)).encode()])
    got = lev_utf8.compute(scope, wide, q.subview(0, 50))
    assert np.array_equal(got, oracle.levenshtein_cross(wide.data, wide.offsets, q.data, q.offsets[:51], oracle.UTF8, threads=8))
    assert "wavefront" in lev_utf8.last_kernel, lev_utf8.last_kernel


@pytest.mark.parametrize("algo", [sw.ALGO_AUTO, sw.ALGO_WAVEFRONT])
def test_very_long_and_lopsided_pairs(scope, lev, oracle, algo):
    """Lane-group bit-vector kernels up to 4096 pattern bytes, wavefront row strips beyond, short-vs-long pairs."""
    rng = np.random.default_rng(23)
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    la = np.array([300, 512, 513, 1024, 1025, 2048, 2500, 4096, 4097, 5000, 7, 64, 1, 0, 700, 3000])
    lb = np.array([511, 512, 700, 1024, 2000, 2048, 2600, 4096, 4100, 5000, 3000, 4000, 5000, 900, 13, 257])
    a = datasets._ragged(la, alphabet, rng, False)
    b = datasets._ragged(lb, alphabet, rng, False)
    _check_pairs(lev, scope, oracle, a, b, algo)
    # similar strings: small distances through the long kernels
    base = bytes(rng.choice(alphabet, size=3000))
    edits = bytearray(base)
    for pos in rng.choice(3000, size=40, replace=False):
        edits[pos] = ord("N")
    a2, b2 = BytesTape.from_strings([base, base[:2000], base]), BytesTape.from_strings([bytes(edits), bytes(edits), base[5:]])
    got = _check_pairs(lev, scope, oracle, a2, b2, algo)
    assert got[0] == 40 and got[2] == 5


@pytest.mark.parametrize("gen", ["ascii64", "words", "utf8"])
def test_load_phase_kernel_bytesums(scope, lev, gen):
    """The pair-load-phase kernel (TMA bulk copies into shared memory) must deliver every byte of every pair:
    its per-pair byte sums are checked against numpy."""
    import torch
    a, b = {"ascii64": lambda: datasets.ascii_pairs(70_001, 64), "words": lambda: datasets.words_pairs(5000),
            "utf8": lambda: datasets.utf8_line_pairs(3000, 10, 200)}[gen]()
    dev = torch.device("cuda", 0)
    pad = lambda t: torch.cat([torch.from_numpy(t.data.copy()), torch.zeros(64, dtype=torch.uint8)]).to(dev)
    ad, bd = pad(a), pad(b)
    ao, bo = torch.from_numpy(a.offsets.view(np.int64).copy()).to(dev), torch.from_numpy(b.offsets.view(np.int64).copy()).to(dev)
    out = torch.zeros(len(a), dtype=torch.int64, device=dev)
    lev.load_phase_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), len(a), out.data_ptr(),
                          stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    csa = np.concatenate([[0], np.cumsum(a.data.astype(np.int64))])
    csb = np.concatenate([[0], np.cumsum(b.data.astype(np.int64))])
    want = (csa[a.offsets[1:].astype(np.int64)] - csa[a.offsets[:-1].astype(np.int64)]) + (csb[b.offsets[1:].astype(np.int64)] - csb[b.offsets[:-1].astype(np.int64)])
    assert np.array_equal(out.cpu().numpy(), want)


def test_unbounded_lengths(scope, lev, oracle):
    """Byte Levenshtein has no length limit: text streams through shared memory in chunks and patterns beyond
    4096 bytes run in row strips (delta bits parked in global scratch between strips)."""
    rng = np.random.default_rng(31)
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    la = np.array([4097, 8192, 8193, 10_000, 20_000, 50, 3000, 12_289, 0, 9000])
    lb = np.array([4097, 8192, 9000, 12_000, 20_001, 100_000, 70_000, 12_288, 5000, 9000])
    a = datasets._ragged(la, alphabet, rng, False)
    b = datasets._ragged(lb, alphabet, rng, False)
    got = _check_pairs(lev, scope, oracle, a, b, sw.ALGO_AUTO)
    assert got[8] == 5000
    # near-identical long strings: small distance through many strips
    base = bytes(rng.choice(alphabet, size=15_000))
    near = bytearray(base)
    for pos in rng.choice(15_000, size=25, replace=False):
        near[pos] = ord("N")
    a2, b2 = BytesTape.from_strings([base, base]), BytesTape.from_strings([bytes(near), base[7:]])
    got = _check_pairs(lev, scope, oracle, a2, b2, sw.ALGO_AUTO)
    assert list(got) == [25, 7]
    # the wavefront family has no limit either: pairs too long to stage in shared memory stream from global memory
    _check_pairs(lev, scope, oracle, a, b, sw.ALGO_WAVEFRONT)


def test_cross_product_table_reuse_kernel(scope, lev, oracle):
    """Cross products large enough to amortise a per-query table build take `lev_bitpar_cross` (one thread per
    query, candidates swept out of shared memory). Covers every K, empty strings, candidates beyond one chunk."""
    rng = np.random.default_rng(41)
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    words = datasets.words_pairs(1400, seed=3)[0]
    cases = [
        (words.subview(0, 600), words.subview(600, 1300)),                                             # K=1, ragged
        (datasets.ascii_pairs(700, 64, seed=2)[0], datasets.ascii_pairs(450, 64, seed=3)[1]),         # K=2
        (datasets._ragged(rng.integers(0, 129, 700), alphabet, rng, False), datasets._ragged(rng.integers(0, 300, 500), alphabet, rng, False)),  # K=4, empties
        (datasets._ragged(rng.integers(100, 257, 900), alphabet, rng, False), datasets._ragged(rng.integers(900, 2600, 300), alphabet, rng, False)),  # K=8, multi-chunk
    ]
    for q, c in cases:
        got = lev.compute(scope, q, c)
        assert "cross" in lev.last_kernel, lev.last_kernel
        want = oracle.levenshtein_cross(q.data, q.offsets, c.data, c.offsets, oracle.RAPIDFUZZ_CLASS, threads=16)
        bad = np.argwhere(got != want)
        assert bad.size == 0, (bad[:5], lev.last_kernel)
    # row stride wider than the candidate count, untouched padding
    q, c = cases[0]
    wide = np.full((len(q), len(c) + 5), 777, dtype=np.uint64)
    lev.compute_into(scope, q, c, wide)
    assert (wide[:, len(c):] == 777).all() and np.array_equal(wide[:, :len(c)], oracle.levenshtein_cross(q.data, q.offsets, c.data, c.offsets, threads=16))


def test_cpp_harness_twin_runs(tmp_path):
    """The compiled C++ twin of the reference harness's GPU blocks builds against the C ABI and runs every
    variant once (STRINGWARS_TIME=0 is the reference's own smoke mode, bench.rs:34-40)."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "harness"), "-B"], check=True, capture_output=True)
    env = dict(os.environ, STRINGWARS_TIME="0", STRINGWARS_WARMUP="0", STRINGWARS_MAX_TOKENS="20000")
    out = subprocess.run([os.path.join(root, "harness", "bench_similarities")], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    for variant in ("uniform/stringwars_b200::levenshtein<1gpu>", "uniform/stringwars_b200::levenshtein_utf8<1gpu>",
                    "linear/stringwars_b200::needleman_wunsch<1gpu>", "affine/stringwars_b200::smith_waterman<1gpu>"):
        assert variant in out.stdout, out.stdout
    assert "CUPS" in out.stdout and "checksum=" in out.stderr
    # the regex filter of the reference (utils.rs:457-483) applies to our variant names too
    env["STRINGWARS_FILTER"] = "needleman"
    out = subprocess.run([os.path.join(root, "harness", "bench_similarities")], env=env, capture_output=True, text=True, timeout=300)
    assert "needleman_wunsch" in out.stdout and "levenshtein<1gpu>" not in out.stdout


def test_every_alignment_and_length_around_piece_boundaries(scope, lev, oracle):
    """The bit-vector kernels read strings as re-aligned 16-byte register pieces (PieceStream): every combination of
    pattern / text alignment (mod 16) and of lengths around the piece and word boundaries, in all four K classes,
    uniform-length batches (the conflict-free lane rotation) and ragged ones (pieces below a shorter pattern)."""
    rng = np.random.default_rng(123)
    alphabet = np.arange(0x20, 0x7F, dtype=np.uint8)
    # fixed odd lengths: consecutive strings walk through every alignment
    for la, lb in ((5, 9), (15, 16), (17, 33), (31, 31), (32, 47), (33, 64), (63, 63), (64, 65), (65, 100), (127, 129), (128, 250), (129, 131),
                   (250, 250), (255, 256), (256, 301)):
        n = 700
        a = datasets._ragged(np.full(n, la), alphabet, rng, False)
        b = datasets._ragged(np.full(n, lb), alphabet, rng, False)
        _check_pairs(lev, scope, oracle, a, b)
        assert "lev_bitpar<" in lev.last_kernel or "lev_short<" in lev.last_kernel
    # ragged lengths inside each class
    for lo, hi in ((1, 32), (20, 64), (50, 128), (100, 256)):
        n = 3000
        a = datasets._ragged(rng.integers(lo, hi + 1, n), alphabet, rng, False)
        b = datasets._ragged(rng.integers(lo, hi + 40, n), alphabet, rng, False)
        _check_pairs(lev, scope, oracle, a, b)
    # near-identical strings (long carry chains in the wide add) at awkward alignments
    a = datasets._ragged(np.full(500, 251), alphabet, rng, False)
    bdata = a.data.copy()
    bdata[rng.integers(0, bdata.size, 300)] ^= 1
    b = BytesTape.from_arrays(bdata, a.offsets.copy())
    _check_pairs(lev, scope, oracle, a, b)


def test_device_tapes_at_unaligned_base_pointers(scope, lev, oracle):
    """Device-resident tapes whose first byte sits 0..15 bytes into the allocation and whose last byte ends the
    payload: the aligned 16-byte pieces under the first and last string reach outside the tape bytes (never outside
    the allocation), which is the read contract stated in include/stringwars_b200.h."""
    import torch
    rng = np.random.default_rng(5)
    alphabet = np.arange(0x20, 0x7F, dtype=np.uint8)
    dev = torch.device("cuda", 0)
    for shift in (0, 1, 7, 8, 15):
        for lo, hi in ((1, 16), (40, 64), (200, 256), (300, 600)):
            n = 257
            a = datasets._ragged(rng.integers(lo, hi + 1, n), alphabet, rng, False)
            b = datasets._ragged(rng.integers(lo, hi + 1, n), alphabet, rng, False)
            want = oracle.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, oracle.RAPIDFUZZ_CLASS, threads=4)
            bufs = []
            for t in (a, b):
                raw = torch.zeros(shift + t.data.size + 16, dtype=torch.uint8, device=dev)
                raw[shift:shift + t.data.size] = torch.from_numpy(np.array(t.data)).to(dev)
                bufs.append((raw, raw[shift:], torch.from_numpy(np.array(t.offsets).view(np.int64)).to(dev)))
            out = torch.zeros(n, dtype=torch.int64, device=dev)
            lev.compute_pairs_device(bufs[0][1].data_ptr(), bufs[0][2].data_ptr(), bufs[1][1].data_ptr(), bufs[1][2].data_ptr(), n,
                                     out.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            got = out.cpu().numpy().view(np.uint64)
            assert np.array_equal(got, want), (shift, lo, hi, np.flatnonzero(got != want)[:5], lev.last_kernel)


def test_unified_mat_cross_product(scope, lev, oracle):
    """`compute` allocates the harness's UnifiedMat equivalent (page-locked matrix) and fills it like compute_into."""
    tape = datasets.words_pairs(600)[0]
    q, c = tape.subview(0, 300), tape.subview(300, 600)
    m = lev.compute(scope, q, c)
    assert m.shape == (300, 300) and m.dtype == np.uint64
    m2 = sw.unified_mat(300, 300, np.uint64)
    lev.compute_into(scope, q, c, m2)
    assert np.array_equal(np.asarray(m), np.asarray(m2))
    for i in (0, 17, 299):
        want = oracle.levenshtein_pairs(*_row_pairs(q, c, i), oracle.RAPIDFUZZ_CLASS, threads=2)
        assert np.array_equal(np.asarray(m)[i], want)


def _row_pairs(q, c, i):
    """Tapes for (q[i], c[j]) over all j, in the oracle's argument order."""
    qi = q[i]
    a = BytesTape.from_strings([qi] * len(c))
    return a.data, a.offsets, c.data, c.offsets  # subview offsets are absolute into the shared data buffer
