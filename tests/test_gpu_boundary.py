"""GPU tests of the round-2 boundary: typed tape views (u32 offsets, any memory kind), narrow result widths, managed
(`UnifiedAlloc`) tapes and matrices used in place, resident tapes, the multi-device sharding path (several scope entries
on one GPU when the box has one), offsets validation, handle lifetimes, the register kernel for short patterns.
Everything goes through the C ABI and is checked bit-exact against the oracle."""
import ctypes

import numpy as np
import pytest

import stringwars_b200 as sw
from stringwars_b200 import _lib, datasets
from stringwars_b200.tape import BytesTape, unified_mat

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scope():
    s = sw.DeviceScope.gpu_device(0)
    yield s
    s.close()


@pytest.fixture(scope="module")
def lev(scope):
    return sw.LevenshteinDistances(scope, 0, 1, 1, 1)


def _want(oracle, a, b):
    return oracle.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, oracle.RAPIDFUZZ_CLASS, threads=8)


# ------------------------------------------------------------------------------------------ short patterns

@pytest.mark.parametrize("max_len", [8, 16, 24])
def test_short_kernel_every_length_pair(scope, lev, oracle, max_len):
    """lev_short<W>: every (la, lb) with la, lb <= max_len, several alphabets, strings at every byte alignment."""
    rng = np.random.default_rng(max_len)
    strings_a, strings_b = [], []
    for la in range(0, max_len + 1):
        for lb in range(0, max_len + 1):
            for alpha in (2, 4, 26):
                strings_a.append(bytes(rng.integers(97, 97 + alpha, size=la, dtype=np.uint8)))
                strings_b.append(bytes(rng.integers(97, 97 + alpha, size=lb, dtype=np.uint8)))
    a, b = BytesTape.from_strings(strings_a), BytesTape.from_strings(strings_b)
    got = lev.compute_pairs(scope, a, b)
    assert "lev_short" in lev.last_kernel, lev.last_kernel
    assert np.array_equal(got, _want(oracle, a, b))


def test_short_kernel_binary_bytes_and_words(scope, lev, oracle):
    rng = np.random.default_rng(5)
    la, lb = rng.integers(0, 17, size=50_000), rng.integers(0, 17, size=50_000)
    every = np.arange(256, dtype=np.uint8)  # all byte values, including 0x00, 0x7F, 0x80, 0xFF
    a = datasets._ragged(la, every, rng, False)
    b = datasets._ragged(lb, every, rng, False)
    got = lev.compute_pairs(scope, a, b)
    assert "lev_short" in lev.last_kernel
    assert np.array_equal(got, _want(oracle, a, b))
    a, b = datasets.words_pairs(200_000, seed=9)
    assert np.array_equal(lev.compute_pairs(scope, a, b), _want(oracle, a, b))
    # near-identical words (small distances exercise the Mv path)
    words = [bytes(rng.integers(97, 123, size=int(rng.integers(1, 17)), dtype=np.uint8)) for _ in range(20_000)]
    muts = []
    for w in words:
        w2 = bytearray(w)
        if len(w2) and rng.random() < 0.7:
            w2[int(rng.integers(0, len(w2)))] = int(rng.integers(97, 123))
        if rng.random() < 0.3 and len(w2) < 16:
            w2.insert(int(rng.integers(0, len(w2) + 1)), int(rng.integers(97, 123)))
        muts.append(bytes(w2))
    a, b = BytesTape.from_strings(words), BytesTape.from_strings(muts)
    assert np.array_equal(lev.compute_pairs(scope, a, b), _want(oracle, a, b))


# ------------------------------------------------------------------------------------------ typed views

@pytest.mark.parametrize("workload", ["words", "ascii64", "lines"])
def test_u32_offsets_and_narrow_results(scope, lev, oracle, workload):
    a, b = {"words": lambda: datasets.words_pairs(5000), "ascii64": lambda: datasets.ascii_pairs(20_000, 64),
            "lines": lambda: datasets.utf8_line_pairs(2000)}[workload]()
    want = _want(oracle, a, b)
    a32, b32 = a.with_offsets(np.uint32), b.with_offsets(np.uint32)
    assert np.array_equal(lev.compute_pairs(scope, a32, b32), want)
    assert np.array_equal(lev.compute_pairs(scope, a32, b), want)  # mixed widths
    for dtype in (np.uint16, np.uint32):
        out = np.full(len(a), 0xABCD, dtype=dtype)
        lev.compute_pairs(scope, a32, b32, out=out)
        assert np.array_equal(out.astype(np.uint64), want)
    # cross product over u32 tapes into a u16 matrix with a row stride
    q, c = a32.subview(3, 60), b32.subview(10, 150)
    m = np.zeros((len(q), len(c) + 5), dtype=np.uint16)
    lev.compute_into(scope, q, c, m)
    ref = oracle.levenshtein_cross(q.data, q.offsets, c.data, c.offsets, oracle.RAPIDFUZZ_CLASS, threads=8)
    assert np.array_equal(m[:, : len(c)].astype(np.uint64), ref) and (m[:, len(c):] == 0).all()


def test_narrow_scores(scope, oracle):
    btc, costs = sw.unary_class_costs(2, -1)
    a, b = datasets.acgt_pairs(300, 200)
    for engine_cls, local in ((sw.NeedlemanWunschScores, False), (sw.SmithWatermanScores, True)):
        eng = engine_cls(scope, btc, costs, -5, -1)
        want = oracle.align_pairs(a.data, a.offsets, b.data, b.offsets, btc, costs, -5, -1, local=local, threads=8)
        for dtype in (np.int64, np.int32, np.int16):
            out = np.zeros(len(a), dtype=dtype)
            eng.compute_pairs(scope, a, b, out=out)
            assert np.array_equal(out.astype(np.int64), want), dtype


def test_managed_tapes_and_matrix_in_place(scope, lev, oracle):
    """The reference's own memory: `BytesTape<u64, UnifiedAlloc>` + `UnifiedMat` (bench.rs:292-295, 130)."""
    a, b = datasets.ascii_pairs(30_000, 64, seed=3)
    want = _want(oracle, a, b)
    ma = BytesTape.from_arrays(a.data, a.offsets, memory="managed")
    mb = BytesTape.from_arrays(b.data, b.offsets, memory="managed")
    out = unified_mat(1, len(a), np.uint64, memory="managed")
    launches = sw.launch_count()
    for _ in range(3):
        lev.compute_pairs(scope, ma, mb, out=out.reshape(-1))
    assert sw.launch_count() > launches
    assert np.array_equal(np.asarray(out).reshape(-1), want)
    # host tapes into a managed matrix, managed tapes into a host vector
    out2 = unified_mat(1, len(a), np.uint32, memory="managed")
    lev.compute_pairs(scope, a, b, out=out2.reshape(-1))
    assert np.array_equal(np.asarray(out2).reshape(-1).astype(np.uint64), want)
    assert np.array_equal(lev.compute_pairs(scope, ma, mb), want)
    # cross product, managed everything, strided
    q, c = ma.subview(0, 300), mb.subview(100, 900)
    m = unified_mat(300, 1000, np.uint64, memory="managed")
    lev.compute_into(scope, q, c, m)
    ref = oracle.levenshtein_cross(q.data, q.offsets, c.data, c.offsets, oracle.RAPIDFUZZ_CLASS, threads=8)
    assert np.array_equal(np.asarray(m)[:, :800], ref)


def test_device_pointers_through_the_view_call(scope, lev, oracle):
    import torch
    a, b = datasets.ascii_pairs(10_000, 256, seed=11)
    want = _want(oracle, a, b)
    dev = torch.device("cuda", 0)
    ad, bd = torch.from_numpy(a.data.copy()).to(dev), torch.from_numpy(b.data.copy()).to(dev)
    ao = torch.from_numpy(a.offsets.astype(np.int64)).to(dev)
    bo32 = torch.from_numpy(b.offsets.astype(np.int32)).to(dev)
    out = torch.zeros(len(a), dtype=torch.int16, device=dev)
    torch.cuda.synchronize()
    va = _lib.TapeView(ad.data_ptr(), ao.data_ptr(), len(a), 64)
    vb = _lib.TapeView(bd.data_ptr(), bo32.data_ptr(), len(b), 32)
    err = ctypes.c_char_p()
    _lib.check(_lib.lib().swb_compute_pairs_view(lev._h, ctypes.byref(va), ctypes.byref(vb), out.data_ptr(), _lib.RESULT_16, ctypes.byref(err)), err)
    assert np.array_equal(out.cpu().numpy().astype(np.uint64), want)
    # device tapes, host result
    host_out = np.zeros(len(a), dtype=np.uint64)
    _lib.check(_lib.lib().swb_compute_pairs_view(lev._h, ctypes.byref(va), ctypes.byref(vb), host_out.ctypes.data, _lib.RESULT_64, ctypes.byref(err)), err)
    assert np.array_equal(host_out, want)


# ------------------------------------------------------------------------------------------ resident tapes

def test_resident_tapes(scope, lev, oracle):
    a, b = datasets.utf8_line_pairs(3000, seed=2)
    want = _want(oracle, a, b)
    ra, rb = scope.upload(a), scope.upload(b.with_offsets(np.uint32))
    assert len(ra) == len(a)
    assert np.array_equal(lev.compute_pairs(scope, ra, rb), want)
    assert np.array_equal(lev.compute_pairs(scope, ra.subview(100, 2100), rb.subview(100, 2100)), want[100:2100])
    out16 = np.zeros(len(a), dtype=np.uint16)
    lev.compute_pairs(scope, ra, rb, out=out16)
    assert np.array_equal(out16.astype(np.uint64), want)
    m = lev.compute(scope, ra.subview(5, 70), rb.subview(40, 200))
    q, c = a.subview(5, 70), b.subview(40, 200)
    ref = oracle.levenshtein_cross(q.data, q.offsets, c.data, c.offsets, oracle.RAPIDFUZZ_CLASS, threads=8)
    assert np.array_equal(m, ref)
    # a subview whose first offset is not zero, uploaded on its own
    rs = scope.upload(a.subview(500, 900))
    assert np.array_equal(lev.compute_pairs(scope, rs, rb.subview(500, 900)), want[500:900])
    ra.close()
    with pytest.raises(sw.Error):
        lev.compute_pairs(scope, ra, rb)


def test_resident_utf8_engine(scope, oracle):
    a, b = datasets.utf8_line_pairs(1500, seed=4)
    want = oracle.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, oracle.UTF8, threads=8)
    eng = sw.LevenshteinDistancesUtf8(scope)
    ra, rb = scope.upload(a), scope.upload(b)
    assert np.array_equal(eng.compute_pairs(scope, ra, rb), want)
    assert np.array_equal(eng.compute_pairs(scope, a.with_offsets(np.uint32), b), want)


# ------------------------------------------------------------------------------------------ multi-device sharding

def _scope_devices(n):
    have = sw.device_count()
    return [i % have for i in range(n)]


@pytest.mark.parametrize("n_dev", [2, 3, 4, 8])
def test_multi_device_scope_matches_single(oracle, n_dev):
    """`DeviceScope.gpu_devices(0, n)` (the `<{n}gpu>` variants, bench.rs:498-501): sharded results == oracle, including
    non-divisible counts, ragged lengths (byte-balanced cuts), fewer pairs than devices, empty strings. On a box with
    fewer GPUs the scope's entries wrap around the devices there are — same host path, same kernels."""
    multi = sw.DeviceScope.device_list(_scope_devices(n_dev))
    try:
        lev = sw.LevenshteinDistances(multi)
        for a, b in (datasets.ascii_pairs(100_003, 64, seed=21), datasets.utf8_line_pairs(1001, seed=22), datasets.words_pairs(n_dev - 1, seed=23),
                     datasets.words_pairs(1, seed=24)):
            assert np.array_equal(lev.compute_pairs(multi, a, b), _want(oracle, a, b))
        empties = BytesTape.from_strings([b"", b"abc", b"", b"", b"kitten"]), BytesTape.from_strings([b"", b"", b"xyz", b"", b"sitting"])
        assert lev.compute_pairs(multi, *empties).tolist() == [0, 3, 3, 0, 3]
        # cross product: query rows sharded, odd counts
        a, b = datasets.ascii_pairs(203, 64, seed=25)
        m = lev.compute(multi, a, b.subview(0, 77))
        ref = oracle.levenshtein_cross(a.data, a.offsets, b.data, b.offsets[:78], oracle.RAPIDFUZZ_CLASS, threads=8)
        assert np.array_equal(m, ref)
        # resident tapes and managed tapes over the same scope
        a, b = datasets.ascii_pairs(50_001, 256, seed=26)
        want = _want(oracle, a, b)
        assert np.array_equal(lev.compute_pairs(multi, multi.upload(a), multi.upload(b)), want)
        ma, mb = BytesTape.from_arrays(a.data, a.offsets, memory="managed"), BytesTape.from_arrays(b.data, b.offsets, memory="managed")
        assert np.array_equal(lev.compute_pairs(multi, ma, mb), want)
        btc, costs = sw.unary_class_costs(2, -1)
        nw = sw.NeedlemanWunschScores(multi, btc, costs, -2, -2)
        a, b = datasets.acgt_pairs(101, 300)
        assert np.array_equal(nw.compute_pairs(multi, a, b), oracle.align_pairs(a.data, a.offsets, b.data, b.offsets, btc, costs, -2, -2, threads=8))
    finally:
        multi.close()


@pytest.mark.parametrize("n_dev", [1, 2, 3, 8])
def test_peer_load_sharding(oracle, n_dev):
    """swb_compute_pairs_peer: the batch stays on the scope's first device, every device's kernels read it in place (over
    NVLink on a multi-GPU box; the scope's entries wrap around the devices there are otherwise) and write into `out`."""
    import torch
    multi = sw.DeviceScope.device_list(_scope_devices(n_dev))
    try:
        lev = sw.LevenshteinDistances(multi)
        dev = torch.device("cuda", 0)
        for (a, b), rtype, tdtype, bits in ((datasets.ascii_pairs(200_003, 64, seed=51), _lib.RESULT_16, torch.int16, 64),
                                            (datasets.utf8_line_pairs(3001, seed=52), _lib.RESULT_64, torch.int64, 32),
                                            (datasets.words_pairs(max(1, n_dev - 1), seed=53), _lib.RESULT_32, torch.int32, 64)):
            want = _want(oracle, a, b)
            ad, bd = torch.from_numpy(np.concatenate([a.data, np.zeros(64, np.uint8)])).to(dev), torch.from_numpy(np.concatenate([b.data, np.zeros(64, np.uint8)])).to(dev)
            odt = np.int64 if bits == 64 else np.int32
            ao, bo = torch.from_numpy(a.offsets.astype(odt)).to(dev), torch.from_numpy(b.offsets.astype(odt)).to(dev)
            out = torch.zeros(len(a), dtype=tdtype, device=dev)
            torch.cuda.synchronize()
            va, vb = _lib.TapeView(ad.data_ptr(), ao.data_ptr(), len(a), bits), _lib.TapeView(bd.data_ptr(), bo.data_ptr(), len(b), bits)
            for _ in range(2):
                stats = lev.compute_pairs_peer(va, vb, out.data_ptr(), rtype)
            got = out.cpu().numpy().astype(np.int64).astype(np.uint64) & np.uint64((1 << (8 * out.element_size())) - 1)
            assert np.array_equal(got, want), (n_dev, len(a))
            assert stats["ms"] > 0
    finally:
        multi.close()


@pytest.mark.parametrize("n_dev", [2, 3, 8])
def test_scatter_call_with_the_copy_engine_transport(oracle, n_dev):
    """swb_compute_pairs_scatter with SWB_TRANSPORT_COPY_ENGINE: same cuts, chunks and gather as the NCCL form, the chunks
    carried by cudaMemcpyPeerAsync — which also runs when the scope's entries share one GPU."""
    import torch
    multi = sw.DeviceScope.device_list(_scope_devices(n_dev)).set_transport(sw.TRANSPORT_COPY_ENGINE)
    try:
        lev = sw.LevenshteinDistances(multi)
        dev = torch.device("cuda", 0)
        for (a, b), rtype, tdtype in ((datasets.ascii_pairs(600_011, 256, seed=61), _lib.RESULT_16, torch.int16),
                                      (datasets.utf8_line_pairs(4001, seed=62), _lib.RESULT_64, torch.int64),
                                      (datasets.words_pairs(2, seed=63), _lib.RESULT_32, torch.int32)):
            want = _want(oracle, a, b)
            ad, bd = torch.from_numpy(a.data.copy()).to(dev), torch.from_numpy(b.data.copy()).to(dev)
            ao, bo = torch.from_numpy(a.offsets.astype(np.int64)).to(dev), torch.from_numpy(b.offsets.astype(np.int64)).to(dev)
            out = torch.zeros(len(a), dtype=tdtype, device=dev)
            torch.cuda.synchronize()
            va, vb = _lib.TapeView(ad.data_ptr(), ao.data_ptr(), len(a), 64), _lib.TapeView(bd.data_ptr(), bo.data_ptr(), len(b), 64)
            for _ in range(2):
                stats = lev.compute_pairs_scatter(va, vb, out.data_ptr(), rtype)
            got = out.cpu().numpy().astype(np.int64).astype(np.uint64) & np.uint64((1 << (8 * out.element_size())) - 1)
            assert np.array_equal(got, want), (n_dev, len(a))
            assert stats["device_ms"] > 0 and (len(a) < 1000 or stats["sent_bytes"] > 0)
    finally:
        multi.close()


@pytest.mark.skipif(sw.device_count() < 2, reason="the NCCL scatter needs two GPUs")
@pytest.mark.parametrize("n_dev", [2, 4, 8])
def test_nccl_scatter_gather_inside_the_boundary(oracle, n_dev):
    import torch
    if sw.device_count() < n_dev:
        pytest.skip(f"{n_dev} GPUs needed")
    multi = sw.DeviceScope.gpu_devices(0, n_dev)
    try:
        lev = sw.LevenshteinDistances(multi)
        dev = torch.device("cuda", 0)
        for (a, b), rtype, tdtype in ((datasets.ascii_pairs(300_007, 256, seed=31), _lib.RESULT_16, torch.int16),
                                      (datasets.utf8_line_pairs(5003, seed=32), _lib.RESULT_64, torch.int64),
                                      (datasets.words_pairs(3, seed=33), _lib.RESULT_32, torch.int32)):
            want = _want(oracle, a, b)
            ad, bd = torch.from_numpy(a.data.copy()).to(dev), torch.from_numpy(b.data.copy()).to(dev)
            ao, bo = torch.from_numpy(a.offsets.astype(np.int64)).to(dev), torch.from_numpy(b.offsets.astype(np.int64)).to(dev)
            out = torch.zeros(len(a), dtype=tdtype, device=dev)
            torch.cuda.synchronize()
            va, vb = _lib.TapeView(ad.data_ptr(), ao.data_ptr(), len(a), 64), _lib.TapeView(bd.data_ptr(), bo.data_ptr(), len(b), 64)
            for _ in range(2):
                stats = lev.compute_pairs_scatter(va, vb, out.data_ptr(), rtype)
            got = out.cpu().numpy().astype(np.int64).astype(np.uint64) & np.uint64((1 << (8 * out.element_size())) - 1)
            assert np.array_equal(got, want), (n_dev, len(a))
            assert stats["device_ms"] > 0
    finally:
        multi.close()


# ------------------------------------------------------------------------------------------ validation, lifetimes

def test_non_monotonic_offsets_are_refused_not_faulted(scope, lev):
    a, b = datasets.ascii_pairs(10_000, 64, seed=41)
    bad = a.offsets.copy()
    bad[5000] = bad[5000] + np.uint64(1_000_000_000)  # string 4999 runs a gigabyte forward, string 5000 backwards
    broken = BytesTape(a.data, bad)
    for _ in range(2):  # the second call would be the speculative one on the device path
        with pytest.raises(sw.Error) as e:
            lev.compute_pairs(scope, broken, b)
        assert e.value.status_name == "INVALID_ARG"
    with pytest.raises(sw.Error) as e:
        lev.compute_into(scope, b.subview(0, 50), broken.subview(4000, 6000), np.zeros((50, 2000), dtype=np.uint64))
    assert e.value.status_name == "INVALID_ARG"
    # device-resident call: same refusal
    import torch
    dev = torch.device("cuda", 0)
    ad, bd = torch.from_numpy(a.data.copy()).to(dev), torch.from_numpy(b.data.copy()).to(dev)
    ao, bo = torch.from_numpy(bad.astype(np.int64)).to(dev), torch.from_numpy(b.offsets.astype(np.int64)).to(dev)
    out = torch.zeros(len(a), dtype=torch.int64, device=dev)
    good = torch.from_numpy(a.offsets.astype(np.int64)).to(dev)
    torch.cuda.synchronize()
    lev.compute_pairs_device(ad.data_ptr(), good.data_ptr(), bd.data_ptr(), bo.data_ptr(), len(a), out.data_ptr())  # arms the speculation
    torch.cuda.synchronize()
    with pytest.raises(sw.Error) as e:
        lev.compute_pairs_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), len(a), out.data_ptr())
    assert e.value.status_name == "INVALID_ARG"
    torch.cuda.synchronize()
    # the code-point engine reads the tapes in its rename pass, before any statistics: it validates for itself
    utf8 = sw.LevenshteinDistancesUtf8(scope)
    with pytest.raises(sw.Error) as e:
        utf8.compute_pairs_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), len(a), out.data_ptr())
    assert e.value.status_name == "INVALID_ARG"
    with pytest.raises(sw.Error) as e:
        utf8.compute_pairs(scope, broken, b)
    assert e.value.status_name == "INVALID_ARG"
    # ... and so does its cross product (the survey pass): candidates 4900..5100 of the broken tape, offsets in device memory
    matrix = torch.zeros((50, 200), dtype=torch.int64, device=dev)
    with pytest.raises(sw.Error) as e:
        utf8.compute_cross_device(bd.data_ptr(), bo.data_ptr(), 50, ad.data_ptr(), ao.data_ptr() + 4900 * 8, 200, matrix.data_ptr(), 200)
    assert e.value.status_name == "INVALID_ARG"
    utf8.compute_cross_device(bd.data_ptr(), bo.data_ptr(), 50, ad.data_ptr(), good.data_ptr() + 4900 * 8, 200, matrix.data_ptr(), 200)
    torch.cuda.synchronize()
    assert np.array_equal(matrix.cpu().numpy().astype(np.uint64), lev.compute(scope, b.subview(0, 50), a.subview(4900, 5100)))
    torch.cuda.synchronize()
    # and the engines still work afterwards
    assert lev.compute_pairs(scope, a.subview(0, 100), b.subview(0, 100)).shape == (100,)
    assert np.array_equal(utf8.compute_pairs(scope, a.subview(0, 100), b.subview(0, 100)), lev.compute_pairs(scope, a.subview(0, 100), b.subview(0, 100)))


def test_scope_destroyed_before_its_engines_and_tapes():
    s = sw.DeviceScope.gpu_device(0)
    eng = sw.LevenshteinDistances(s)
    a, b = datasets.words_pairs(10)
    tape = s.upload(a)
    err = ctypes.c_char_p()
    _lib.lib().swb_scope_destroy(s._h)  # behind the Python wrappers' back, as a C caller could
    s._h = ctypes.c_void_p()
    va, vb = a.view(), b.view()
    out = np.zeros(10, dtype=np.uint64)
    st = _lib.lib().swb_compute_pairs_view(eng._h, ctypes.byref(va), ctypes.byref(vb), out.ctypes.data, 0, ctypes.byref(err))
    assert st == 2 and b"scope" in err.value
    _lib.lib().swb_engine_destroy(eng._h)
    eng._h = ctypes.c_void_p()
    _lib.lib().swb_tape_destroy(tape._h)
    tape._h = ctypes.c_void_p()


def test_gap_costs_must_fit_int8(scope):
    btc, costs = sw.unary_class_costs(2, -1)
    with pytest.raises(sw.Error) as e:
        sw.NeedlemanWunschScores(scope, btc, costs, -5, -1000)
    assert e.value.status_name == "INVALID_ARG"


def test_large_gap_costs_on_tiny_strings(scope, oracle):
    """extreme (int8) gap costs on a same-shape batch must not wrap inside the packed s16x2 kernels."""
    btc, costs = sw.unary_class_costs(127, -128)
    a = BytesTape.from_strings([b"ACGT", b"AAAA", b"ACGA", b"TTTT", b"ACGT", b"GGGG"])
    b = BytesTape.from_strings([b"AGGT", b"AAAT", b"TCGA", b"TTTT", b"TGCA", b"GAGA"])
    for open_, extend in ((-128, -128), (-128, -1), (-1, -128 + 127), (-100, -128)):
        for cls, local in ((sw.NeedlemanWunschScores, False), (sw.SmithWatermanScores, True)):
            for conv in (sw.GAP_STRINGZILLA, sw.GAP_BIO):
                eng = cls(scope, btc, costs, open_, extend).set_gap_convention(conv)
                want = oracle.align_pairs(a.data, a.offsets, b.data, b.offsets, btc, costs, open_, extend, convention=conv, local=local)
                assert np.array_equal(eng.compute_pairs(scope, a, b), want), (open_, extend, local, conv, eng.last_kernel)


# ------------------------------------------------------------------------------------------ long alignments

def test_alignments_longer_than_shared_memory(scope, oracle):
    """NW / SW (linear, affine, general table) and the code-point engine on pairs far beyond what a block can stage
    (~10 KB strings used to fail with SWB_TOO_LONG): the global-memory variant of the wavefront kernel,
    mixed in one batch with short pairs, bit-exact with the oracle."""
    rng = np.random.default_rng(77)
    lens_a = [30_000, 12_345, 50, 21_000, 1]
    lens_b = [28_500, 40_000, 60, 9, 16_000]
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    a = datasets._ragged(np.array(lens_a), alphabet, rng, False)
    b = datasets._ragged(np.array(lens_b), alphabet, rng, False)
    # make pair 0 near-identical so that long runs of matches and real gaps both occur
    b.data[: 20_000] = a.data[: 20_000]
    btc, costs = sw.unary_class_costs(2, -1)
    rng2 = np.random.default_rng(5)
    table = rng2.integers(-4, 5, size=(32, 32)).astype(np.int8)
    for cls, local in ((sw.NeedlemanWunschScores, False), (sw.SmithWatermanScores, True)):
        for (bt, cs), (op, ex) in (((btc, costs), (-2, -2)), ((btc, costs), (-5, -1)), ((btc, table), (-3, -1))):
            eng = cls(scope, bt, cs, op, ex)
            got = eng.compute_pairs(scope, a, b)
            want = oracle.align_pairs(a.data, a.offsets, b.data, b.offsets, bt, cs, op, ex, local=local, threads=8)
            assert np.array_equal(got, want), (cls.__name__, op, ex, got, want, eng.last_kernel)
    # byte Levenshtein through the wavefront family and the code-point engine on long multi-byte text
    lev = sw.LevenshteinDistances(scope).set_algorithm(sw.ALGO_WAVEFRONT)
    want = oracle.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, oracle.RAPIDFUZZ_CLASS, threads=8)
    assert np.array_equal(lev.compute_pairs(scope, a, b), want)
    words = ["жук", "naïve", "日本語", "x", "ß", "Ω"]
    long_a = " ".join(words[int(i)] for i in rng.integers(0, len(words), size=9000))
    long_b = " ".join(words[int(i)] for i in rng.integers(0, len(words), size=8000))
    ua, ub = BytesTape.from_strings([long_a, "короткая", long_a[:3000]]), BytesTape.from_strings([long_b, long_b[:5000], long_a[:2990] + "Ж"])
    utf8 = sw.LevenshteinDistancesUtf8(scope)
    want = oracle.levenshtein_pairs(ua.data, ua.offsets, ub.data, ub.offsets, oracle.UTF8, threads=8)
    assert np.array_equal(utf8.compute_pairs(scope, ua, ub), want), utf8.last_kernel


def test_codepoint_engine_alphabet_compaction_and_its_fallback(scope, oracle):
    """Pairs whose symbols are renamed to bytes run the bit-vector kernels; a pair whose shorter string holds more than
    255 distinct code points keeps its batch on the 32-bit wavefront kernel. Both bit-exact with the oracle."""
    rng = np.random.default_rng(99)
    utf8 = sw.LevenshteinDistancesUtf8(scope)
    a, b = datasets.utf8_line_pairs(3000, seed=8)
    want = oracle.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, oracle.UTF8, threads=8)
    assert np.array_equal(utf8.compute_pairs(scope, a, b), want)
    assert "bitpar" in utf8.last_kernel or "short" in utf8.last_kernel, utf8.last_kernel
    # symbols shared and not shared between the two strings, 2/3/4-byte code points, repeated symbols, empties
    pool = [chr(c) for c in list(range(0x430, 0x460)) + list(range(0x4E00, 0x4E40)) + list(range(0x1F600, 0x1F620)) + list(range(0x61, 0x7B))]
    sa = ["".join(pool[int(i)] for i in rng.integers(0, len(pool), size=int(rng.integers(0, 90)))) for _ in range(4000)]
    sb = ["".join(pool[int(i)] for i in rng.integers(0, len(pool), size=int(rng.integers(0, 90)))) for _ in range(4000)]
    for i in range(0, 4000, 3):  # near-identical pairs
        chars = list(sa[i])
        if chars:
            chars[int(rng.integers(0, len(chars)))] = pool[int(rng.integers(0, len(pool)))]
        sb[i] = "".join(chars)
    ta, tb = BytesTape.from_strings(sa), BytesTape.from_strings(sb)
    want = oracle.levenshtein_pairs(ta.data, ta.offsets, tb.data, tb.offsets, oracle.UTF8, threads=8)
    assert np.array_equal(utf8.compute_pairs(scope, ta, tb), want), utf8.last_kernel
    # 128..255 distinct multi-byte code points in the shorter string of ONE pair: the second rename tier (a 255-slot
    # dictionary for every symbol) keeps the batch on the byte kernels; arbitrary bytes (not UTF-8 at all) split like the oracle
    mid = "".join(chr(0x4E00 + i) for i in range(200))
    junk = [bytes(rng.choice([0x41, 0x00, 0x7F, 0x80, 0xBF, 0xC3, 0xE4, 0xF0, 0xFF, 0x20], size=int(rng.integers(0, 70))).astype(np.uint8)) for _ in range(400)]
    ta = BytesTape.from_strings([x.encode() for x in sa[:50]] + [mid.encode()] + junk[:200])
    tb = BytesTape.from_strings([x.encode() for x in sb[:50]] + [(mid[::-1] + "ж" + mid[:10]).encode()] + junk[200:])
    want = oracle.levenshtein_pairs(ta.data, ta.offsets, tb.data, tb.offsets, oracle.UTF8, threads=8)
    assert np.array_equal(utf8.compute_pairs(scope, ta, tb), want)
    assert "wavefront" not in utf8.last_kernel, utf8.last_kernel
    # more than 255 distinct code points in the shorter string of ONE pair: whole batch falls back, still exact
    wide = "".join(chr(0x4E00 + i) for i in range(300))
    sa2, sb2 = sa[:50] + [wide], sb[:50] + [wide[::-1] + "ж"]
    ta, tb = BytesTape.from_strings(sa2), BytesTape.from_strings(sb2)
    want = oracle.levenshtein_pairs(ta.data, ta.offsets, tb.data, tb.offsets, oracle.UTF8, threads=8)
    assert np.array_equal(utf8.compute_pairs(scope, ta, tb), want)
    assert "wavefront" in utf8.last_kernel, utf8.last_kernel


# ------------------------------------------------------------------------------------------ software-pipelined kernel

@pytest.mark.parametrize("la,lb,shift", [(64, 64, 0), (64, 100, 0), (90, 64, 3), (64, 64, 7), (64, 17 * 16, 1)])
def test_large_full_pattern_batches_and_lost_bets(scope, oracle, la, lb, shift):
    """Large batches whose every shorter string has exactly 64 symbols (texts longer, either side the shorter one, tapes at
    odd byte alignments), device-resident (first call: regular path; repeated calls: the speculative launch) and resident
    tapes through the host call; then a batch of another shape, where the bet is lost. (Written for the experimental
    software-pipelined kernel, `-DSWB_EXPERIMENTS` + SWB_PIPE=1, which takes exactly these batches; the shipped library
    runs them on lev_bitpar<K=2>.)"""
    import torch
    n = 400_003
    rng = np.random.default_rng(la * 1000 + lb + shift)
    alphabet = np.arange(0x20, 0x7F, dtype=np.uint8)
    a = datasets._ragged(np.full(n, la), alphabet, rng, False)
    b = datasets._ragged(np.full(n, lb), alphabet, rng, False)
    # a third of the pairs nearly identical: small distances, long carry chains
    nb = b.data.reshape(n, lb).copy()
    k = min(la, lb)
    nb[::3, :k] = a.data.reshape(n, la)[::3, :k]
    nb[::3, 5] ^= 1
    b = BytesTape.from_arrays(nb.reshape(-1), b.offsets)
    want = _want(oracle, a, b)
    lev = sw.LevenshteinDistances(scope)
    dev = torch.device("cuda", 0)
    pad = lambda x: torch.cat([torch.zeros(shift, dtype=torch.uint8), torch.from_numpy(x.copy()), torch.zeros(64, dtype=torch.uint8)]).to(dev)
    ad, bd = pad(a.data), pad(b.data)
    ao, bo = torch.from_numpy(a.offsets.astype(np.int64)).to(dev), torch.from_numpy(b.offsets.astype(np.int64)).to(dev)
    out = torch.zeros(n, dtype=torch.int64, device=dev)
    kernels = []
    for _ in range(3):
        out.zero_()
        lev.compute_pairs_device(ad.data_ptr() + shift, ao.data_ptr(), bd.data_ptr() + shift, bo.data_ptr(), n, out.data_ptr())
        torch.cuda.synchronize()
        kernels.append(lev.last_kernel)
        assert np.array_equal(out.cpu().numpy().astype(np.uint64), want), kernels
    assert "lev_bitpar" in kernels[-1], kernels
    got = lev.compute_pairs(scope, scope.upload(a), scope.upload(b))
    assert np.array_equal(got, want) and "lev_bitpar" in lev.last_kernel
    # a batch of another shape right after: the bet is lost, the result is not
    a2, b2 = datasets.ascii_pairs(400_003, 48, seed=5)
    ad2, bd2 = pad(a2.data), pad(b2.data)
    ao2, bo2 = torch.from_numpy(a2.offsets.astype(np.int64)).to(dev), torch.from_numpy(b2.offsets.astype(np.int64)).to(dev)
    lev.compute_pairs_device(ad.data_ptr() + shift, ao.data_ptr(), bd.data_ptr() + shift, bo.data_ptr(), n, out.data_ptr())
    lev.compute_pairs_device(ad2.data_ptr() + shift, ao2.data_ptr(), bd2.data_ptr() + shift, bo2.data_ptr(), n, out.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().astype(np.uint64), _want(oracle, a2, b2))
