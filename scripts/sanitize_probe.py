"""Small run through every kernel family, meant to be executed under compute-sanitizer (memcheck / racecheck).
Exits non-zero on any mismatch against the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stringwars_b200 as sw
from stringwars_b200 import datasets
from stringwars_b200.tape import BytesTape
from oracle import oracle as orc

scope = sw.DeviceScope.gpu_device(0)
lev = sw.LevenshteinDistances(scope)
rng = np.random.default_rng(1)
alpha = np.frombuffer(b"ACGT", np.uint8)
ok = True

def check(name, got, want):
    global ok
    good = np.array_equal(got, want)
    ok &= good
    print(f"{name}: {'ok' if good else 'MISMATCH'} ({lev.last_kernel})", flush=True)

cases = {
    "bitpar K=1 ragged": datasets.words_pairs(300),
    "bitpar K=2": datasets.ascii_pairs(300, 64),
    "bitpar K=4 misaligned": (datasets._ragged(rng.integers(65, 129, 200), alpha, rng, False), datasets._ragged(rng.integers(65, 300, 200), alpha, rng, False)),
    "bitpar K=8": datasets.ascii_pairs(100, 256),
    "coop G=8": datasets.ascii_pairs(40, 500),
    "coop G=16": datasets.acgt_pairs(24, 1024),
    "coop G=32 K=2/4 + strips": (datasets._ragged(np.array([2048, 3000, 4096, 4500, 5]), alpha, rng, False), datasets._ragged(np.array([2048, 3100, 4096, 4600, 4800]), alpha, rng, False)),
    "mixed (bucketed)": datasets.utf8_line_pairs(300, 10, 400),
}
ONLY_CHANGED = os.environ.get("SWB_PROBE_ONLY") == "utf8"  # racecheck of the rename pass and the forked class streams alone
if ONLY_CHANGED:
    cases = {"mixed (bucketed)": cases["mixed (bucketed)"]}
long_cases = {} if ONLY_CHANGED else {"coop strips + chunks": (datasets._ragged(np.array([9000, 4097, 30]), alpha, rng, False), datasets._ragged(np.array([9500, 5000, 40000]), alpha, rng, False))}
for name, (a, b) in long_cases.items():
    check(name, lev.compute_pairs(scope, a, b), orc.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, orc.RAPIDFUZZ_CLASS, threads=4))
for name, (a, b) in cases.items():
    want = orc.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, orc.RAPIDFUZZ_CLASS, threads=4)
    check(name, lev.compute_pairs(scope, a, b), want)
    lev.set_algorithm(sw.ALGO_WAVEFRONT)
    check(name + " [wavefront]", lev.compute_pairs(scope, a, b), want)
    lev.set_algorithm(sw.ALGO_AUTO)
u8 = sw.LevenshteinDistancesUtf8(scope)
a, b = datasets.utf8_line_pairs(200, 10, 300)
check("utf8 engine", u8.compute_pairs(scope, a, b), orc.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, orc.UTF8, threads=4))
# second rename tier (more than 127 distinct multi-byte symbols in one pair), strings longer than one 512-byte round, bytes that are not UTF-8
from stringwars_b200.tape import BytesTape
_rng = np.random.default_rng(5)
_mid = "".join(chr(0x4E00 + i) for i in range(200))
_junk = [bytes(_rng.choice([0x41, 0x00, 0x7F, 0x80, 0xBF, 0xC3, 0xE4, 0xF0, 0xFF, 0x20], size=int(_rng.integers(0, 900))).astype(np.uint8)) for _ in range(40)]
a = BytesTape.from_strings([_mid.encode(), ("жук naïve 日本語 " * 90).encode()] + _junk[:20])
b = BytesTape.from_strings([(_mid[::-1] + "ж").encode(), ("naïve жук 日本 " * 80).encode()] + _junk[20:])
check("utf8 engine (second tier, long, lenient)", u8.compute_pairs(scope, a, b), orc.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, orc.UTF8, threads=4))
# code-point cross products: batch-wide byte alphabet (survey -> scan -> translate -> byte cross kernels), same tape, too wide -> wavefront
w = datasets.alphabetic_words(260, seed=3)
q, c = w.subview(0, 100), w.subview(100, 260)
check("utf8 cross (batch dictionary)", u8.compute(scope, q, c), orc.levenshtein_cross(q.data, q.offsets, c.data, c.offsets, orc.UTF8, threads=4))
check("utf8 cross (same tape)", u8.compute(scope, q, q), orc.levenshtein_cross(q.data, q.offsets, q.data, q.offsets, orc.UTF8, threads=4))
wide = BytesTape.from_strings([bytes(w[i]) for i in range(40)] + ["".join(chr(0x4E00 + i) for i in range(200)).encode()])
check("utf8 cross (too wide: wavefront)", u8.compute(scope, wide, q), orc.levenshtein_cross(wide.data, wide.offsets, q.data, q.offsets, orc.UTF8, threads=4))
if ONLY_CHANGED:
    print("ALL OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
btc, costs = sw.unary_class_costs(2, -1)
a, b = datasets.acgt_pairs(30, 700)
for cls, local in ((sw.NeedlemanWunschScores, False), (sw.SmithWatermanScores, True)):
    for o, e in ((-2, -2), (-5, -1)):
        eng = cls(scope, btc, costs, o, e)
        got = eng.compute_pairs(scope, a, b)
        want = orc.align_pairs(a.data, a.offsets, b.data, b.offsets, btc, costs, o, e, 0, local, threads=4)
        ok &= np.array_equal(got, want)
        print(f"{cls.__name__} ({o},{e}): {'ok' if np.array_equal(got, want) else 'MISMATCH'} ({eng.last_kernel})", flush=True)
q = datasets.words_pairs(80)[0]
check("cross", lev.compute(scope, q.subview(0, 40), q.subview(40, 80)), orc.levenshtein_cross(q.data, q.offsets[:41], q.data, q.offsets[40:81]))
q = datasets.words_pairs(1200)[0]  # large enough for the table-reusing cross kernel (tile-buffered result rows)
check("cross (table reuse)", lev.compute(scope, q.subview(0, 512), q.subview(512, 1152)), orc.levenshtein_cross(q.data, q.offsets[:513], q.data, q.offsets[512:1153]))
import torch
dev = torch.device("cuda", 0)
a, b = datasets.ascii_pairs(1000, 64)
pad = lambda t: torch.cat([torch.from_numpy(t.data.copy()), torch.zeros(64, dtype=torch.uint8)]).to(dev)
ad, bd = pad(a), pad(b)
ao, bo = torch.from_numpy(a.offsets.view(np.int64).copy()).to(dev), torch.from_numpy(b.offsets.view(np.int64).copy()).to(dev)
out = torch.zeros(1000, dtype=torch.int64, device=dev)
lev.load_phase_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), 1000, out.data_ptr())
torch.cuda.synchronize()
want = a.data.reshape(1000, 64).astype(np.int64).sum(1) + b.data.reshape(1000, 64).astype(np.int64).sum(1)
ok &= np.array_equal(out.cpu().numpy(), want)
print("load_phase:", "ok" if np.array_equal(out.cpu().numpy(), want) else "MISMATCH", flush=True)
for _ in range(3):  # speculative path (second and third device-resident calls of the same shape)
    lev.compute_pairs_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), 1000, out.data_ptr())
torch.cuda.synchronize()
check("device-resident speculative", out.cpu().numpy().view(np.uint64), orc.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets))
# round 2: register kernel (both orientations), long alignment (global-memory variant), u32 offsets + u16 results,
# managed tapes, resident tapes, multi-entry scope, peer-load call, load phase with several threads per pair
a, b = datasets.words_pairs(2000, seed=3)
want = orc.levenshtein_pairs(a.data, a.offsets, b.data, b.offsets, orc.RAPIDFUZZ_CLASS, threads=4)
check("short (long pattern)", lev.compute_pairs(scope, a, b), want)
a2 = datasets._ragged(rng.integers(0, 20, 500), alpha, rng, False)
b2 = datasets._ragged(rng.integers(30, 90, 500), alpha, rng, False)
check("short (short pattern)", lev.compute_pairs(scope, a2, b2), orc.levenshtein_pairs(a2.data, a2.offsets, b2.data, b2.offsets, orc.RAPIDFUZZ_CLASS, threads=4))
out16 = np.zeros(len(a), np.uint16)
lev.compute_pairs(scope, a.with_offsets(np.uint32), b.with_offsets(np.uint32), out=out16)
check("u32 offsets, u16 results", out16.astype(np.uint64), want)
ma, mb = BytesTape.from_arrays(a.data, a.offsets, memory="managed"), BytesTape.from_arrays(b.data, b.offsets, memory="managed")
check("managed tapes", lev.compute_pairs(scope, ma, mb), want)
check("resident tapes", lev.compute_pairs(scope, scope.upload(a), scope.upload(b)), want)
la, lb = datasets._ragged(np.array([12000, 40]), alpha, rng, False), datasets._ragged(np.array([15000, 11000]), alpha, rng, False)
nw = sw.NeedlemanWunschScores(scope, btc, costs, -5, -1)
got = nw.compute_pairs(scope, la, lb)
wantl = orc.align_pairs(la.data, la.offsets, lb.data, lb.offsets, btc, costs, -5, -1, threads=4)
ok &= np.array_equal(got, wantl)
print(f"long alignment: {'ok' if np.array_equal(got, wantl) else 'MISMATCH'} ({nw.last_kernel})", flush=True)
multi = sw.DeviceScope.device_list([0, 0, 0])
mlev = sw.LevenshteinDistances(multi)
good = np.array_equal(mlev.compute_pairs(multi, a, b), want)
ok &= good
print("multi-entry scope:", "ok" if good else "MISMATCH", flush=True)
from stringwars_b200 import _lib
ad, bd = pad(a), pad(b)
ao, bo = torch.from_numpy(a.offsets.view(np.int64).copy()).to(dev), torch.from_numpy(b.offsets.view(np.int64).copy()).to(dev)
out = torch.zeros(len(a), dtype=torch.int64, device=dev)
torch.cuda.synchronize()
mlev.compute_pairs_peer(_lib.TapeView(ad.data_ptr(), ao.data_ptr(), len(a), 64), _lib.TapeView(bd.data_ptr(), bo.data_ptr(), len(b), 64), out.data_ptr(), _lib.RESULT_64)
good = np.array_equal(out.cpu().numpy().view(np.uint64), want)
ok &= good
print("peer-load call:", "ok" if good else "MISMATCH", flush=True)
a, b = datasets.ascii_pairs(500, 256)
ad, bd = pad(a), pad(b)
ao, bo = torch.from_numpy(a.offsets.view(np.int64).copy()).to(dev), torch.from_numpy(b.offsets.view(np.int64).copy()).to(dev)
out = torch.zeros(500, dtype=torch.int64, device=dev)
lev.load_phase_device(ad.data_ptr(), ao.data_ptr(), bd.data_ptr(), bo.data_ptr(), 500, out.data_ptr())
torch.cuda.synchronize()
want = a.data.reshape(500, 256).astype(np.int64).sum(1) + b.data.reshape(500, 256).astype(np.int64).sum(1)
ok &= np.array_equal(out.cpu().numpy(), want)
print("load_phase 256 B (threads per pair):", "ok" if np.array_equal(out.cpu().numpy(), want) else "MISMATCH", flush=True)
multi.close()
sys.exit(0 if ok else 1)
