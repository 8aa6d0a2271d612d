"""Code-point engine on the harness's shape (side x side cross product, managed tape and matrix): alphabetic words (batch-wide
byte alphabet -> byte kernels) next to mixed-script words with ideographs (too wide: 32-bit wavefront).
    python scripts/utf8_cross_probe.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stringwars_b200 as sw
from stringwars_b200 import datasets

scope = sw.DeviceScope.gpu_device(0)
utf8, lev = sw.LevenshteinDistancesUtf8(scope), sw.LevenshteinDistances(scope)


def timed(engine, tq, tc, m, budget=0.3):
    for _ in range(3):
        engine.compute_into(scope, tq, tc, m)
    t0, calls = time.perf_counter(), 0
    while time.perf_counter() - t0 < budget:
        engine.compute_into(scope, tq, tc, m)
        calls += 1
    return (time.perf_counter() - t0) / calls * 1e3


for side in (195, 1557):
    alpha = datasets.alphabetic_words(2 * side, seed=14)
    rng = np.random.default_rng(2)
    vocab = datasets._vocabulary(rng, 2 * side)
    mixed = sw.BytesTape.from_strings(vocab)
    for name, words in (("alphabetic", alpha), ("with ideographs", mixed)):
        cp = np.add.reduceat(((words.data & 0xC0) != 0x80).astype(np.int64), words.offsets[:-1].astype(np.int64)).astype(np.float64)
        cut = int(words.offsets[side])
        tq = sw.BytesTape.from_arrays(words.data[:cut], words.offsets[: side + 1], memory="managed")
        tc = sw.BytesTape.from_arrays(words.data[cut:], words.offsets[side:] - words.offsets[side], memory="managed")
        m = sw.unified_mat(side, side, np.uint64, memory="managed")
        ms = timed(utf8, tq, tc, m)
        cells = float(np.outer(cp[:side], cp[side:]).sum())
        ms_b = timed(lev, tq, tc, m)
        print(f"side {side:5d} {name:16s}: code points {ms * 1e3:8.1f} us = {cells / ms / 1e6:8.1f} GCUPS ({utf8.last_kernel}); bytes engine {ms_b * 1e3:8.1f} us", flush=True)
