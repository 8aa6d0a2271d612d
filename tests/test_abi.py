"""CPU-side checks of the drop-in boundary: the native library builds, loads, exports every symbol that
include/stringwars_b200.h declares, and fails loudly (never falls back) when there is no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import stringwars_b200 as sw
from stringwars_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    header = open(os.path.join(ROOT, "include", "stringwars_b200.h")).read()
    return sorted(set(re.findall(r"SWB_API\s+[\w\s\*]+?\b(swb_\w+)\s*\(", header)))


def test_library_is_built_in_tree():
    from stringwars_b200 import build
    path = build.build()
    assert path == sw.LIB_PATH and os.path.exists(path)
    assert os.path.commonpath([path, ROOT]) == ROOT  # in-tree, so it ships to the GPU box


def test_exports_every_declared_symbol():
    declared = _declared_symbols()
    assert len(declared) >= 20
    raw = ctypes.CDLL(sw.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert set(declared) == set(_lib.SIGNATURES), "ctypes binding and header disagree"


def test_version_and_device_query():
    L = _lib.lib()
    assert L.swb_version() >= 100
    assert L.swb_device_count() >= 0


def test_product_never_touches_the_oracle():
    """The product package must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "stringwars_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "liblev_oracle" not in text and "orc_" not in text, f


@pytest.mark.skipif(sw.device_count() > 0, reason="needs a box WITHOUT a GPU")
def test_no_gpu_fails_loudly():
    with pytest.raises(sw.Error) as e:
        sw.DeviceScope.gpu_device(0)
    assert e.value.status_name == "NO_DEVICE"
    with pytest.raises(sw.Error):
        sw.multiprocessor_count(0)
    with pytest.raises(sw.Error):
        sw.BytesTape.from_strings([b"a"], pinned=True)


def test_cpu_scope_is_refused():
    with pytest.raises(sw.Error):
        sw.DeviceScope.cpu_cores(4)


def test_null_arguments_are_rejected():
    L = _lib.lib()
    err = ctypes.c_char_p()
    assert L.swb_scope_create(0, 1, None, ctypes.byref(err)) == 2  # INVALID_ARG
    assert b"NULL" in err.value
    assert L.swb_compute_pairs(None, None, None, None, None, 1, None, ctypes.byref(err)) == 2
    out = ctypes.c_void_p()
    assert L.swb_levenshtein_create(None, 0, 1, 1, 1, ctypes.byref(out), ctypes.byref(err)) == 2


def test_unary_class_costs_mirror():
    btc, costs = sw.unary_class_costs(2, -1)
    assert btc[65] == 1 and btc[97] == 1 and btc[255] == 31  # byte % 32 (bench.rs:99-102)
    assert costs[3, 3] == 2 and costs[3, 4] == -1 and costs.dtype == np.int8
