"""Builds the in-tree native library `stringwars_b200/lib/libstringwars_b200.so` for sm_100a with nvcc.

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box with the
repo snapshot. The library is several translation units (csrc/engine_internal.h lists them) compiled in
parallel and linked once. Run as `python -m stringwars_b200.build [--force] [--experiments]`.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libstringwars_b200.so")
SOURCES = ["engine.cu", "pipeline.cu", "launch_lev.cu", "launch_align.cu", "launch_packed.cu", "multi_gpu.cu"]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
]


def _headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h"))) + [
        os.path.join(PKG_DIR, "..", "include", "stringwars_b200.h")]


def _newest_header() -> float:
    return max(os.path.getmtime(h) for h in _headers())


def _obj(src: str) -> str:
    return os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")


def _stale_objects(force: bool):
    newest = _newest_header()
    stale = []
    for src in SOURCES:
        obj = _obj(src)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(newest, os.path.getmtime(os.path.join(CSRC, src))):
            stale.append(src)
    return stale


def build(force: bool = False, verbose: bool = False, experiments: bool = False) -> str:
    """Compile what is out of date and link. `experiments=True` adds -DSWB_EXPERIMENTS (A/B environment switches that
    the shipped library does not read)."""
    stale = _stale_objects(force)
    if not stale and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(_obj(s)) for s in SOURCES):
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    flags = NVCC_FLAGS + (["-DSWB_EXPERIMENTS"] if experiments else [])

    def compile_one(src: str):
        cmd = [nvcc, *flags, "-c", "-o", _obj(src), os.path.join(CSRC, src)]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src} ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
        return proc.stderr

    with ThreadPoolExecutor(max_workers=min(len(stale), os.cpu_count() or 1) or 1) as pool:
        for warnings in pool.map(compile_one, stale):
            if verbose and warnings.strip():
                print(warnings, file=sys.stderr)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *[_obj(s) for s in SOURCES], "-ldl"]
    if verbose:
        print(" ".join(link), file=sys.stderr)
    proc = subprocess.run(link, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"link failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, experiments="--experiments" in sys.argv))
