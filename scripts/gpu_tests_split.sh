#!/bin/bash
# Runs every GPU test file in its own process (a faulting kernel poisons its CUDA context: the other files still report).
cd "$(dirname "$0")/.."
for f in tests/test_gpu_boundary.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_bench_contract.py; do
  echo "=== $f"
  timeout 900 python -m pytest "$f" -m gpu -q -x 2>&1 | tail -${TAIL:-25}
done
