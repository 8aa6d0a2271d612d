set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "utf8" 2>&1 | tail -5
python scripts/utf8_cross_probe.py 2>&1 | tee gpurun_out/r2h_utf8_cross_probe.log
