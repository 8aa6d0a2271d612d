#!/bin/bash
# Runs on the GPU box (under gpurun): GPU parity suite, default bench line, launch list, ncu captures of the two
# kernels the bench line reports on. Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( time python bench.py ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --cpu-seconds 1 > gpurun_out/bench_under_ncu.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
cap() { # name, kernel regex, skip, command...
    local name=$1 regex=$2 skip=$3; shift 3
    $NCU -k regex:$regex -s $skip -c 1 -o gpurun_out/$name "$@" > /dev/null 2>&1
    python scripts/ncu_summary.py gpurun_out/$name.ncu-rep gpurun_out/$name.txt > /dev/null 2>&1
    [ "${KEEP_REP:-}" = "$name" ] || rm -f gpurun_out/$name.ncu-rep
}
cap prof_bitpar_k2_ascii64 lev_bitpar_kernel 3 python bench.py --steps 2 --warmup 3 --cpu-seconds 1
cap prof_load_phase_ascii64 load_phase 2 python bench.py --steps 2 --warmup 3 --cpu-seconds 1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_n1.json
