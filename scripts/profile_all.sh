#!/bin/bash
# Runs on the GPU box (under gpurun): GPU suite, default bench line, launch list of the default bench command, one
# `ncu --set full` capture per kernel family (summarised on the box: reports exceed gpurun_out's 64 MiB cap), probes.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/smi.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --workload ascii256 --pairs 262144 --cpu-seconds 5 > gpurun_out/bench_ascii256.json 2> gpurun_out/bench_ascii256.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --cpu-seconds 1 > gpurun_out/bench_under_ncu.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
cap() { # name, kernel regex, skip, command...
    local name=$1 regex=$2 skip=$3; shift 3
    $NCU -k regex:$regex -s $skip -c 1 -o gpurun_out/$name "$@" > /dev/null 2>&1
    python scripts/ncu_summary.py gpurun_out/$name.ncu-rep gpurun_out/$name.txt > /dev/null 2>&1
    rm -f gpurun_out/$name.ncu-rep
}
cap prof_bitpar_k2_ascii64 lev_bitpar_kernel 3 python bench.py --steps 2 --warmup 3 --cpu-seconds 1
cap prof_bitpar_k8_ascii256 lev_bitpar_kernel 3 python bench.py --workload ascii256 --pairs 262144 --steps 2 --warmup 3 --cpu-seconds 1
cap prof_load_phase_ascii64 load_phase 2 python bench.py --steps 2 --warmup 3 --cpu-seconds 1
cap prof_schedule_count schedule_count 3 python bench.py --steps 2 --warmup 3 --cpu-seconds 1
cap prof_coop_g16_acgt1024 lev_bitpar_coop 2 python scripts/coop_probe.py
cap prof_wavefront_packed_nw1024 align_wavefront 1 python scripts/nw_probe.py
cap prof_wavefront_packed_ascii64 align_wavefront 3 python bench.py --algorithm wavefront --steps 2 --warmup 3 --cpu-seconds 1
python scripts/gpu_probe.py > gpurun_out/probe.log 2>&1
python scripts/unaligned_probe.py > gpurun_out/unaligned_probe.log 2>&1
python scripts/nw_probe.py > gpurun_out/nw_probe.log 2>&1
python scripts/cross_probe.py > gpurun_out/cross_probe.log 2>&1
python scripts/coop_probe.py 400 40000 > gpurun_out/coop_probe.log 2>&1; python scripts/coop_probe.py >> gpurun_out/coop_probe.log 2>&1; python scripts/coop_probe.py 3000 2000 >> gpurun_out/coop_probe.log 2>&1
bash scripts/ncu_source.sh src_bitpar_k2 lev_bitpar_kernel 3 python bench.py --steps 2 --warmup 3 --cpu-seconds 1 > /dev/null 2>&1
python scripts/ncu_lines.py gpurun_out/src_bitpar_k2_source.csv 1.0 > gpurun_out/src_bitpar_k2_lines.txt 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_n1.json
