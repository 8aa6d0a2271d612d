#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of the default bench command + one `ncu --set full` capture per
# kernel family, each summarised on the box (scripts/ncu_summary.py) because gpurun_out/ is capped at 64 MiB.
set -u
mkdir -p gpurun_out
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
cap prof_bitpar_k8_ascii256 lev_bitpar_kernel 3 python bench.py --workload ascii256 --pairs 262144 --steps 2 --warmup 3 --cpu-seconds 1
cap prof_wavefront_g4_ascii64 align_wavefront 3 python bench.py --algorithm wavefront --steps 2 --warmup 3 --cpu-seconds 1
cap prof_load_phase_ascii64 load_phase 2 python bench.py --steps 2 --warmup 3 --cpu-seconds 1
cap prof_wavefront_g32_nw1024 align_wavefront 1 python scripts/nw_probe.py
cap prof_coop_g16_acgt1024 lev_bitpar_coop 1 python scripts/gpu_probe.py
cap prof_schedule_count schedule_count 3 python bench.py --steps 2 --warmup 3 --cpu-seconds 1
ls -la gpurun_out
