#!/bin/bash
# Runs on the GPU box: per-source-line / per-SASS instruction counts and stall samples for one kernel.
# usage: ncu_source.sh <name> <kernel regex> <skip> <command...>; writes gpurun_out/<name>_{source,sass}.csv
set -u
name=$1 regex=$2 skip=$3; shift 3
mkdir -p gpurun_out
ncu --section SourceCounters --section WarpStateStats --import-source on --clock-control none -k regex:$regex -s $skip -c 1 \
    -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
ncu -i gpurun_out/$name.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${name}_source.csv 2>> gpurun_out/$name.log
ncu -i gpurun_out/$name.ncu-rep --page source --csv --print-source sass > gpurun_out/${name}_sass.csv 2>> gpurun_out/$name.log
ls -la gpurun_out/$name*
rm -f gpurun_out/$name.ncu-rep
