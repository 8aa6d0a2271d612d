# final validation of the round-2 tree on one GPU: the GPU suite, smoke, the default bench line, sanitizers over the kernels changed last
set -u
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; tail -c 600 gpurun_out/r2_final_bench_n1.json; tail -3 gpurun_out/r2_final_bench_n1.err
SWB_PROBE_ONLY=utf8 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_probe.py > gpurun_out/r2_final_memcheck_changed.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r2_final_memcheck_changed.log
SWB_PROBE_ONLY=utf8 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_probe.py > gpurun_out/r2_final_racecheck_changed.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r2_final_racecheck_changed.log
tail -3 gpurun_out/r2_final_memcheck_changed.log gpurun_out/r2_final_racecheck_changed.log
