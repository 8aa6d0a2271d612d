# round-2 follow-up run: ragged-batch timing, sanitizers over the changed kernels
set -u
mkdir -p gpurun_out
python scripts/utf8_probe.py > gpurun_out/r2e_utf8_probe.log 2>&1; cat gpurun_out/r2e_utf8_probe.log
compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_probe.py > gpurun_out/r2e_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r2e_memcheck.log
tail -4 gpurun_out/r2e_memcheck.log
SWB_PROBE_ONLY=utf8 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_probe.py > gpurun_out/r2e_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r2e_racecheck.log
tail -6 gpurun_out/r2e_racecheck.log
