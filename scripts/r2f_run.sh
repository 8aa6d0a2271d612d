set -u
mkdir -p gpurun_out
bash scripts/ncu_source.sh src_k8_ragged lev_bitpar_kernel 3 python scripts/utf8_probe.py 100000 > /dev/null 2>&1
python scripts/ncu_lines.py gpurun_out/src_k8_ragged_source.csv 1.0 > gpurun_out/src_k8_ragged_lines.txt; rm -f gpurun_out/src_k8_ragged_sass.csv gpurun_out/src_k8_ragged_source.csv
ncu --set full --clock-control none -k regex:lev_bitpar_kernel -s 3 -c 1 -o gpurun_out/r2f_k8_ragged python scripts/utf8_probe.py 100000 > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/r2f_k8_ragged.ncu-rep gpurun_out/r2f_k8_ragged.txt > /dev/null 2>&1; rm -f gpurun_out/r2f_k8_ragged.ncu-rep
ncu --set full --clock-control none -k regex:lev_bitpar_coop_kernel -s 1 -c 1 -o gpurun_out/r2f_coop_ragged python scripts/utf8_probe.py 100000 > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/r2f_coop_ragged.ncu-rep gpurun_out/r2f_coop_ragged.txt > /dev/null 2>&1; rm -f gpurun_out/r2f_coop_ragged.ncu-rep
cat gpurun_out/src_k8_ragged_lines.txt | head -70
