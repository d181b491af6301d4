#!/bin/bash
# Round-2 GPU call 17: per-kernel times of cp_gram (gen 2) on conv4_2 + ncu --set full of the GEMM kernel.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== launch list"; CP_GRAM_MODE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2q_gram2_launches.csv python profiles/prof_kernels.py gram 3 > /dev/null 2>&1
grep -v "^==" gpurun_out/r2q_gram2_launches.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    if r['Metric Name']=='gpu__time_duration.sum': print(r['Kernel Name'][:50], r['Grid Size'], r['Block Size'], r['Metric Value'], r['Metric Unit'])
" | tail -16
echo "== ncu full"; CP_GRAM_MODE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_tc2_kernel -s 1 -c 1 -o gpurun_out/r2q_gram_tc2_full -f python profiles/prof_kernels.py gram 2 > gpurun_out/r2q_ncu.log 2>&1; tail -3 gpurun_out/r2q_ncu.log
