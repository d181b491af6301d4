#!/bin/bash
# Round-2 GPU call 19: gram_tc2 with the vectorised split reduction; ncu --set full of the pair kernel; step A/B of the
# refinement residual on the tensor cores (now the second-generation kernel) vs on the FP64 pipe; GPU tests.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== pair"; timeout 180 python profiles/prof_gram2.py 2>&1 | tail -5 | tee gpurun_out/r2s_pair.log
echo "== launch list (pair)"; CP_GRAM_MODE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_gram2_launches.csv python profiles/prof_kernels.py gram 3 > /dev/null 2>&1
grep -v "^==" gpurun_out/r2s_gram2_launches.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    if r['Metric Name']=='gpu__time_duration.sum': print(r['Kernel Name'][:50], r['Grid Size'], r['Block Size'], r['Metric Value'], r['Metric Unit'])
" | tail -6 | tee gpurun_out/r2s_gram2_kernels.log
echo "== ncu full"; CP_GRAM_MODE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_tc2_pair_kernel -s 1 -c 1 -o gpurun_out/r2s_gram_tc2_pair_full -f python profiles/prof_kernels.py gram 2 > gpurun_out/r2s_ncu.log 2>&1; tail -2 gpurun_out/r2s_ncu.log
echo "== bench A (residual auto)"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | tee gpurun_out/r2s_bench_a.json | cut -c1-220
echo "== bench B (residual tc)"; CPB200_LS_RESID=tc timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | tee gpurun_out/r2s_bench_b.json | cut -c1-220
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -4 | tee gpurun_out/r2s_tests.log
