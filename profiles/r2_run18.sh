#!/bin/bash
# Round-2 GPU call 18: gram_tc2 CTA-pair kernel (cta_group::2, 256x256 tiles) vs single-CTA tiles; merged statistic passes.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== pair"; timeout 180 python profiles/prof_gram2.py 2>&1 | tail -20 | tee gpurun_out/r2r_pair.log
echo "== single"; CPB200_GRAM_PAIR=0 timeout 180 python profiles/prof_gram2.py 2>&1 | tail -20 | tee gpurun_out/r2r_single.log
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_gram_tc.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2r_tests.log
echo "== launch list (pair)"; CP_GRAM_MODE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2r_gram2_launches.csv python profiles/prof_kernels.py gram 3 > /dev/null 2>&1
grep -v "^==" gpurun_out/r2r_gram2_launches.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    if r['Metric Name']=='gpu__time_duration.sum': print(r['Kernel Name'][:50], r['Grid Size'], r['Block Size'], r['Metric Value'], r['Metric Unit'])
" | tail -6
