#!/bin/bash
# Round-2 GPU call 2: fp64 latency microbenchmark, 3C + refinement tests, LS launch list, bench.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== dp latency"; timeout 120 ./profiles/dp_latency 2>&1 | tee gpurun_out/r2_dp_latency.log
echo "== quick tests"; timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullsize.py 2>&1 | tail -30 | tee gpurun_out/r2b_test_quick.log
echo "== prof_ls"; timeout 300 python profiles/prof_ls.py 512 28 2>&1 | tee gpurun_out/r2b_prof_ls.log
timeout 300 python profiles/prof_ls.py 256 56 2>&1 | tee -a gpurun_out/r2b_prof_ls.log
echo "== ncu launch list of one solve"; timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_ls.csv python profiles/prof_ls.py 512 28 > gpurun_out/r2b_ncu_ls.log 2>&1; tail -2 gpurun_out/r2b_ncu_ls.log
echo "== fullsize conv4_2/conv3_2"; timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s -k "conv4_2 or conv3_2" 2>&1 | grep -E "relW|passed|failed|Error" | tee gpurun_out/r2b_test_full.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/r2b_bench.log | tail -2 | cut -c1-600
