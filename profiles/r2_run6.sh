#!/bin/bash
# Round-2 GPU call 6: LASSO register-ring update warps, residual split, inversion micro-tiles, data-form kernel, R3 debug.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== quick tests"; timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py --deselect tests/test_gpu_r3.py 2>&1 | tail -40 | tee gpurun_out/r2f_test_quick.log
echo "== r3 debug"; timeout 600 python profiles/r3_debug.py 0 2>&1 | tail -20 | tee gpurun_out/r2f_r3_debug.log
echo "== phases default"; timeout 300 python profiles/time_phases.py 1 2>&1 | tee gpurun_out/r2f_phases_default.log | grep select
echo "== prof_ls"; timeout 300 python profiles/prof_ls.py 512 28 2>&1 | tee gpurun_out/r2f_prof_ls.log
echo "== timeline"; CPB200_LIBRARY=$PWD/channel-pruning_b200/libcpb200_timing.so timeout 300 python profiles/kernel_timeline.py 512 28 2>&1 | tee gpurun_out/r2f_timeline.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/r2f_bench.log | tail -1 | cut -c1-300
echo "== sweep"; timeout 1500 python bench.py --workload sweep 2>&1 | tee gpurun_out/r2f_sweep.log | tail -6 | cut -c1-600
