#!/bin/bash
# Round-2 GPU call 1: correctness of the rewritten least squares / LASSO search, phase timings per LAG variant,
# conditioning map, bench line.  Everything lands in gpurun_out/.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt 2>&1
echo "== quick tests"; timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullsize.py 2>&1 | tail -15 | tee gpurun_out/r2_test_quick.log
echo "== phases (LAG 5)"; timeout 300 python profiles/time_phases.py 1 2>&1 | tee gpurun_out/r2_phases_lag5.log | grep -v gather
for l in 3 4; do echo "== phases (LAG $l)"; CPB200_LIBRARY=$PWD/channel-pruning_b200/libcpb200_lag$l.so timeout 300 python profiles/time_phases.py 1 2>&1 | tee gpurun_out/r2_phases_lag$l.log | grep -v gather; done
echo "== conditioning"; timeout 600 python profiles/conditioning_map.py 2>&1 | tee gpurun_out/r2_conditioning.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/r2_bench1.log | tail -3
echo "== fullsize tests"; timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s 2>&1 | tail -25 | tee gpurun_out/r2_test_full.log
