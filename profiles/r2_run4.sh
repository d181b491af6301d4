#!/bin/bash
# Round-2 GPU call 4: potrf128 v2 / pair-batched far updates / GEMM epilogues, LASSO warp placement + back-off variants,
# truncated solve, 3C + R3.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== quick tests"; timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullsize.py 2>&1 | tail -40 | tee gpurun_out/r2d_test_quick.log
echo "== phases default"; timeout 300 python profiles/time_phases.py 1 2>&1 | tee gpurun_out/r2d_phases_default.log
for v in c0s0 c3s0 c0s32 c3s64 c3s32l3; do echo "== phases $v"; CPB200_LIBRARY=$PWD/channel-pruning_b200/libcpb200_$v.so timeout 300 python profiles/time_phases.py 1 2>&1 | tee gpurun_out/r2d_phases_$v.log; done
echo "== prof_ls"; timeout 300 python profiles/prof_ls.py 512 28 2>&1 | tee gpurun_out/r2d_prof_ls.log
echo "== timeline"; CPB200_LIBRARY=$PWD/channel-pruning_b200/libcpb200_timing.so timeout 300 python profiles/kernel_timeline.py 512 28 2>&1 | tee gpurun_out/r2d_timeline.log
echo "== fullsize conv4_2"; timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s -k "conv4_2" 2>&1 | grep -E "relW|passed|failed|Error" | tee gpurun_out/r2d_test_full.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/r2d_bench.log | tail -1 | cut -c1-400
