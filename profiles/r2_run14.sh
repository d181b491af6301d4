#!/bin/bash
# Round-2 GPU call 14: panel factorisation with FP64-MMA block products (tests, in-kernel timeline, solver timings, bench).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== ls tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_3c.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2n_tests.log
echo "== timeline"; CPB200_LIBRARY=$PWD/channel-pruning_b200/libcpb200_timing.so timeout 300 python profiles/kernel_timeline.py 512 28 2>&1 | tail -10 | tee gpurun_out/r2n_timeline.log
echo "== prof_ls"; timeout 300 python profiles/prof_ls.py 512 28 2>&1 | sed -n 1,8p | tee gpurun_out/r2n_prof_ls.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | cut -c1-200 | tee gpurun_out/r2n_bench_ab.log
echo "== fullsize"; timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s 2>&1 | grep -E "relW|passed|failed|Error" | tee gpurun_out/r2n_test_full.log
