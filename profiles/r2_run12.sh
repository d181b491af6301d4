#!/bin/bash
# Round-2 GPU call 12: cp.async-staged fp64 GEMM (tests, microbenchmark, solver timings, bench A/B).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== gemm / ls tests"; timeout 900 python -m pytest tests/test_gpu_3c.py tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r2l_tests.log
for v in async dmma; do echo "== gemm bench CPB200_GEMM=$v"; CPB200_GEMM=$v timeout 300 python profiles/gemm_bench.py 2>&1 | tail -8; done | tee gpurun_out/r2l_gemm_bench.log
for v in async dmma; do echo "== prof_ls CPB200_GEMM=$v"; CPB200_GEMM=$v timeout 300 python profiles/prof_ls.py 512 28 2>&1 | sed -n 2,8p; done | tee gpurun_out/r2l_prof_ls.log
echo "== chain timeline"; CPB200_LIBRARY=$PWD/channel-pruning_b200/libcpb200_timing.so timeout 300 python profiles/kernel_timeline.py 512 28 2>&1 | tail -2 | tee gpurun_out/r2l_timeline.log
echo "== bench A/B"; for v in async dmma; do echo "gemm=$v"; CPB200_GEMM=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | cut -c1-200; done | tee gpurun_out/r2l_bench_ab.log
echo "== quick tests"; timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py 2>&1 | tail -8 | tee gpurun_out/r2l_test_quick.log
