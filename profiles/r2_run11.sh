#!/bin/bash
# Round-2 GPU call 11: bulk trailing updates on their own stream (chain timeline, solver timings, tests, bench).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== ls tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_3c.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2k_tests.log
for v in -1 0 64; do echo "== chain timeline rest_ctas=$v"; if [ "$v" = "-1" ]; then unset CPB200_LS_REST_CTAS; else export CPB200_LS_REST_CTAS=$v; fi; CPB200_LIBRARY=$PWD/channel-pruning_b200/libcpb200_timing.so timeout 300 python profiles/kernel_timeline.py 512 28 2>&1 | tail -2; done | tee gpurun_out/r2k_timeline.log
unset CPB200_LS_REST_CTAS
for v in -1 0; do echo "== prof_ls rest_ctas=$v"; if [ "$v" = "-1" ]; then unset CPB200_LS_REST_CTAS; else export CPB200_LS_REST_CTAS=$v; fi; timeout 300 python profiles/prof_ls.py 512 28 2>&1 | sed -n 2,8p; done | tee gpurun_out/r2k_prof_ls.log
unset CPB200_LS_REST_CTAS
timeout 300 python profiles/prof_ls.py 256 56 2>&1 | sed -n 1,8p | tee -a gpurun_out/r2k_prof_ls.log
echo "== bench A/B"; for v in -1 0 64; do echo "rest_ctas=$v"; if [ "$v" = "-1" ]; then unset CPB200_LS_REST_CTAS; else export CPB200_LS_REST_CTAS=$v; fi; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | cut -c1-200; done | tee gpurun_out/r2k_bench_ab.log
unset CPB200_LS_REST_CTAS
echo "== fullsize"; timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s 2>&1 | grep -E "relW|passed|failed|Error" | tee gpurun_out/r2k_test_full.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/r2k_bench.log | tail -1 | cut -c1-250
