#!/bin/bash
# Round-2 GPU call 10: fp64 GEMM inner loop on DMMA vs DFMA (microbenchmark, tests, solver timings, bench); capped
# grid of the bulk trailing updates (Cholesky chain timeline with and without).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== gemm / ls tests"; timeout 900 python -m pytest tests/test_gpu_3c.py tests/test_gpu_kernels.py -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r2j_tests.log
for v in dmma dfma; do echo "== gemm bench CPB200_GEMM=$v"; CPB200_GEMM=$v timeout 300 python profiles/gemm_bench.py 2>&1 | tail -8; done | tee gpurun_out/r2j_gemm_bench.log
for v in "dmma -1" "dfma -1" "dmma 0" "dmma 64" "dmma 120"; do set -- $v; echo "== prof_ls CPB200_GEMM=$1 rest_ctas=$2 (-1: default 2/3 of the SMs)"; if [ "$2" = "-1" ]; then unset CPB200_LS_REST_CTAS; else export CPB200_LS_REST_CTAS=$2; fi; CPB200_GEMM=$1 timeout 300 python profiles/prof_ls.py 512 28 2>&1 | sed -n 2,6p; done | tee gpurun_out/r2j_prof_ls.log
unset CPB200_LS_REST_CTAS
for v in -1 0; do echo "== chain timeline rest_ctas=$v"; if [ "$v" = "-1" ]; then unset CPB200_LS_REST_CTAS; else export CPB200_LS_REST_CTAS=$v; fi; CPB200_LIBRARY=$PWD/channel-pruning_b200/libcpb200_timing.so timeout 300 python profiles/kernel_timeline.py 512 28 2>&1 | tail -2; done | tee gpurun_out/r2j_timeline.log
unset CPB200_LS_REST_CTAS
for v in dmma dfma; do echo "== phases CPB200_GEMM=$v"; CPB200_GEMM=$v timeout 300 python profiles/time_phases.py 2>&1 | tail -5; done | tee gpurun_out/r2j_phases.log
echo "== bench A/B"; for v in "dmma -1" "dfma -1" "dmma 0"; do set -- $v; echo "gemm=$1 rest_ctas=$2"; if [ "$2" = "-1" ]; then unset CPB200_LS_REST_CTAS; else export CPB200_LS_REST_CTAS=$2; fi; CPB200_GEMM=$1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | cut -c1-200; done | tee gpurun_out/r2j_bench_ab.log
unset CPB200_LS_REST_CTAS
echo "== quick tests"; timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py 2>&1 | tail -8 | tee gpurun_out/r2j_test_quick.log
