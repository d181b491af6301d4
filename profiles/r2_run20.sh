#!/bin/bash
# Round-2 GPU call 20: split-precision tensor-core GEMM for the solver's bulk products (gemm_tc.cu): unit test, solver
# timings and accuracy with it on / off, step A/B, GPU tests.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== gemm_tc unit"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm_tc_split" -s 2>&1 | grep -E "max err|passed|failed|Error|error" | tee gpurun_out/r2t_unit.log
echo "== prof_ls tc"; timeout 300 python profiles/prof_ls.py 512 28 2>&1 | sed -n 1,9p | tee gpurun_out/r2t_prof_ls_tc.log
echo "== prof_ls fp64 bulk"; CPB200_LS_TC=0 timeout 300 python profiles/prof_ls.py 512 28 2>&1 | sed -n 1,9p | tee gpurun_out/r2t_prof_ls_dmma.log
echo "== bench A (LS bulk on TC)"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | tee gpurun_out/r2t_bench_a.json | cut -c1-200
echo "== bench B (LS bulk fp64)"; CPB200_LS_TC=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | tee gpurun_out/r2t_bench_b.json | cut -c1-200
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -6 | tee gpurun_out/r2t_tests.log
