#!/bin/bash
# Round-2 GPU call 8: TMA im2col (tests, timing, ncu), staged R3 parity, residual-mode A/B of the bench, step timeline.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== gather tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gather" 2>&1 | tail -30 | tee gpurun_out/r2h_test_gather.log
echo "== gather timing"; for l in nhwc nchw; do CP_LAYOUT=$l timeout 300 python profiles/prof_kernels.py gather 5 2>&1 | tail -2; done | tee gpurun_out/r2h_gather_timing.log
echo "== quick tests"; timeout 1500 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_fullsize.py 2>&1 | grep -v "^run for\|^Extracting\|^Reconstruction\|^channel_\|^spatial" | tail -150 > gpurun_out/r2h_test_quick.log; tail -5 gpurun_out/r2h_test_quick.log
echo "== bench (nhwc, tc residual)"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/r2h_bench.log | tail -1 | cut -c1-250
echo "== bench A/B"; for v in tc fp64; do CPB200_LS_RESID=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-parity 2>&1 | tail -1 | cut -c1-200; done | tee gpurun_out/r2h_bench_ab.log
echo "== timeline"; timeout 600 python profiles/e2e_breakdown.py 2>&1 | tee gpurun_out/r2h_e2e_breakdown.log | tail -40
echo "== ncu gather"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:patch_gather_nhwc_tma -c 2 -o gpurun_out/r2h_gather_tma -f python profiles/prof_kernels.py gather 2 2>&1 | tail -3
