#!/bin/bash
# Round-2 GPU call 9: TMA im2col v2 (CTAs per SM / stages), stream-priority A/B, Cholesky chain timeline, failing tests.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== gather tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gather" 2>&1 | tail -5 | tee gpurun_out/r2i_test_gather.log
echo "== gather timing"; for cfg in "0 0" "1 0" "2 2" "1 3"; do set -- $cfg; echo "ctas_per_sm<=$1 stages<=$2 (0 = automatic)"; CPB200_GATHER_CTAS_PER_SM=$1 CPB200_GATHER_STAGES=$2 CP_LAYOUT=nhwc timeout 300 python profiles/prof_kernels.py gather 5 2>&1 | grep "^gather"; done | tee gpurun_out/r2i_gather_timing.log
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r3.py -m gpu -q -s -k "conv4_shape or r3" 2>&1 | grep -v "^run for\|^Extracting\|^Reconstruction\|^channel_\|^spatial" | tail -60 > gpurun_out/r2i_tests.log; tail -4 gpurun_out/r2i_tests.log
echo "== chain timeline"; CPB200_LIBRARY=$PWD/channel-pruning_b200/libcpb200_timing.so timeout 300 python profiles/kernel_timeline.py 512 28 2>&1 | tail -12 | tee gpurun_out/r2i_timeline.log
echo "== bench priority A/B"; for v in graded 1 0; do echo "priority=$v"; CPB200_STREAM_PRIORITY=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-parity 2>&1 | tail -1 | cut -c1-200; done | tee gpurun_out/r2i_bench_ab.log
echo "== timeline graded"; timeout 600 python profiles/e2e_breakdown.py 2>&1 | tee gpurun_out/r2i_e2e_breakdown.log | sed -n 1,20p
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/r2i_bench.log | tail -1 | cut -c1-250
