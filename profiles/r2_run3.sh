#!/bin/bash
# Round-2 GPU call 3: in-kernel timelines (instrumented build), 3C / R3 tests.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for c in "512 28" "128 112"; do
  echo "== timeline c=$c"; CPB200_LIBRARY=$PWD/channel-pruning_b200/libcpb200_timing.so timeout 300 python profiles/kernel_timeline.py $c 2>&1 | tee -a gpurun_out/r2c_timeline.log
done
echo "== 3C + R3 tests"; timeout 1500 python -m pytest tests/test_gpu_3c.py tests/test_gpu_r3.py -m gpu -q -s 2>&1 | tail -40 | tee gpurun_out/r2c_test_3c.log
