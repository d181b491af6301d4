#!/bin/bash
# Round-2 GPU call 29: the full-size property test in tensor-core mode, with the far updates on / off the tensor cores
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== default"; timeout 600 python -m pytest "tests/test_gpu_parity.py::test_full_size_properties_conv4_shape" -m gpu -q -x 2>&1 | grep -E "^E |assert|passed|failed" | head -20 | tee gpurun_out/r3c_a.log
echo "== min_nn 192"; CPB200_LS_TC_MIN_NN=192 timeout 600 python -m pytest "tests/test_gpu_parity.py::test_full_size_properties_conv4_shape" -m gpu -q -x 2>&1 | grep -E "^E |assert|passed|failed" | head -20 | tee gpurun_out/r3c_b.log
