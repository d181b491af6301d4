#!/bin/bash
# Round-2 GPU call 23: step timeline with the tensor-core least squares; full-size parity against the CPU oracle.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== timeline"; timeout 300 python profiles/step_timeline.py 2>&1 | tail -16 | tee gpurun_out/r2w_timeline.log
echo "== fullsize"; timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s 2>&1 | grep -E "relW|rel W|passed|failed|Error" | tee gpurun_out/r2w_fullsize.log
