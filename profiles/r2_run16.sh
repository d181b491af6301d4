#!/bin/bash
# Round-2 GPU call 16: first run of the second-generation tensor-core Gram (gram_tc2.cu): accuracy on edge shapes, timing vs gen 1.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== gen2"; timeout 300 python profiles/prof_gram2.py 2>&1 | tail -22 | tee gpurun_out/r2p_gram2.log
echo "== gen1"; CPB200_GRAM_TC=1 timeout 300 python profiles/prof_gram2.py 2>&1 | tail -4 | tee gpurun_out/r2p_gram1.log
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_gram_tc.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2p_tests.log
