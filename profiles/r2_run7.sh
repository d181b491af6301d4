#!/bin/bash
# Round-2 GPU call 7: tensor-core residual, substitution tiles, data-form staging, R3 golden (wide conv2_2), sweep.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== quick tests"; timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py 2>&1 | tail -40 | tee gpurun_out/r2g_test_quick.log
echo "== r3 debug"; timeout 600 python profiles/r3_debug.py 1 2>&1 | tail -20 | tee gpurun_out/r2g_r3_debug.log
echo "== prof_ls"; timeout 300 python profiles/prof_ls.py 512 28 2>&1 | tee gpurun_out/r2g_prof_ls.log
echo "== fullsize"; timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s 2>&1 | grep -E "relW|passed|failed|Error" | tee gpurun_out/r2g_test_full.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/r2g_bench.log | tail -1 | cut -c1-300
echo "== sweep"; timeout 1500 python bench.py --workload sweep 2>&1 | tee gpurun_out/r2g_sweep.log | grep "^N=" 
echo "== conditioning"; timeout 600 python profiles/conditioning_map.py 2>&1 | tee gpurun_out/r2g_conditioning.log | tail -8
