#!/bin/bash
# Round-2 GPU call 22: backward substitution on the tensor cores (transposed operand preparation); launch list of one
# reconstruction; step; tests.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== gemm_tc unit"; timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm_tc_split" -s 2>&1 | grep -E "max err|passed|failed|Error|error" | tee gpurun_out/r2v_unit.log
echo "== prof_ls tc"; timeout 300 python profiles/prof_ls.py 512 28 2>&1 | sed -n 1,9p | tee gpurun_out/r2v_prof_ls_tc.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | tee gpurun_out/r2v_bench.json | cut -c1-200
echo "== launch list of prof_ls"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2v_prof_ls_launches.csv python profiles/prof_ls.py 512 28 > /dev/null 2>&1; gzip -f gpurun_out/r2v_prof_ls_launches.csv
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -6 | tee gpurun_out/r2v_tests.log
