#!/bin/bash
# Round-2 GPU call 15 (session 2): evidence run at HEAD -- GPU tests (all but the full-size file), bench line,
# ncu launch list of one bench step.  Logs are copied to profiles/r2_logs/ afterwards.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q -x --durations=12 --deselect tests/test_gpu_fullsize.py 2>&1 | tail -25 | tee gpurun_out/r2o_tests.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/r2o_bench.err | tail -1 | tee gpurun_out/r2o_bench.json | cut -c1-400
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2o_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/r2o_ncu_bench.log 2>&1; tail -2 gpurun_out/r2o_ncu_bench.log | cut -c1-200
gzip -f gpurun_out/r2o_launches.csv
