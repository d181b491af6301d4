#!/bin/bash
# Round-2 GPU call 25: phase 2 (reconstructions) issued from worker threads: step time against the thread count, timeline, tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for n in 1 2 4 6 8; do
  echo "== threads $n"; CPB200_PHASE2_THREADS=$n timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-parity 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])" | tee -a gpurun_out/r2y_threads.log
done
echo "== timeline"; timeout 300 python profiles/step_timeline.py 2>&1 | tail -15 | tee gpurun_out/r2y_timeline.log
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -4 | tee gpurun_out/r2y_tests.log
