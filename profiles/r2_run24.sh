#!/bin/bash
# Round-2 GPU call 25: phase 2 issued from worker threads (CPB200_PHASE2_THREADS)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in "A:" "B:CPB200_LS_TC_MIN_NN=512" "C:CPB200_LS_REST_CTAS=50" "D:CPB200_LS_REST_CTAS=148" "E:CPB200_LS_TC_MIN_NN=512 CPB200_LS_REST_CTAS=148"; do
  tag=${v%%:*}; envs=${v#*:}
  echo "== $tag $envs"; env $envs timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-parity 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])" | tee -a gpurun_out/r2y_ab.log
done
