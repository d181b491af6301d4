#!/bin/bash
# Round-2 GPU call 27: deferred full Gram with the event recorded before the search
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in "A:" "C:CPB200_DEFER_GRAM=0" "A2:" "C2:CPB200_DEFER_GRAM=0"; do
  tag=${v%%:*}; envs=${v#*:}
  echo "== $tag $envs"; env $envs timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-parity 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])" | tee -a gpurun_out/r3a_ab.log
done
echo "== timeline"; timeout 300 python profiles/step_timeline.py 2>&1 | tail -15 | tee gpurun_out/r3a_timeline.log
