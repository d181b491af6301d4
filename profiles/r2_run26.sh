#!/bin/bash
# Round-2 GPU call 26: L2 prefetch of the C tiles in gemm_tc; far updates (128 columns) on the tensor cores; full Gram
# deferred behind the search (lowest-priority stream); tests with the threaded phase 2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in "A:" "B:CPB200_LS_TC_MIN_NN=128" "C:CPB200_DEFER_GRAM=0" "A2:" ; do
  tag=${v%%:*}; envs=${v#*:}
  echo "== $tag $envs"; env $envs timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-parity 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])" | tee -a gpurun_out/r2z_ab.log
done
echo "== timeline"; timeout 300 python profiles/step_timeline.py 2>&1 | tail -15 | tee gpurun_out/r2z_timeline.log
echo "== prof_ls"; timeout 300 python profiles/prof_ls.py 512 28 2>&1 | sed -n 1,9p | tee gpurun_out/r2z_prof_ls.log
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -4 | tee gpurun_out/r2z_tests.log
