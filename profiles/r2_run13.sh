#!/bin/bash
# Round-2 GPU call 13: transfer plan by simulated makespan (e2e A/B), full bench line.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== e2e breakdown"; timeout 900 python profiles/e2e_breakdown.py 2>&1 | tee gpurun_out/r2m_e2e_breakdown.log | tail -40
echo "== host-path tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host_resident" 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/r2m_bench.log | tail -1 | cut -c1-250
