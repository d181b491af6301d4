#!/bin/bash
# Round-2 GPU call 30: evidence at HEAD -- complete GPU test suite, smoke(), the default bench line (value, e2e, rooflines, parity, cpu baseline)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== tests (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r3e_tests.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r3e_smoke.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/r3e_bench.err | tail -1 | tee gpurun_out/r3e_bench.json | cut -c1-300
