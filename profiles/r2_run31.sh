#!/bin/bash
# Round-2 GPU call 31: ncu launch list of one bench step at HEAD (every launch with its device time; cold-cache, serialised) + ncu --set full of the solver's tensor-core GEMM
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== ncu launches"; timeout 840 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r3f_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/r3f_ncu_bench.log 2>&1; tail -1 gpurun_out/r3f_ncu_bench.log | cut -c1-160
gzip -f gpurun_out/r3f_launches.csv
echo "== ncu full gemm_tc"; timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair_kernel -s 40 -c 1 -o gpurun_out/r3f_gemm_tc_full -f python profiles/prof_ls.py 512 28 > gpurun_out/r3f_ncu2.log 2>&1; tail -2 gpurun_out/r3f_ncu2.log
