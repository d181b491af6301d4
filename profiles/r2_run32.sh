#!/bin/bash
# Round-2 GPU call 32: ncu --set full of the TMA im2col kernel (DRAM traffic against the 184 MB algorithmic figure)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CP_LAYOUT=nhwc timeout 150 ncu --set full --clock-control none --import-source on -k regex:patch_gather_nhwc_tma -s 1 -c 1 -o gpurun_out/r3g_gather_tma_full -f python profiles/prof_kernels.py gather 2 > gpurun_out/r3g_ncu.log 2>&1; tail -2 gpurun_out/r3g_ncu.log
