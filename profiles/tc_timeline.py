"""In-kernel clock64 timeline of gram_tc_kernel (profiling build: `make -C channel-pruning_b200/csrc timing`).
    CPB200_LIBRARY=channel-pruning_b200/libcpb200_timing.so python profiles/tc_timeline.py [N K n]
Prints, for CTAs blockIdx.x < 64 of row chunk 3 of the X'X launch, cycles from kernel entry to each stage and
the cycles each role spent waiting on its barriers."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cpb200
from cpb200 import _cabi

N, K, n = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (5000, 4608, 512)
eng = cpb200.Engine(gram_mode=1)
X = torch.rand(N, K, device="cuda")
L = ctypes.CDLL(_cabi.LIBRARY)
names = ["entry", "setup_done", "first_raw_full", "first_ops_full", "mma_all_issued", "conv_loop_done", "partial_written", "exit"]
waits = ["producer waits raw_empty", "mma waits ops_full", "converter waits raw_full", "converter waits acc_full (drain)",
         "converter waits ops_empty"]
MASKS = ((0, "normal"),)
for mask, what in MASKS:
    for _ in range(2):
        g = eng.gram(X, None, mode=1)   # G only -> the last gram_tc_kernel launch is the X'X one
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g = eng.gram(X, None, mode=1)
    b.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (64 * 16))()
    assert L.cp_debug_tc_times(buf) == 0
    T = np.array(buf, dtype=np.int64).reshape(64, 16)
    rel = T[:, :8] - T[:, :1]
    print("== %s: cp_gram %.3f ms; N=%d K=%d: cycles from CTA entry (median over 64 CTAs)" % (what, a.elapsed_time(b), N, K))
    print("   " + "  ".join("%s %d" % (nm, np.median(rel[:, i])) for i, nm in enumerate(names)))
    print("   " + "  ".join("%s %d" % (nm, np.median(T[:, i])) for i, nm in zip(range(8, 13), waits)))
