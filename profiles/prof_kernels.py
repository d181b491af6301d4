"""Runs single hot kernels of libcpb200 on BASELINE-sized inputs, for `ncu --set full -k regex:...`.
    python profiles/prof_kernels.py gram|lasso|ls|gather [reps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cpb200

what = sys.argv[1] if len(sys.argv) > 1 else "gram"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = int(os.environ.get("CP_GRAM_MODE", "0"))
eng = cpb200.Engine(gram_mode=mode)
s = cpb200.synth.LayerShape("conv4_2", 512, 512, 28, N=5000)
layout = os.environ.get("CP_LAYOUT", "nhwc")  # HBM layout of the bottom blob (nhwc: TMA gather)
d = cpb200.synth.make_problem_device(s, 5, eng, layout=layout)
X = eng.patch_gather(d["fmap"], d["randx"], d["randy"], s.B, s.P, s.k, s.pad, s.stride, relu=True, layout=layout)
torch.cuda.synchronize()
if what == "gather":
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)
    ts = []
    for it in range(reps):
        flush.fill_(it)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.patch_gather(d["fmap"], d["randx"], d["randy"], s.B, s.P, s.k, s.pad, s.stride, relu=True, out=X, layout=layout)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("gather %s: %s ms -> %.0f GB/s (8NK bytes)" % (layout, ["%.4f" % t for t in ts], 8.0 * s.N * s.K / (min(ts) / 1e3) / 1e9))
elif what == "gram":
    for _ in range(reps):
        eng.gram(X, d["feats"], y_bias=d["b2"], want_sums=True)
else:
    W2m = d["W2"].reshape(s.n, s.K)
    g_full, res = eng.select_channels_async(X, W2m, d["feats"], d["b2"], d["samples"], s.c, 9, s.rank, .1, 1e-3, d["seeds"])
    idxs = res.idxs.cpu().numpy().astype(bool)
    if what == "ls":
        for _ in range(reps):
            eng.reconstruct_async(g_full, X, d["feats"], d["b2"], idxs, 9)
torch.cuda.synchronize()
print("done", what)
