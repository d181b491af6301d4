"""One LASSO selection (cp_lasso_select) on a VGG-shaped problem, for ncu source-level profiling.
    python profiles/prof_lasso.py <c> <H>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cpb200

c = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 56
eng = cpb200.Engine()
s = cpb200.synth.LayerShape("L", c, c, H, N=5000)
d = cpb200.synth.make_problem_device(s, 7, eng)
W2m = d["W2"].reshape(s.n, s.K)
X = eng.patch_gather(d["fmap"], d["randx"], d["randy"], s.B, s.P, s.k, s.pad, s.stride, relu=True)
g_s = eng.gram(X, d["feats"], y_bias=d["b2"], rows=d["samples"], want_yy=True, mode=0)
g_w = eng.gram(W2m, None, want_B=False, mode=0)
Q, qv, yn2 = eng.lasso_build(g_s, g_w, W2m, s.c, 9, s.S)
lb, rb = cpb200.engine.window(s.rank, .1)
res = eng.lasso_select(Q, qv, yn2, float(s.S) * s.n, s.rank, lb, rb, 1e-3, d["seeds"])
torch.cuda.synchronize()
print("probes", int(res.scalars.cpu()[1]))
