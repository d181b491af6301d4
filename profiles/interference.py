"""Does the in-place (zero-copy, PCIe) gather slow down unrelated kernels running next to it?
    python profiles/interference.py
Times conv4_2's LS, Gram and LASSO search alone and while another stream keeps a zero-copy gather running."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cpb200

eng = cpb200.Engine(nstreams=2)
s = cpb200.synth.LayerShape("conv4_2", 512, 512, 28, N=5000)
d = cpb200.synth.make_problem_device(s, 7, eng, pinned_host=True)
s2 = cpb200.synth.LayerShape("conv3_2", 256, 256, 56, N=5000)
d2 = cpb200.synth.make_problem_device(s2, 8, eng, pinned_host=True)
W2m = d["W2"].reshape(s.n, s.K)
X = eng.patch_gather(d["fmap"], d["randx"], d["randy"], s.B, s.P, s.k, s.pad, s.stride, relu=True)
g_full = eng.gram(X, d["feats"], y_bias=d["b2"])
g_s = eng.gram(X, d["feats"], y_bias=d["b2"], rows=d["samples"], want_yy=True, mode=0)
g_w = eng.gram(W2m, None, want_B=False, mode=0)
Q, qv, yn2 = eng.lasso_build(g_s, g_w, W2m, s.c, 9, s.S)
lb, rb = cpb200.engine.window(s.rank, .1)
res = eng.lasso_select(Q, qv, yn2, float(s.S) * s.n, s.rank, lb, rb, 1e-3, d["seeds"])
idxs = res.idxs.cpu().numpy().astype(bool)
torch.cuda.synchronize()
side = torch.cuda.Stream()


def timed(fn, background):
    fn()
    torch.cuda.synchronize()
    if background:
        with torch.cuda.stream(side):
            for _ in range(6):   # ~60 ms of PCIe gathers
                eng.patch_gather(d2["fmap_host"], d2["randx"], d2["randy"], s2.B, s2.P, s2.k, s2.pad, s2.stride)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b)


ops = {
    "ls": lambda: eng.reconstruct_async(g_full, X, d["feats"], d["b2"], idxs, 9),
    "gram": lambda: eng.gram(X, d["feats"], y_bias=d["b2"]),
    "select": lambda: eng.lasso_select(Q, qv, yn2, float(s.S) * s.n, s.rank, lb, rb, 1e-3, d["seeds"]),
    "gather(HBM)": lambda: eng.patch_gather(d["fmap"], d["randx"], d["randy"], s.B, s.P, s.k, s.pad, s.stride),
}
for ctas in (64, ):
    for name, fn in ops.items():
        print("%-12s alone %7.3f ms | next to a zero-copy gather %7.3f ms" % (name, timed(fn, False), timed(fn, True)))
