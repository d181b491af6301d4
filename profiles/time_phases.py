"""CUDA-event timing of each phase of one layer problem (device-resident inputs), per VGG shape class.
    python profiles/time_phases.py [mode]      mode 0 = fp64 Gram, 1 = 3xTF32 tensor-core Gram
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cpb200

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
eng = cpb200.Engine(gram_mode=mode)


def t(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


shapes = [("conv1_2", 64, 64, 224), ("conv2_2", 128, 128, 112), ("conv3_2", 256, 256, 56), ("conv4_2", 512, 512, 28)]
for name, c, n, H in shapes:
    s = cpb200.synth.LayerShape(name, c, n, H, N=5000)
    d = cpb200.synth.make_problem_device(s, 7, eng)
    W2m = d["W2"].reshape(s.n, s.K)
    tg, X = t(lambda: eng.patch_gather(d["fmap"], d["randx"], d["randy"], s.B, s.P, s.k, s.pad, s.stride, relu=True))
    fm_nhwc = d["fmap"].permute(0, 2, 3, 1).contiguous()
    tg2, X2 = t(lambda: eng.patch_gather(fm_nhwc, d["randx"], d["randy"], s.B, s.P, s.k, s.pad, s.stride, relu=True,
                                         layout="nhwc"))
    assert torch.equal(X, X2)
    del fm_nhwc, X2
    gb = 8.0 * s.N * s.K / 1e9
    print("   gather %s: NCHW %.3f ms = %.0f GB/s | NHWC %.3f ms = %.0f GB/s (algorithmic 8NK bytes)"
          % (name, tg, gb / tg * 1e3, tg2, gb / tg2 * 1e3), flush=True)
    tgram, g_full = t(lambda: eng.gram(X, d["feats"], y_bias=d["b2"]))
    tgs, g_s = t(lambda: eng.gram(X, d["feats"], y_bias=d["b2"], rows=d["samples"], want_yy=True, mode=0))
    tgw, g_w = t(lambda: eng.gram(W2m, None, want_B=False, mode=0))
    tb, (Q, qv, yn2) = t(lambda: eng.lasso_build(g_s, g_w, W2m, s.c, 9, s.S))
    lb, rb = cpb200.engine.window(s.rank, .1)
    tsel, res = t(lambda: eng.lasso_select(Q, qv, yn2, float(s.S) * s.n, s.rank, lb, rb, 1e-3, d["seeds"]), reps=2)
    plog = res.probe_log.cpu().numpy()
    npb = int(res.scalars.cpu()[1])
    sweeps = int(plog[:npb, 2].sum())
    idxs = res.idxs.cpu().numpy().astype(bool)
    tls, _ = t(lambda: eng.reconstruct_async(g_full, X, d["feats"], d["b2"], idxs, 9), reps=2)
    steps = sweeps * s.c
    print("%s c=%d K=%d | gather %.3f | gram %.3f | gram_s %.3f | gram_w %.3f | build %.3f | select %.3f (probes %d sweeps %d ~%.0f ns/coord) | ls %.3f ms"
          % (name, c, s.K, tg, tgram, tgs, tgw, tb, tsel, npb, sweeps, 1e6 * tsel / max(1, steps), tls), flush=True)
    del d, X, g_full, g_s, g_w
    torch.cuda.empty_cache()
