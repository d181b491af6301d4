"""In-kernel clock64 timelines from the instrumented build (make -C channel-pruning_b200/csrc timing):
   * potrf128 (ls.cu): cycles per phase of the first 128-wide panel of a conv4_2-sized factorisation
   * lasso_select_kernel (lasso.cu): per coordinate step, where the chain warp and an update warp spend their time
     and how long a shared-memory hand-off between them takes
    CPB200_LIBRARY=channel-pruning_b200/libcpb200_timing.so python profiles/kernel_timeline.py [c] [H]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cffi
import numpy as np
import torch

import cpb200

c = int(sys.argv[1]) if len(sys.argv) > 1 else 512
H = int(sys.argv[2]) if len(sys.argv) > 2 else 28
ffi2 = cffi.FFI()
ffi2.cdef("int cp_debug_ls_times(long long*); int cp_debug_lasso_times(long long*); int cp_debug_ls_chain(long long*);")
dbg = ffi2.dlopen(cpb200._cabi.LIBRARY)

eng = cpb200.Engine()
s = cpb200.synth.LayerShape("L", c, c, H, N=5000)
d = cpb200.synth.make_problem_device(s, 7, eng)
W2m = d["W2"].reshape(s.n, s.K)
X = eng.patch_gather(d["fmap"], d["randx"], d["randy"], s.B, s.P, s.k, s.pad, s.stride, relu=True)
g_full = eng.gram(X, d["feats"], y_bias=d["b2"])
g_s = eng.gram(X, d["feats"], y_bias=d["b2"], rows=d["samples"], want_yy=True, mode=0)
g_w = eng.gram(W2m, None, want_B=False, mode=0)
Q, qv, yn2 = eng.lasso_build(g_s, g_w, W2m, s.c, 9, s.S)
lb, rb = cpb200.engine.window(s.rank, .1)
for _ in range(2):
    res = eng.lasso_select(Q, qv, yn2, float(s.S) * s.n, s.rank, lb, rb, 1e-3, d["seeds"])
torch.cuda.synchronize()
buf = ffi2.new("long long[]", 384 * 8)
assert dbg.cp_debug_lasso_times(buf) == 0
T = np.array(list(buf), dtype=np.int64).reshape(384, 8)
n = min(384, c) - 8
T = T[4:n]
ch_top, ch_fet, ch_conf, ch_pub = T[:, 0], T[:, 1], T[:, 2], T[:, 3]
up_poll, up_got, up_pub, up_end = T[:, 4], T[:, 5], T[:, 6], T[:, 7]
print("LASSO c=%d (third sweep of the first fit, steps 4..%d), cycles:" % (c, n))
print("  chain warp : step period %.0f | fetch of next operands %.0f | wait for x (tag confirm) %.0f | arithmetic + publish %.0f | loop tail %.0f"
      % (np.diff(ch_top).mean(), (ch_fet - ch_top).mean(), (ch_conf - ch_fet).mean(), (ch_pub - ch_conf).mean(),
         (ch_top[1:] - ch_pub[:-1]).mean()))
print("  update warp: iteration period %.0f | waiting for delta %.0f | update + publish %.0f | prefetch + row load %.0f"
      % (np.diff(up_poll).mean(), (up_got - up_poll).mean(), (up_pub - up_got).mean(), (up_end - up_pub).mean()))
print("  hand-off   : delta published by the chain -> seen by the update warp %.0f cycles (min %d, max %d)"
      % ((up_got - ch_pub).mean(), (up_got - ch_pub).min(), (up_got - ch_pub).max()))
lag = cpb200._cabi  # noqa
print("  update warp runs %.1f steps behind the chain on average" % (((up_got[:, None] > ch_pub[None, :]).sum(1) - np.arange(len(up_got)) - 1).mean()))

idxs = res.idxs.cpu().numpy().astype(bool)
cols = eng._cols_device(idxs, 9, s.K)
for _ in range(2):
    eng.ls_solve(g_full, cols)
torch.cuda.synchronize()
buf2 = ffi2.new("long long[]", 32)
assert dbg.cp_debug_ls_times(buf2) == 0
t = np.array(list(buf2), dtype=np.int64)
print("potrf128 (first panel), cycles: total %d" % (t[22] - t[0]))
print("  load %d" % (t[1] - t[0]))
for sp in range(4):
    a = t[2 + 4 * sp]
    line = "  sub-panel %d: potrf32 (one warp) %d" % (sp, t[3 + 4 * sp] - a)
    if sp < 3:
        line += " | row solves %d | trailing update %d" % (t[4 + 4 * sp] - t[3 + 4 * sp], t[5 + 4 * sp] - t[4 + 4 * sp])
    print(line)
print("  ratio reduce + L write-out %d | 32x32 inverses %d | off-diagonal inverse blocks %d | Linv write-out %d"
      % (t[19] - t[18], t[20] - t[19], t[21] - t[20], t[22] - t[21]))

buf3 = ffi2.new("long long[]", 128)
assert dbg.cp_debug_ls_chain(buf3) == 0
ch = np.array(list(buf3), dtype=np.int64).reshape(64, 2)
npan = (len(cols) + 127) // 128
ch = ch[:npan]
dur = (ch[:, 1] - ch[:, 0]) / 1e3
gap = (ch[1:, 0] - ch[:-1, 1]) / 1e3
print("Cholesky chain, %d panels (us): potrf128 mean %.1f | gap to the next panel's potrf128 (TRSM + next-column update + "
      "launches + waits on the side stream) mean %.1f, first five %s, last five %s | factorisation %.0f us"
      % (npan, dur.mean(), gap.mean(), np.round(gap[:5], 1).tolist(), np.round(gap[-5:], 1).tolist(),
         (ch[-1, 1] - ch[0, 0]) / 1e3))
print("  gaps by panel:", " ".join("%.0f" % g for g in gap))
