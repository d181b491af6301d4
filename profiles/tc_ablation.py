"""What bounds gram_tc_kernel?  Times cp_gram (G only, conv4_2 shape) with compile-time ablated builds of the kernel
(`make -C channel-pruning_b200/csrc ablate`): each build drops one or more of the pipeline's roles.
    python profiles/tc_ablation.py            (runs itself once per build in a subprocess)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILDS = [("", "product build"), ("_abl1", "converters idle (TMA + MMA on stale operands)"), ("_abl2", "no MMAs (TMA + converters)"),
          ("_abl5", "MMA only"), ("_abl6", "converters only"), ("_abl7", "barrier protocol + prologue/epilogue only")]
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    import torch
    import cpb200
    eng = cpb200.Engine(gram_mode=1)
    X = torch.rand(5000, 4608, device="cuda")
    for _ in range(3):
        eng.gram(X, None, mode=1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        eng.gram(X, None, mode=1)
    b.record()
    torch.cuda.synchronize()
    print("%-50s cp_gram %.3f ms" % (sys.argv[1], a.elapsed_time(b) / 5))
else:
    for suffix, what in BUILDS:
        env = dict(os.environ, CPB200_LIBRARY=os.path.join(ROOT, "channel-pruning_b200", "libcpb200%s.so" % suffix))
        subprocess.run([sys.executable, os.path.abspath(__file__), what], env=env, check=False)
