"""Where the end-to-end (host-resident inputs/results) step spends its time beyond the device-resident step.
    python profiles/e2e_breakdown.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cpb200
from cpb200 import pruner

eng = cpb200.Engine(nstreams=13)
shapes = cpb200.synth.vgg16_layers()
datas = [cpb200.synth.make_problem_device(s, 1000 + i, eng, pinned_host=True) for i, s in enumerate(shapes)]
torch.cuda.synchronize()


def run(fh, th, reps=3):
    for _ in range(2):
        pruner.prune_layers(eng, shapes, datas, from_host=fh, to_host=th)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        pruner.prune_layers(eng, shapes, datas, from_host=fh, to_host=th)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for fh, th in [(False, False), (False, True), ("zc", True)]:
    print("from_host=%-5s to_host=%-5s : %7.2f ms/step (wall)" % (fh, th, run(fh, th)))
for cap, ratio in ((300, 0.8),):
    os.environ["CPB200_DMA_MAX_MB"] = str(cap)
    os.environ["CPB200_DMA_RATIO"] = str(ratio)
    print("from_host=auto (DMA maps <= %4d MB: %s) to_host=True : %7.2f ms/step (wall)" % (
        cap, "".join("D" if p == "dma" else "z" for p in pruner.h2d_plan(shapes, datas, True)), run(True, True)))
os.environ["CPB200_DMA_MAX_MB"], os.environ["CPB200_DMA_RATIO"] = "300", "0.8"
for fh in (False, True):
    tr = {}
    pruner.prune_layers(eng, shapes, datas, from_host=fh, to_host=True, trace=tr)
    torch.cuda.synchronize()
    t0 = tr.pop("_t0")
    print("timeline (ms after step start) from_host=%s" % fh)
    for s in shapes:
        print("   %-8s " % s.name + "  ".join("%s %6.1f" % (lab, t0.elapsed_time(e)) for lab, e in tr.get(s.name, [])))
if len(sys.argv) < 2:
    sys.exit(0)

# the gathers alone, from pinned host memory vs from HBM
for s, d in zip(shapes, datas):
    for src in ("fmap", "fmap_host"):
        f = d[src]
        eng.patch_gather(f, d["randx"], d["randy"], s.B, s.P, s.k, s.pad, s.stride, relu=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            eng.patch_gather(f, d["randx"], d["randy"], s.B, s.P, s.k, s.pad, s.stride, relu=True)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print("  %-8s gather from %-9s %7.3f ms  (%.1f MB gathered, map %.0f MB -> %.1f GB/s useful)" %
              (s.name, src, ms, s.N * s.K * 4 / 1e6, f.numel() * 4 / 1e6, s.N * s.K * 4 / ms / 1e6))
