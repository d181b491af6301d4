"""Device timeline of one 13-layer step with the feature maps resident in HBM: when every layer's channel search and
its reconstruction finish (ms after the step started), sorted by reconstruction end.
    python profiles/step_timeline.py [nhwc|nchw]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cpb200
from cpb200 import pruner

layout = sys.argv[1] if len(sys.argv) > 1 else "nhwc"
eng = cpb200.Engine(nstreams=13)
shapes = cpb200.synth.vgg16_layers()
datas = [cpb200.synth.make_problem_device(s, 1000 + i, eng, pinned_host=False, layout=layout) for i, s in enumerate(shapes)]
for _ in range(3):
    pruner.prune_layers(eng, shapes, datas)
torch.cuda.synchronize()
tr = {}
end = torch.cuda.Event(enable_timing=True)
pruner.prune_layers(eng, shapes, datas, trace=tr)
end.record()
torch.cuda.synchronize()
t0 = tr.pop("_t0")
rows = []
for s in shapes:
    marks = {lab: t0.elapsed_time(e) for lab, e in tr.get(s.name, [])}
    rows.append((marks.get("ls_done", 0.0), s.name, s.c, marks))
print("step: %.2f ms" % t0.elapsed_time(end))
for ls_done, name, c, marks in sorted(rows):
    print("  %-8s c=%4d  select_done %6.2f  ls_done %6.2f  (reconstruction %5.2f ms after its search)" % (
        name, c, marks.get("select_done", float("nan")), ls_done, ls_done - marks.get("select_done", float("nan"))))
