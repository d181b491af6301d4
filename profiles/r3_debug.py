"""Stage-by-stage comparison of Net.R3 on the device against the oracle's R3 (which reproduces the reference's golden)
on the r3_small stack: prints, per layer and stage, the relative deviation of every intermediate."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch

import cases
import cp_oracle as O
import cpb200
from cpb200.lib import cfgs
from test_gpu_r3 import NumpyConvForward, build_net

eng = cpb200.get_engine()
eng.gram_mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
spec = cases.R3_CASES["r3_small"]
images, specs, weights, biases = cases.r3_inputs(**spec["gen"])
onet = O.NumpyNet(specs, weights, biases)
np.random.seed(spec["np_seed"])
feats, pd = O.extract_features(lambda b: onet.forward_blobs(images[b % len(images)]), onet.convs, spec["nBatches"], spec["P"])
for b in range(spec["nBatches"]):
    pd[(b, 0)] = images[b % len(images)]
onet._feats_dict, onet._points_dict = feats, pd
otrace = []
rng_state = np.random.get_state()
O.R3(onet, state=O.DictState(alpha=1e-3), trace=otrace)

net, _ = build_net(eng, spec, NumpyConvForward)
cfgs.c.nBatches, cfgs.c.nPointsPerLayer = spec["nBatches"], spec["P"]
cfgs.c.dic.vh, cfgs.c.dic.keep = 1, 3.
cfgs.alpha = 1e-3
net.load_frozen(feats_dict=feats, points_dict=pd)
net._trace = []
np.random.set_state(rng_state)
net.R3()
for (c1, st1, a1), (c2, st2, a2) in zip(net._trace, otrace):
    assert (c1, st1) == (c2, st2)
    parts = []
    for k in a2:
        x, y = np.asarray(a1[k], dtype=np.float64), np.asarray(a2[k], dtype=np.float64)
        if k == "idxs":
            parts.append("idxs equal=%s" % np.array_equal(a1[k], a2[k]))
        elif x.shape != y.shape:
            parts.append("%s shape %s vs %s" % (k, x.shape, y.shape))
        else:
            parts.append("%s %.2e" % (k, np.linalg.norm(x - y) / max(1e-300, np.linalg.norm(y))))
    extra = ""
    if "ls" in a1:
        extra = " | ls " + str(a1["ls"])
    if st1 == "vh":
        Xv = a2["X"].reshape(a2["X"].shape[0], -1)
        sv = np.linalg.svd(Xv - Xv.mean(0), compute_uv=False)
        extra += " | sigma_min/max of X %.2e" % (sv[-1] / sv[0])
    print("%-8s %-5s %s%s" % (c1, st1, "  ".join(parts), extra), flush=True)
