"""GPU: the headline configuration (BASELINE.json configs[1], VGG-16 conv stack, N = 5000 sampled patches) checked
against the oracle AT FULL SIZE on the bench's own problems -- same generator, same seeds as bench.py -- through
the pipeline bench.py times (pruner.prune_layers), in both Gram arithmetic modes.

Gates (north_star): selected-channel set identical, identical alpha-probe sequence (alpha and survivor count of
every Lasso.fit of the search), reconstructed weights within 1e-4 relative Frobenius error, bias within 1e-4.
The oracle runs sklearn's DATA-form coordinate descent (cd_oracle.c:cp_enet_cd_dense, the arithmetic the reference
executes, lib/decompose.py:457) and gelsd least squares (lib/decompose.py:665-666): ~5 s of host time per c = 256
problem, ~20-60 s per c = 512 problem; results are cached per layer so that the second arithmetic mode is free."""
import numpy as np
import pytest

import cp_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

LAYERS = ["conv2_2", "conv3_2", "conv4_1", "conv4_2", "conv5_1"]
_ORACLE = {}


def oracle_layer(s, d):
    """Oracle results for one bench problem (host copies of the device data)."""
    fm = d["fmap"].cpu().numpy()
    randx, randy = d["randx"].cpu().numpy(), d["randy"].cpu().numpy()
    pd = {"nPointsPerLayer": s.P, "nBatches": s.nbatch}
    for b in range(s.nbatch):
        pd[(b, "y", "randx")] = randx[b]
        pd[(b, "y", "randy")] = randy[b]
    forward = lambda b: {"x": fm[b * s.B:(b + 1) * s.B]}  # noqa: E731
    spec = O.ConvSpec("y", "x", s.k, s.pad, s.stride)
    info = {}
    idxs, W, B = O.dictionary_kernel(forward, "x", spec, d["W2"].cpu().numpy(), d["b2"].cpu().numpy(),
                                     d["feats"].cpu().numpy().astype(np.float64), pd, s.rank,
                                     state=O.DictState(alpha=1e-3), samples=d["samples"].cpu().numpy(),
                                     form="dense", info=info, rng=O.SeedFeeder(d["seeds"]))
    return dict(idxs=idxs, W=W, B=B, probes=info["probes"], alpha=info["alpha"])


def compare(res, ref, s):
    assert np.array_equal(res.idxs, ref["idxs"]), "selected-channel set differs (%d vs %d kept)" % (
        res.idxs.sum(), ref["idxs"].sum())
    plog = res.probes.probe_log[:res.nprobe].cpu().numpy()
    got = [(float(a), int(z)) for a, z, _, _ in plog]
    assert got == ref["probes"], "alpha-probe sequence differs: %r vs %r" % (got, ref["probes"])
    assert res.alpha == ref["alpha"]
    W = res.W.cpu().numpy().reshape(ref["W"].shape)
    relW = np.linalg.norm(W - ref["W"]) / np.linalg.norm(ref["W"])
    relB = np.abs(res.b.cpu().numpy() - ref["B"]).max() / max(1.0, np.abs(ref["B"]).max())
    assert relW <= 1e-4 and relB <= 1e-4, (relW, relB)
    return relW, relB


@pytest.mark.parametrize("mode", [1, 0], ids=["3xtf32", "fp64"])
@pytest.mark.parametrize("name", LAYERS)
def test_bench_layer_matches_oracle_at_full_size(engine, name, mode):
    import cpb200
    from cpb200 import pruner

    shapes = cpb200.synth.vgg16_layers()
    i = [s.name for s in shapes].index(name)
    s = shapes[i]
    d = cpb200.synth.make_problem_device(s, 1000 + i, engine)  # bench.py: seed 1000 + problem index
    engine.gram_mode = mode
    res = pruner.prune_layers(engine, [s], [d], right0=1e-3, rank_tol=.1)[0]
    torch.cuda.synchronize()
    if name not in _ORACLE:
        _ORACLE[name] = oracle_layer(s, d)
    relW, relB = compare(res, _ORACLE[name], s)
    assert res.info["verdict"] == "ok", res.info  # iid bench data stays on the fast path
    print("%s mode=%d kept %d/%d probes %d relW %.2e relB %.2e pivot_ratio %.3f"
          % (name, mode, res.idxs.sum(), s.c, res.nprobe, relW, relB, res.info["pivot_ratio"]))
