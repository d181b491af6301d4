"""CPU: host-side logic and the C-ABI surface (no compute calls -- there is no GPU here)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
import cp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    import cpb200

    syms = cpb200._cabi.declared_symbols()
    assert len(syms) >= 12
    lib = ctypes.CDLL(cpb200._cabi.LIBRARY)  # fails loudly if the .so is missing
    for s in syms:
        assert hasattr(lib, s), "libcpb200.so does not export %s declared in include/cpb200.h" % s
    ffi, l2 = cpb200._cabi.load()
    assert l2.cp_version() >= 100
    assert isinstance(ffi.string(l2.cp_last_error()), bytes)


def test_header_cites_reference_for_every_entry_point():
    src = open(os.path.join(ROOT, "include", "cpb200.h")).read()
    for fn in ("cp_patch_gather", "cp_point_gather", "cp_gram", "cp_lasso_build", "cp_lasso_select", "cp_ls_solve",
               "cp_ls_solve_dual"):
        head = src[:src.index("int " + fn + "(")]
        comment = head[head.rindex("/*"):]
        assert "lib/" in comment and ".py:" in comment, "no reference file:line cited for " + fn


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "channel-pruning_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            txt = open(os.path.join(dp, f), errors="ignore").read() if f.endswith((".py", ".cu", ".cuh", ".h")) else ""
            if f.endswith(".py"):
                assert "cp_oracle" not in txt and "cd_oracle" not in txt and "ref_shims" not in txt, f
                import re

                assert not re.search(r"^\s*(import|from)\s+\S*oracle", txt, flags=re.M), f
                assert not re.search(r"""["']oracle["'/]""", txt), f  # no path into oracle/ either
                assert "import sklearn" not in txt and "from sklearn" not in txt and "import scipy" not in txt, f
            elif txt:
                for line in txt.splitlines():  # C/CUDA sources may cite the oracle in comments only
                    if "#include" in line:
                        assert "oracle" not in line, f


def test_engine_refuses_to_run_without_cuda():
    import torch

    import cpb200

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        cpb200.Engine()


def test_window_matches_reference_quirks():
    from cpb200.engine import window

    assert window(27, .1) == (27, 27 + 2.7)
    assert window(30, .2) == (30 + 3.0, 30 + 6.0)  # lib/decompose.py:498-501
    assert window(10, 2) == (10, 12)  # rank_tol >= 1 is absolute (:493-494)


def test_lpt_assignment_is_balanced_and_deterministic():
    import cpb200

    shapes = cpb200.synth.vgg16_layers()
    costs = [s.cost() for s in shapes]
    for world in (1, 2, 4, 8):
        owner = cpb200.pruner.assign_layers(costs, world)
        assert owner == cpb200.pruner.assign_layers(costs, world)
        loads = [sum(c for c, o in zip(costs, owner) if o == r) for r in range(world)]
        assert max(loads) <= max(sum(costs) / world * 1.6, max(costs) * 1.0001)
        assert set(owner) <= set(range(world))


def test_pack_unpack_roundtrip():
    import torch

    import cpb200

    pr = cpb200.pruner
    c, n, k2, rank = 40, 12, 9, 34
    size = pr.slot_size(c, n, k2, rank, .1)
    r = np.random.RandomState(0)
    idxs = np.zeros(c, bool)
    idxs[r.choice(c, 36, replace=False)] = True
    W = torch.as_tensor(r.standard_normal((n, 36 * k2)))
    b = torch.as_tensor(r.standard_normal(n))
    buf = torch.zeros(size + 5, dtype=torch.float64)
    pr.pack_result(buf, 5, idxs, W, b, 0.004, 7, c, n, k2)
    out = pr.unpack_result(buf, 5, c, n, k2)
    assert np.array_equal(out["idxs"], idxs) and out["alpha"] == 0.004 and out["nprobe"] == 7
    np.testing.assert_array_equal(out["W"].reshape(n, -1), W.numpy())
    np.testing.assert_array_equal(out["b"], b.numpy())


def test_synth_patch_layout_equals_oracle_extract_XY():
    import cpb200

    s = cpb200.synth.LayerShape("t", 6, 4, 9, k=3, pad=1, stride=1, N=60, B=3, P=5)
    d = cpb200.synth.make_problem_numpy(s, 4)
    pd = {"nPointsPerLayer": s.P, "nBatches": s.nbatch}
    for b in range(s.nbatch):
        pd[(b, "y", "randx")] = d["randx"][b]
        pd[(b, "y", "randy")] = d["randy"][b]
    forward = lambda b: {"x": d["fmap"][b * s.B:(b + 1) * s.B]}  # noqa: E731
    XY = O.extract_XY(forward, "x", O.ConvSpec("y", "x", 3, 1, 1), pd)
    Xo = np.rollaxis(XY.reshape((-1, 3, 3, XY.shape[1])), 3, 1)
    np.testing.assert_array_equal(np.maximum(Xo, 0), d["X"].astype(np.float64))


_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import cpb200
pr = cpb200.pruner
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
shapes = [cpb200.synth.LayerShape("L%%d" %% i, c, n, 8, N=200, B=2, P=5) for i, (c, n) in enumerate([(8, 4), (16, 8), (12, 6), (10, 5), (20, 4)])]
owner = pr.assign_layers([s.cost() for s in shapes], world)
sizes = [pr.slot_size(s.c, s.n, 9, s.rank, .1) for s in shapes]
per_rank = [sum(sizes[i] for i in range(len(shapes)) if owner[i] == r) for r in range(world)]
buf = torch.zeros(max(per_rank), dtype=torch.float64)
off = 0
def fake(i):  # deterministic stand-in for a solved layer
    s = shapes[i]; r = np.random.RandomState(100 + i)
    idxs = np.zeros(s.c, bool); idxs[r.choice(s.c, s.rank, replace=False)] = True
    return idxs, torch.as_tensor(r.standard_normal((s.n, s.rank * 9))), torch.as_tensor(r.standard_normal(s.n))
for i in range(len(shapes)):
    if owner[i] == rank:
        idxs, W, b = fake(i)
        pr.pack_result(buf, off, idxs, W, b, 0.001 * (i + 1), i, shapes[i].c, shapes[i].n, 9)
        off += sizes[i]
allbuf = pr.allgather_results(buf, world)
res = pr.unpack_network(shapes, owner, sizes, allbuf)
for i in range(len(shapes)):
    idxs, W, b = fake(i)
    assert np.array_equal(res[i]["idxs"], idxs) and res[i]["nprobe"] == i
    assert np.array_equal(res[i]["W"].reshape(shapes[i].n, -1), W.numpy())
    assert np.array_equal(res[i]["b"], b.numpy())
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_allgather_reassembles_every_layer(tmp_path):
    """world_size-2 gloo run of the sharding protocol (assignment, static packing, ONE all_gather)."""
    script = tmp_path / "w.py"
    script.write_text(_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_h2d_plan_stages_small_maps_and_reads_large_ones_in_place(monkeypatch):
    """Transfer plan of the host-resident input path: maps that the sampled windows mostly cover (conv5_x: 14x14)
    are DMA'd whole, the others are read in place; explicit policies override."""
    import torch

    import cpb200
    from cpb200 import pruner

    monkeypatch.delenv("CPB200_DMA_MAX_MB", raising=False)
    monkeypatch.delenv("CPB200_DMA_RATIO", raising=False)
    shapes = cpb200.synth.vgg16_layers()

    class FakeMap:  # only .numel() is used by the plan
        def __init__(self, n):
            self._n = n

        def numel(self):
            return self._n

    datas = [dict(fmap_host=FakeMap(s.nbatch * s.B * s.c * s.H * s.W)) for s in shapes]
    plan = pruner.h2d_plan(shapes, datas, True)
    by_name = {s.name: p for s, p in zip(shapes, plan)}
    assert all(by_name[n] == "dma" for n in ("conv5_1", "conv5_2", "conv5_3"))
    assert all(by_name[n] == "zc" for n in ("conv1_2", "conv2_2", "conv3_2", "conv4_2"))
    assert pruner.h2d_plan(shapes, datas, "zc") == ["zc"] * len(shapes)
    assert pruner.h2d_plan(shapes, datas, "copy") == ["dma"] * len(shapes)
    monkeypatch.setenv("CPB200_DMA_MAX_MB", "1")
    assert pruner.h2d_plan(shapes, datas, True) == ["zc"] * len(shapes)


def test_reference_arm_reports_what_it_ran(monkeypatch, capsys):
    """bench.py --impl reference: a step is ONE class problem (round-robin, cost-ascending), ms_per_step is the
    measured wall time of the timed steps (so steps x ms_per_step is what the run really took), and the value
    extrapolates the class timings to the 13-layer stack by multiplicity."""
    import importlib.util
    import json
    import types

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    import cpb200

    cost = {64: .01, 128: .04, 256: .10, 512: .40, 3: .001}
    calls = []

    def fake_seconds(shape, seed):
        calls.append((shape.c, shape.n))
        return cost[shape.c], {"t_gather": 0.1 * cost[shape.c], "t_lasso": 0.4 * cost[shape.c], "t_ls": 0.5 * cost[shape.c]}

    monkeypatch.setattr(bench, "cpu_layer_seconds", fake_seconds)
    monkeypatch.delenv("RANK", raising=False)
    args = types.SimpleNamespace(gpus=1, steps=10, warmup=2, workload="vgg16", layers="")
    bench.run_reference(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    shapes = cpb200.synth.vgg16_layers()
    classes = bench.shape_classes(shapes)
    assert len(calls) == 12 and set(calls) == set((c, n) for c, n, _ in classes)  # 12 steps cover the 8 classes
    want = len(shapes) / sum(cost[c] * len(m) for (c, n, k), m in classes.items())
    assert abs(line["value"] - want) <= 1e-9 * want
    assert line["impl"] == "reference" and line["steps"] == 10 and line["cpu_baseline"]["kind"] == "port"
    assert line["ms_per_step"] * line["steps"] < 5e3  # the fake problems take no time: what was RUN, not 13 layers/step
    st = line["cpu_baseline"]["stack_seconds"]
    assert abs(st["lasso"] + st["ls"] + st["gather"] - len(shapes) / want) <= 1e-9
    assert line["config"] == bench.config_dict(args, shapes, 1)


def test_numa_local_context_restores_affinity():
    """engine.numa_local binds the calling thread to the GPU's NUMA node only for the duration of a pinned allocation;
    without a GPU (or without NUMA information) it must be a no-op, and the affinity must come back either way."""
    import cpb200

    before = os.sched_getaffinity(0)
    with cpb200.engine.numa_local(0) as ctx:
        inside = os.sched_getaffinity(0)
        assert inside <= before and len(inside) >= 1
        assert ctx.cpus is None or (ctx.cpus & before)
    assert os.sched_getaffinity(0) == before


def test_header_documents_the_round2_entry_points():
    src = open(os.path.join(ROOT, "include", "cpb200.h")).read()
    for fn in ("cp_gemm_tc_split", "cp_ls_tensor_cores"):
        head = src[:src.index("int " + fn + "(")]
        comment = head[head.rindex("/*"):]
        assert "lib/decompose.py:" in comment, "no reference file:line cited for " + fn


def test_combineHP_folds_P_into_H_like_the_reference_rule():
    """lib/net.py:1473-1504 on the dictionaries Net.R3 returns: a conv whose channel decomposition kept 3 m >= 2 o is
    merged (W = P.H, b = P_b + P.H_b): the merged layer must reproduce H followed by P exactly; others stay."""
    from cpb200.lib import net as N

    r = np.random.RandomState(5)
    WPQ, layers = {}, []
    for name, (m, o, c, k) in {"conv2_1": (48, 64, 24, 3), "conv3_1": (40, 128, 32, 3)}.items():
        WPQ[name + "_V"] = r.standard_normal((c, 16, k, 1))
        WPQ[(name + "_H", 0)] = r.standard_normal((m, c, 1, k))
        WPQ[(name + "_H", 1)] = r.standard_normal(m)
        WPQ[(name + "_P", 0)] = r.standard_normal((o, m, 1, 1))
        WPQ[(name + "_P", 1)] = r.standard_normal(o)
        layers.append({"V": name + "_V", "H": name + "_H", "P": name + "_P", "rank": m, "num_output": o})
    out, pt = N.combineHP(WPQ, {"prefix": "3C5x", "layers": layers})
    # conv2_1: 3*48 >= 2*64 -> merged; conv3_1: 3*40 < 2*128 -> untouched
    assert ("conv2_1_P", 0) not in out and out[("conv2_1_H", 0)].shape == (64, 24, 1, 3)
    assert ("conv3_1_P", 0) in out and out[("conv3_1_H", 0)].shape == (40, 32, 1, 3)
    x = r.standard_normal(24 * 3)  # one flattened 1 x k patch of the H layer's input
    h = WPQ[("conv2_1_H", 0)].reshape(48, -1) @ x + WPQ[("conv2_1_H", 1)]
    y = WPQ[("conv2_1_P", 0)].reshape(64, -1) @ h + WPQ[("conv2_1_P", 1)]
    y2 = out[("conv2_1_H", 0)].reshape(64, -1) @ x + out[("conv2_1_H", 1)]
    np.testing.assert_allclose(y2, y, rtol=1e-12, atol=1e-12)
    assert pt["layers"][0]["P"] is None and pt["layers"][1]["P"] == "conv3_1_P"
    assert ("conv2_1_P", 0) in WPQ  # the input dictionaries are left alone


def test_computation_matches_the_reference_formula():
    """lib/net.py:1049-1081: VGG-16 conv1_1 (224x224x3 -> 64, 3x3) = 224*224*64*3*9 multiply-accumulates."""
    from cpb200.lib import net as N

    total, per = N.computation([("conv1_1", (1, 3, 224, 224), (64, 3, 3, 3), 1),
                                ("conv1_2", (1, 64, 224, 224), (64, 64, 3, 3), 1),
                                ("down", (1, 64, 112, 112), (128, 64, 3, 3), 2)])
    assert per["conv1_1"] == 224 * 224 * 64 * 3 * 9 and per["conv1_2"] == 224 * 224 * 64 * 64 * 9
    assert per["down"] == 112 * 112 * 128 * 64 * 9 // 4 and total == sum(per.values())


def test_split_fp16_scheme_keeps_22_bits():
    """The operand split of gram_tc2.cu / gemm_tc.cu, restated in numpy: v = x * 2^e with max|v| in [2^9, 2^10),
    hi = fp16(v), lo = fp16(v - hi).  hi + lo must reproduce v to 2^-22 relative (2^-25 absolute below 2^-3, where the
    lo half goes subnormal), and the three products the kernels issue (hi.hi + hi.lo + lo.hi) must reproduce a
    dot product to ~2^-21 of sum|a||b| -- the bound DESIGN.md and the GPU tests quote."""
    r = np.random.RandomState(11)
    x = (r.standard_normal((64, 512)) * np.exp(2.0 * r.standard_normal((64, 1)))).astype(np.float32)
    mx = np.abs(x).max(1, keepdims=True)
    e = 9 - np.floor(np.log2(mx)).astype(np.int64)
    v = (x.astype(np.float64) * np.exp2(e)).astype(np.float32)
    assert np.all(np.abs(v).max(1) >= 2.0 ** 9) and np.all(np.abs(v).max(1) < 2.0 ** 10)
    hi = v.astype(np.float16)
    lo = (v - hi.astype(np.float32)).astype(np.float16)
    rec = hi.astype(np.float64) + lo.astype(np.float64)
    err = np.abs(rec - v.astype(np.float64))
    assert np.all(err <= np.maximum(np.abs(v) * 2.0 ** -22, 2.0 ** -25))
    H, L = hi.astype(np.float64), lo.astype(np.float64)
    exact = (v.astype(np.float64) @ v.astype(np.float64).T)
    three = H @ H.T + H @ L.T + L @ H.T          # what the tensor cores accumulate (here without fp32 rounding)
    bound = np.abs(v.astype(np.float64)) @ np.abs(v.astype(np.float64)).T
    assert np.abs(three - exact).max() <= 2.0 ** -21 * bound.max()
    assert (np.abs(three - exact) / bound).max() <= 2.0 ** -20
