"""BASELINE.json configs[4]: patch-count sweep N in {1k, 5k, 20k, 100k} on the conv4_3 shape (c = n = 512, k = 3,
K = 4608) -- Gram GEMM against the tensor roofline, im2col against HBM, the LASSO search (Gram form, latency bound:
ns per coordinate; data form: streamed bytes per second) and the least squares, one JSON line.

    python bench.py --workload sweep            (or python profiles/sweep_config5.py)
"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(torch, fn, reps=4, skip=1, flush=None):
    times = []
    out = None
    for it in range(reps):
        if flush is not None:
            flush.fill_(it)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        if it >= skip:
            times.append(a.elapsed_time(b))
    return statistics.median(times), out  # median: the first repetition after a workspace growth is slow


def main(args=None):
    import numpy as np
    import torch

    import bench
    import cpb200

    torch.cuda.set_device(0)
    eng = cpb200.Engine(device=0)
    dev = eng.device
    peaks, which = bench.measured_peaks()
    lib_peaks = bench.library_peaks(torch, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    points = []
    for N in (1000, 5000, 20000, 100000):
        s = cpb200.synth.LayerShape("conv4_3", 512, 512, 28, N=N)
        d = cpb200.synth.make_problem_device(s, 4300 + N // 1000, eng)
        W2m = d["W2"].reshape(s.n, s.K)
        t_g, X = timed(torch, lambda: eng.patch_gather(d["fmap"], d["randx"], d["randy"], s.B, s.P, s.k, s.pad, s.stride,
                                                        relu=True), flush=flush)
        t_gram, g_full = timed(torch, lambda: eng.gram(X, d["feats"], y_bias=d["b2"], want_sums=False), flush=flush)
        g_full = eng.gram(X, d["feats"], y_bias=d["b2"])
        flops = float(N) * s.K * (s.K + 1) + 2.0 * N * s.K * s.n
        g_s = eng.gram(X, d["feats"], y_bias=d["b2"], rows=d["samples"], want_yy=True, mode=0)
        g_w = eng.gram(W2m, None, want_B=False, mode=0)
        Q, qv, yn2 = eng.lasso_build(g_s, g_w, W2m, s.c, 9, s.S)
        lb, rb = cpb200.engine.window(s.rank, .1)
        t_sel, res = timed(torch, lambda: eng.lasso_select(Q, qv, yn2, float(s.S) * s.n, s.rank, lb, rb, 1e-3, d["seeds"]),
                           reps=2, skip=0)
        scal = res.scalars.cpu().numpy()
        plog = res.probe_log[:int(scal[1])].cpu().numpy()
        sweeps = int(plog[:, 2].sum())
        idxs = res.idxs.cpu().numpy().astype(bool)
        t_ls, out = timed(torch, lambda: eng.reconstruct_async(g_full, X, d["feats"], d["b2"], idxs, 9), reps=2, skip=0)
        fail = int(out[2].cpu()[0])
        pt = {"N": N, "S": s.S, "X_bytes": 4 * N * s.K,
              "gather_ms": t_g, "gather_gbs": 8.0 * N * s.K / (t_g / 1e3) / 1e9,
              "gather_frac_of_hbm": 8.0 * N * s.K / (t_g / 1e3) / 1e9 / peaks["hbm_gbs"],
              "gram_ms": t_gram, "gram_tflops": flops / (t_gram / 1e3) / 1e12,
              "gram_frac_of_tf32": flops / (t_gram / 1e3) / 1e12 / lib_peaks["tf32_tflops"],
              "gram_frac_of_bf16_peak": flops / (t_gram / 1e3) / 1e12 / peaks["bf16_tflops"],  # 3 MMAs per product: ceiling 1/3
              "select_ms": t_sel, "probes": int(scal[1]), "cd_sweeps": sweeps,
              "select_ns_per_coordinate": 1e6 * t_sel / max(1, sweeps * s.c), "kept": int(idxs.sum()),
              "ls_ms": t_ls, "ls_path": "dual (N-1 < K')" if N - 1 < int(idxs.sum()) * 9 else "primal + refinement",
              "ls_info": fail}
        if hasattr(eng, "lasso_cd_dataform"):
            pt["dataform"] = eng.dataform_benchmark(X, W2m, d["feats"], d["b2"], d["samples"], s, float(scal[0]))
        points.append(pt)
        print("N=%6d gather %.3f ms (%.0f GB/s) | gram %.3f ms (%.0f TF/s) | select %.2f ms (%.0f ns/coord) | ls %.2f ms"
              % (N, t_g, pt["gather_gbs"], t_gram, pt["gram_tflops"], t_sel, pt["select_ns_per_coordinate"], t_ls),
              file=sys.stderr, flush=True)
        del d, X, g_full, g_s, g_w
        torch.cuda.empty_cache()
    line = {"metric": "conv4_3_patch_count_sweep", "unit": "per-kernel", "n_gpus": 1, "data": "synthetic",
            "config": {"workload": bench.WORKLOADS["sweep"], "c": 512, "n": 512, "k": 3, "K": 4608},
            "peaks": {"hbm_gbs": peaks["hbm_gbs"], "bf16_tflops": peaks["bf16_tflops"], "hbm_source": which, **lib_peaks}, "points": points}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
