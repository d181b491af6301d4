"""Maps the error of the least-squares reconstruction computed from tensor-core (3xTF32) Gram statistics against the
conditioning signal the Cholesky reports (smallest pivot / original diagonal), on features of increasing
correlation: the data behind engine.LS_RATIO_MIN (when is the tensor-core Gram accurate enough for the 1e-4 weight
tolerance, when must the layer be re-solved from exact-product fp64 statistics).

    python profiles/conditioning_map.py           (GPU)
Truth = the same solve from fp64 statistics (exact products; error ~ cond * 2e-16), cross-checked against numpy
lstsq for the smaller cases."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cpb200

eng = cpb200.Engine()
dev = eng.device


def features(c, N, k, passes, seed, collinear_eps=None, H=24):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    nimg = -(-N // 16)
    img = torch.randn((nimg, 3, H + 8, H + 8), generator=g, device=dev, dtype=torch.float32)
    for _ in range(passes):
        img = (img + img.roll(1, 2) + img.roll(-1, 2)) / 3
        img = (img + img.roll(1, 3) + img.roll(-1, 3)) / 3
    img = img / img.std()
    w1 = torch.randn((c, 3, 3, 3), generator=g, device=dev) * (2.0 / 27) ** .5
    feat = torch.relu(torch.nn.functional.conv2d(img, w1) + 0.1 * torch.randn((c,), generator=g, device=dev)[None, :, None, None])
    if collinear_eps is not None:
        feat[:, 1] = feat[:, 0] * 1.5 * (1 + collinear_eps * torch.randn(feat[:, 0].shape, generator=g, device=dev))
    Hf = feat.shape[-1]
    r = np.random.RandomState(seed)
    ys, xs, ims = r.randint(0, Hf - k + 1, N), r.randint(0, Hf - k + 1, N), r.randint(0, nimg, N)
    idx = torch.as_tensor(ims, device=dev)
    patches = torch.stack([feat[idx, :, torch.as_tensor(ys + dy, device=dev), torch.as_tensor(xs + dx, device=dev)]
                           for dy in range(k) for dx in range(k)], dim=2)  # (N, c, k*k)
    return patches.reshape(N, c * k * k).contiguous()


print("%-28s %10s %10s %10s %10s" % ("case", "pivot_ratio", "relW(tc)", "relW(f64)", "cond(Gc)"))
for c, n, N in ((32, 32, 2000), (64, 64, 5000), (128, 128, 5000)):
    for label, passes, eps in (("iid-like p=0", 0, None), ("smooth p=1", 1, None), ("smooth p=2", 2, None),
                               ("smooth p=3", 3, None), ("smooth p=3 + pair 1e-2", 3, 1e-2),
                               ("smooth p=3 + pair 1e-3", 3, 1e-3), ("smooth p=3 + pair 1e-4", 3, 1e-4)):
        X = features(c, N, 3, passes, 7 + passes, eps)
        K = X.shape[1]
        g = torch.Generator(device=dev)
        g.manual_seed(99)
        W0 = torch.randn((n, K), generator=g, device=dev, dtype=torch.float64) * (2.0 / K) ** .5
        Y = X.double() @ W0.T
        Y = (Y + 0.01 * Y.std() * torch.randn(Y.shape, generator=g, device=dev, dtype=torch.float64)).float()
        cols = torch.arange(K, dtype=torch.int32, device=dev)
        out = {}
        for mode in (1, 0):
            gg = eng.gram(X, Y, mode=mode)
            W, b, info, stat = eng.ls_solve(gg, cols)
            out[mode] = (W.cpu().numpy(), int(info.cpu()[0]), float(stat.cpu()[0]))
        Xh = X.double().cpu().numpy()
        Xc = Xh - Xh.mean(0)
        sv = np.linalg.svd(Xc, compute_uv=False)
        ref = None
        if K <= 600:
            Yh = Y.double().cpu().numpy()
            ref = np.linalg.lstsq(Xc, Yh - Yh.mean(0), rcond=None)[0].T
        truth = ref if ref is not None else out[0][0]
        rel = lambda a: np.linalg.norm(a - truth) / np.linalg.norm(truth)  # noqa: E731
        print("%-28s %10.2e %10.2e %10.2e %10.2e  info tc/f64 %d/%d  c=%d N=%d%s"
              % (label, out[1][2], rel(out[1][0]), rel(out[0][0]), (sv[0] / sv[-1]) ** 2, out[1][1], out[0][1], c, N,
                 "" if ref is not None else "  (truth = fp64 solve)"), flush=True)
