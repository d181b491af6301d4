"""Seeded input generators shared by oracle/make_golden.py (which runs the real reference on
them) and the tests (which regenerate the same inputs and compare against the stored
reference outputs).  numpy's legacy RandomState stream is stable across versions."""
import numpy as np


def dictionary_inputs(c, n, N, k, seed, noise=0.01):
    """X (N,c,k,k) fp32 post-ReLU patches, W2 (n,c,k,k) fp32, Y (N,n) float64 holding fp32 values."""
    r = np.random.RandomState(seed)
    X = np.maximum(r.standard_normal((N, c, k, k)).astype(np.float32), 0)
    W2 = (r.standard_normal((n, c, k, k)) * np.sqrt(2.0 / (c * k * k))).astype(np.float32)
    Y = X.reshape(N, -1).astype(np.float64) @ W2.reshape(n, -1).T.astype(np.float64)
    Y = Y + noise * Y.std() * r.standard_normal(Y.shape)
    Y = Y.astype(np.float32).astype(np.float64)
    return X, W2, Y


DICTIONARY_CASES = {
    # name: generator args, target rank, np.random.seed before the call, cfgs.alpha on entry
    "c32": dict(gen=dict(c=32, n=16, N=600, k=3, seed=11), rank=27, np_seed=5, alpha0=1e-3),
    "c64": dict(gen=dict(c=64, n=48, N=1000, k=3, seed=12), rank=55, np_seed=6, alpha0=1e-3),
    "k1": dict(gen=dict(c=96, n=32, N=800, k=1, seed=13), rank=83, np_seed=7, alpha0=1e-3),
    "full": dict(gen=dict(c=3, n=8, N=400, k=3, seed=14), rank=3, np_seed=8, alpha0=1e-3),  # rank == c shortcut
    "under": dict(gen=dict(c=64, n=16, N=400, k=3, seed=15), rank=55, np_seed=9, alpha0=1e-3),  # N-1 < K'
    "carry": dict(gen=dict(c=48, n=24, N=800, k=3, seed=16), rank=41, np_seed=10, alpha0=0.016),  # carried alpha
    "tol2": dict(gen=dict(c=40, n=24, N=800, k=3, seed=17), rank=30, np_seed=11, alpha0=1e-3, rank_tol=.2),
    # rank_tol >= 1 is an ABSOLUTE slack on the channel count (decompose.py:493-494): window [24, 27]
    "tolabs": dict(gen=dict(c=36, n=20, N=700, k=3, seed=18), rank=24, np_seed=12, alpha0=1e-3, rank_tol=3),
}


def net_inputs(B, H, c1, c2, k, pad, stride, nimgbatches, seed):
    """A two-conv network: data (B,3,H,H) -> conv1 (3->c1, 3x3 pad 1) -> ReLU -> conv2 (c1->c2, k, pad, stride)."""
    r = np.random.RandomState(seed)
    images = [r.standard_normal((B, 3, H, H)).astype(np.float32) for _ in range(nimgbatches)]
    specs = [dict(name="conv1", bottom="data", k=3, pad=1, stride=1),
             dict(name="conv2", bottom="conv1_relu", k=k, pad=pad, stride=stride)]
    weights = {"conv1": (r.standard_normal((c1, 3, 3, 3)) * np.sqrt(2.0 / 27)).astype(np.float32),
               "conv2": (r.standard_normal((c2, c1, k, k)) * np.sqrt(2.0 / (c1 * k * k))).astype(np.float32)}
    biases = {"conv1": (0.1 * r.standard_normal(c1)).astype(np.float32),
              "conv2": (0.1 * r.standard_normal(c2)).astype(np.float32)}
    return images, specs, weights, biases


NET_CASES = {
    "k3s1": dict(gen=dict(B=4, H=12, c1=24, c2=12, k=3, pad=1, stride=1, nimgbatches=6, seed=21), nBatches=6, P=10,
                 np_seed=31, xy=("conv1", "conv2"), dictionary_kernel=True),
    "k3s2": dict(gen=dict(B=3, H=11, c1=8, c2=6, k=3, pad=1, stride=2, nimgbatches=3, seed=22), nBatches=3, P=5,
                 np_seed=32, xy=("conv1", "conv2")),
    "k5s1": dict(gen=dict(B=2, H=9, c1=6, c2=5, k=5, pad=2, stride=1, nimgbatches=3, seed=23), nBatches=3, P=4,
                 np_seed=33, xy=("conv1", "conv2")),
    "k1s1": dict(gen=dict(B=3, H=8, c1=10, c2=7, k=1, pad=0, stride=1, nimgbatches=2, seed=24), nBatches=2, P=6,
                 np_seed=34, xy=("conv1", "conv2")),
}
