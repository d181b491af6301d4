"""Synthetic workloads for BASELINE.json's configs (no datasets / checkpoints offline).

A *layer problem* is what one iteration of the reference's pruning loop sees
(lib/net.py:1406-1459): the bottom blob of ``convnext`` for ``nBatches`` batches of ``B``
images, the sampled output points, ``convnext``'s weights/bias and its frozen output features
at those points.  Shapes follow ``temp/vgg.prototxt`` (VGG-16 @224x224): each of the 13 conv
layers has its *input* channels pruned to ``int(c / 1.15)`` (``conv1_1``: c = 3 -> the
``rank == c`` shortcut, lib/decompose.py:487).

Data: pre-ReLU bottom blob ~ N(0,1) fp32 (ReLU is fused into the gather, lib/net.py:1720),
weights ~ N(0, 2/(c k^2)) (MSRA filler, lib/builder.py:385), bias ~ 0.01 N(0,1), frozen output
features = conv(relu(blob)) + bias + 1% noise, rounded to fp32 like a Caffe blob.  Every
batch is an independent draw (as 10 new images per batch are in the reference), so the N =
nBatches*P*B sampled patches are distinct and the least-squares systems are well posed.
"""
from __future__ import annotations

import numpy as np

# (name, c_in, n_out, H=W of the layer's input/output map)  -- temp/vgg.prototxt:53-306
VGG16 = [
    ("conv1_1", 3, 64, 224), ("conv1_2", 64, 64, 224),
    ("conv2_1", 64, 128, 112), ("conv2_2", 128, 128, 112),
    ("conv3_1", 128, 256, 56), ("conv3_2", 256, 256, 56), ("conv3_3", 256, 256, 56),
    ("conv4_1", 256, 512, 28), ("conv4_2", 512, 512, 28), ("conv4_3", 512, 512, 28),
    ("conv5_1", 512, 512, 14), ("conv5_2", 512, 512, 14), ("conv5_3", 512, 512, 14),
]

C_RATIO = 1.15  # lib/net.py:1327

# BASELINE config 4: ResNet-50 bottleneck problems.  (name, c_in, n_out, k, H_in, stride, pad, kept) where `kept`
# is the number of surviving INPUT channels recorded in the reference's released pruned net
# temp/resnet-50-cp.prototxt (branch2a: the block's input selector `<prev>_Filter.num_output`; branch2b /
# branch2c: `num_output` of the producing conv).  kept == c_in means the layer was left alone (LS only).
RESNET50 = [
    ('res2a_branch2a', 64, 64, 1, 56, 1, 0, 35), ('res2a_branch2b', 64, 64, 3, 56, 1, 1, 64),
    ('res2a_branch2c', 64, 256, 1, 56, 1, 0, 55), ('res2b_branch2a', 256, 64, 1, 56, 1, 0, 101),
    ('res2b_branch2b', 64, 64, 3, 56, 1, 1, 51), ('res2b_branch2c', 64, 256, 1, 56, 1, 0, 39),
    ('res2c_branch2a', 256, 64, 1, 56, 1, 0, 97), ('res2c_branch2b', 64, 64, 3, 56, 1, 1, 50),
    ('res2c_branch2c', 64, 256, 1, 56, 1, 0, 37), ('res3a_branch2a', 256, 128, 1, 56, 2, 0, 144),
    ('res3a_branch2b', 128, 128, 3, 28, 1, 1, 128), ('res3a_branch2c', 128, 512, 1, 28, 1, 0, 106),
    ('res3b_branch2a', 512, 128, 1, 28, 1, 0, 205), ('res3b_branch2b', 128, 128, 3, 28, 1, 1, 105),
    ('res3b_branch2c', 128, 512, 1, 28, 1, 0, 72), ('res3c_branch2a', 512, 128, 1, 28, 1, 0, 198),
    ('res3c_branch2b', 128, 128, 3, 28, 1, 1, 105), ('res3c_branch2c', 128, 512, 1, 28, 1, 0, 72),
    ('res3d_branch2a', 512, 128, 1, 28, 1, 0, 288), ('res3d_branch2b', 128, 128, 3, 28, 1, 1, 128),
    ('res3d_branch2c', 128, 512, 1, 28, 1, 0, 110), ('res4a_branch2a', 512, 256, 1, 28, 2, 0, 278),
    ('res4a_branch2b', 256, 256, 3, 14, 1, 1, 256), ('res4a_branch2c', 256, 1024, 1, 14, 1, 0, 225),
    ('res4b_branch2a', 1024, 256, 1, 14, 1, 0, 418), ('res4b_branch2b', 256, 256, 3, 14, 1, 1, 209),
    ('res4b_branch2c', 256, 1024, 1, 14, 1, 0, 147), ('res4c_branch2a', 1024, 256, 1, 14, 1, 0, 407),
    ('res4c_branch2b', 256, 256, 3, 14, 1, 1, 204), ('res4c_branch2c', 256, 1024, 1, 14, 1, 0, 158),
    ('res4d_branch2a', 1024, 256, 1, 14, 1, 0, 423), ('res4d_branch2b', 256, 256, 3, 14, 1, 1, 212),
    ('res4d_branch2c', 256, 1024, 1, 14, 1, 0, 155), ('res4e_branch2a', 1024, 256, 1, 14, 1, 0, 412),
    ('res4e_branch2b', 256, 256, 3, 14, 1, 1, 211), ('res4e_branch2c', 256, 1024, 1, 14, 1, 0, 148),
    ('res4f_branch2a', 1024, 256, 1, 14, 1, 0, 595), ('res4f_branch2b', 256, 256, 3, 14, 1, 1, 256),
    ('res4f_branch2c', 256, 1024, 1, 14, 1, 0, 213), ('res5a_branch2a', 1024, 512, 1, 14, 2, 0, 606),
    ('res5a_branch2b', 512, 512, 3, 7, 1, 1, 512), ('res5a_branch2c', 512, 2048, 1, 7, 1, 0, 433),
    ('res5b_branch2a', 2048, 512, 1, 7, 1, 0, 1222), ('res5b_branch2b', 512, 512, 3, 7, 1, 1, 512),
    ('res5b_branch2c', 512, 2048, 1, 7, 1, 0, 437), ('res5c_branch2a', 2048, 512, 1, 7, 1, 0, 1147),
    ('res5c_branch2b', 512, 512, 3, 7, 1, 1, 512), ('res5c_branch2c', 512, 2048, 1, 7, 1, 0, 440),
]


class LayerShape:
    def __init__(self, name, c, n, H, k=3, pad=1, stride=1, N=5000, B=10, P=10, rank=None):
        self.name, self.c, self.n, self.H, self.W = name, c, n, H, H
        self.k, self.pad, self.stride = k, pad, stride
        self.B, self.P = B, P
        assert N % (B * P) == 0, "N must be a multiple of B*P"
        self.nbatch = N // (B * P)
        self.N = N
        self.rank = int(c / C_RATIO) if rank is None else rank
        if c <= 3:
            self.rank = c
        self.K = c * k * k
        self.S = min(400, N // 20)
        self.Ho = (H + 2 * pad - k) // stride + 1  # output map side

    def cost(self):
        """Rough relative cost (Gram + Cholesky flops) for load balancing across GPUs."""
        kp = self.rank * self.k * self.k
        return self.N * self.K * (self.K + 2 * self.n) + kp ** 3 / 3 + 2.0 * kp * kp * self.n


def vgg16_layers(N=5000, B=10, P=10):
    return [LayerShape(nm, c, n, H, N=N, B=B, P=P) for nm, c, n, H in VGG16]


def resnet50_layers(N=5000, B=10, P=10):
    return [LayerShape(nm, c, n, H, k=k, pad=pad, stride=st, N=N, B=B, P=P, rank=kept)
            for nm, c, n, k, H, st, pad, kept in RESNET50]


def make_problem_numpy(shape: LayerShape, seed: int, noise=0.01):
    """Host (numpy) instance of a layer problem -- used by CPU tests and by the oracle leg.
    Returns dict(fmap (nbatch*B,c,H,W) f32, randx/randy (nbatch,P) i32, W2, b2, feats (N,n) f32,
    samples (S,), X (N,c,k,k) f32 relu'd patches)."""
    r = np.random.RandomState(seed)
    s = shape
    fmap = r.standard_normal((s.nbatch * s.B, s.c, s.H, s.W)).astype(np.float32)
    randx = r.randint(0, s.Ho, (s.nbatch, s.P)).astype(np.int32)
    randy = r.randint(0, s.Ho, (s.nbatch, s.P)).astype(np.int32)
    W2 = (r.standard_normal((s.n, s.c, s.k, s.k)) * np.sqrt(2.0 / (s.c * s.k * s.k))).astype(np.float32)
    b2 = (0.01 * r.standard_normal(s.n)).astype(np.float32)
    X = gather_patches_numpy(fmap, randx, randy, s.B, s.k, s.pad, s.stride, relu=True)
    Y = X.reshape(s.N, -1).astype(np.float64) @ W2.reshape(s.n, -1).T.astype(np.float64) + b2
    Y = Y + noise * Y.std() * r.standard_normal(Y.shape)
    feats = Y.astype(np.float32)
    samples = r.randint(0, s.N, s.S)
    return dict(fmap=fmap, randx=randx, randy=randy, W2=W2, b2=b2, feats=feats, samples=samples, X=X)


def gather_patches_numpy(fmap, randx, randy, B, k, pad, stride, relu):
    """Plain numpy statement of the patch layout (rows (batch, point, image); columns (c,kh,kw))
    used to build synthetic targets.  (The *checked* restatement of the reference's
    extract_XY lives in the oracle directory; tests compare the two.)"""
    nimg, c, H, W = fmap.shape
    nbatch, P = randx.shape
    fp = np.zeros((nimg, c, H + 2 * pad, W + 2 * pad), dtype=fmap.dtype)
    fp[:, :, pad:H + pad, pad:W + pad] = fmap
    out = np.empty((nbatch * P * B, c, k, k), dtype=fmap.dtype)
    for b in range(nbatch):
        imgs = fp[b * B:(b + 1) * B]
        for p in range(P):
            y0, x0 = stride * randx[b, p], stride * randy[b, p]
            out[(b * P + p) * B:(b * P + p + 1) * B] = imgs[:, :, y0:y0 + k, x0:x0 + k]
    if relu:
        np.maximum(out, 0, out=out)
    return out


def fmap_nchw(d):
    """The feature map of a problem dict in the reference's blob order (nimg, c, H, W), whatever its HBM layout."""
    return d["fmap"].permute(0, 3, 1, 2) if d.get("layout", "nchw") == "nhwc" else d["fmap"]


def make_problem_device(shape: LayerShape, seed: int, eng, noise=0.01, pinned_host=False, layout="nchw"):
    """Device instance (torch CUDA generator), sized for BASELINE configs (GBs of feature maps).
    feats are produced with the library's own gather + a torch fp64 matmul: this is data
    generation, outside any timed region.
    layout: how the bottom blob sits in HBM.  'nhwc' (channels last, what a device-side forward provider hands over)
    takes the TMA gather; the VALUES are those of the 'nchw' instance of the same seed.  The pinned host copy
    (fmap_host) always keeps the reference's NCHW blob order."""
    import torch

    s = shape
    dev = eng.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    fmap = torch.randn((s.nbatch * s.B, s.c, s.H, s.W), generator=g, device=dev, dtype=torch.float32)
    r = np.random.RandomState(seed)
    randx = torch.as_tensor(r.randint(0, s.Ho, (s.nbatch, s.P)).astype(np.int32), device=dev)
    randy = torch.as_tensor(r.randint(0, s.Ho, (s.nbatch, s.P)).astype(np.int32), device=dev)
    W2 = torch.randn((s.n, s.c, s.k, s.k), generator=g, device=dev, dtype=torch.float32) * float(
        np.sqrt(2.0 / (s.c * s.k * s.k)))
    b2 = 0.01 * torch.randn((s.n,), generator=g, device=dev, dtype=torch.float32)
    X = eng.patch_gather(fmap, randx, randy, s.B, s.P, s.k, s.pad, s.stride, relu=True)
    Y = X.to(torch.float64) @ W2.reshape(s.n, -1).T.to(torch.float64) + b2.to(torch.float64)
    Y = Y + noise * Y.std() * torch.randn(Y.shape, generator=g, device=dev, dtype=torch.float64)
    feats = Y.to(torch.float32)
    samples = torch.as_tensor(r.randint(0, s.N, s.S).astype(np.int32), device=dev)
    seeds = r.randint(0, 2147483647, size=64)
    out = dict(fmap=fmap, randx=randx, randy=randy, W2=W2, b2=b2, feats=feats, samples=samples, seeds=seeds,
               layout=layout)
    del X, Y
    if pinned_host:
        out["fmap_host"] = torch.empty(fmap.shape, dtype=torch.float32, pin_memory=True)
        out["fmap_host"].copy_(fmap)
    if layout == "nhwc":
        out["fmap"] = fmap.permute(0, 2, 3, 1).contiguous()
        del fmap
    return out
