"""Drop-in for the pruning-path methods of the reference's ``lib.net.Net`` -- without Caffe.

The reference's ``Net`` wraps a pycaffe handle; the hot path only needs (a) the blobs of a
forward pass, (b) conv hyper-parameters, (c) weights.  Here a ``Net`` is built from a plain
list of conv specs + a weight dict + a *feature provider* ``forward(net, batch) -> {blob:
CUDA tensor (B, C, H, W)}`` (``ConvStackForward`` below runs a sequential conv/ReLU/pool stack
with torch.nn.functional, standing in for Caffe's GPU forward, which is also library code in
the reference).  Methods keep the reference's names, arguments and return conventions:

  extract_features(names, nBatches=None, points_dict=None, save=False)   lib/net.py:368-532
  extract_XY(X, Y, DEBUG=False, w1=None)                                 lib/net.py:534-684
  load_frozen(DEBUG=False, feats_dict=None, points_dict=None)            lib/net.py:839-876
  dictionary_kernel(X_name, weights, d_prime, Y_name, Y, DEBUG=0)        lib/net.py:1685-1735
  R3() -> (WPQ, new_pt)    channel-pruning block only                     lib/net.py:1292-1471

The gathers run on the device (cp_point_gather / cp_patch_gather); sampled points come from the
numpy global RNG with the reference's call sequence, so a seeded run draws the same points.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import cfgs
from .cfgs import c as dcfgs
from .decompose import _dictionary_device, rel_error
from ..engine import get_engine


class ConvSpec:
    """What the reference reads from the prototxt for a Convolution layer (net.py:542-553)."""

    def __init__(self, name, bottom, num_output, kernel_size=3, pad=1, stride=1, pool_after=False):
        self.name, self.bottom = name, bottom
        self.num_output = num_output
        self.kernel_size, self.pad, self.stride = kernel_size, pad, stride
        self.pool_after = pool_after  # a 2x2/2 max-pool follows the ReLU (VGG)


class ConvStackForward:
    """Feature provider for a sequential conv -> ReLU (-> pool) stack.  Blob names follow the
    reference after ``seperateConvReLU`` (net.py:1228): blob ``<conv>`` holds the PRE-ReLU conv
    output; the next conv's bottom is ``<conv>_relu`` or ``pool<k>`` (post-ReLU)."""

    def __init__(self, images_by_batch):
        self.images_by_batch = images_by_batch  # callable batch -> (B,3,H,W) CUDA fp32 tensor

    def __call__(self, net, batch, upto=None):
        x = self.images_by_batch(batch)
        blobs = {"data": x}
        for spec in net._specs:
            w = net._w[spec.name]
            b = net._b[spec.name]
            y = F.conv2d(blobs[spec.bottom], w, b, stride=spec.stride, padding=spec.pad)
            blobs[spec.name] = y
            r = F.relu(y)
            blobs[spec.name + "_relu"] = r
            if spec.pool_after:
                blobs[net._pool_name[spec.name]] = F.max_pool2d(r, 2, 2)
            if upto is not None and spec.name == upto:
                break
        return blobs


class Net:
    def __init__(self, specs, weights, biases, forward, pool_names=None):
        """specs: ordered list of ConvSpec; weights/biases: {name: array (n,c,k,k) / (n,)};
        forward: callable (net, batch) -> {blob name: CUDA tensor (B,C,H,W) fp32}."""
        self.eng = get_engine()
        dev = self.eng.device
        self._specs = list(specs)
        self._spec = {s.name: s for s in self._specs}
        self.convs = [s.name for s in self._specs]
        self._w = {k: torch.as_tensor(np.asarray(v, dtype=np.float32), device=dev).clone() if not isinstance(
            v, torch.Tensor) else v.to(dev, torch.float32).clone() for k, v in weights.items()}
        self._b = {k: torch.as_tensor(np.asarray(v, dtype=np.float32), device=dev).clone() if not isinstance(
            v, torch.Tensor) else v.to(dev, torch.float32).clone() for k, v in biases.items()}
        self._forward = forward
        self._pool_name = pool_names or {}
        self.bottom_names = {s.name: [s.bottom] for s in self._specs}
        self._mem = True
        self.WPQ = {}
        self.selection = {}
        self._feats_dict = None
        self._points_dict = None
        self._feats_dev = {}
        self.num = None

    # ---- accessors with the reference's names (net.py:174-286)
    def param_data(self, name):
        return self._w[name]

    def param_b_data(self, name):
        return self._b[name]

    def param_shape(self, name):
        return tuple(self._w[name].shape)

    def set_param_data(self, name, d):
        self._w[name].copy_(torch.as_tensor(np.asarray(d), device=self._w[name].device))

    def set_param_b(self, name, d):
        self._b[name].copy_(torch.as_tensor(np.asarray(d), device=self._b[name].device))

    def conv_param_pad(self, name):
        return self._spec[name].pad

    def conv_param_kernel_size(self, name):
        return self._spec[name].kernel_size

    def conv_param_stride(self, name):
        return self._spec[name].stride

    def forward(self, batch):
        return self._forward(self, batch)

    # ---- extract_features, net.py:368-532 (conv blobs)
    def extract_features(self, names=[], nBatches=None, points_dict=None, save=False):
        assert nBatches is None, "deprecate"  # net.py:369
        nBatches = dcfgs.nBatches
        nPointsPerLayer = dcfgs.nPointsPerLayer
        if not isinstance(names, list):
            names = [names]
        assert len(names) > 0
        frozen_points = False
        if save:
            if points_dict is None:
                points_dict = dict()
                points_dict["nPointsPerLayer"] = nPointsPerLayer
                points_dict["nBatches"] = nBatches
            else:
                frozen_points = True
                nPointsPerLayer = points_dict["nPointsPerLayer"]
                nBatches = points_dict["nBatches"]
        eng = self.eng
        feats_dev = {}
        P = nPointsPerLayer
        for batch in range(nBatches):
            blobs = self.forward(batch)
            for name in names:
                feat = blobs[name]
                B, n, H, W = feat.shape
                self.num = B
                if name not in feats_dev:
                    feats_dev[name] = eng.empty(nBatches * P * B, n, dtype=torch.float32)
                if save and frozen_points and (batch, name, "randx") in points_dict:
                    randx = points_dict[(batch, name, "randx")]
                    randy = points_dict[(batch, name, "randy")]
                else:
                    randx = np.random.randint(0, H - 0, P)  # net.py:464-465 / 506-507
                    randy = np.random.randint(0, W - 0, P)
                    if save:
                        points_dict[(batch, name, "randx")] = randx.copy()
                        points_dict[(batch, name, "randy")] = randy.copy()
                rx = torch.as_tensor(np.asarray(randx, dtype=np.int32), device=eng.device)
                ry = torch.as_tensor(np.asarray(randy, dtype=np.int32), device=eng.device)
                out = feats_dev[name][batch * P * B:(batch + 1) * P * B]
                eng.point_gather(feat.contiguous(), rx, ry, B, P, out=out)
        self._last_feats_dev = feats_dev
        feats_dict = {k: v.cpu().numpy().astype(np.float64) for k, v in feats_dev.items()}  # fp64, net.py:426
        if save:
            return feats_dict, points_dict
        return feats_dict

    # ---- load_frozen, net.py:839-876 (in-memory branch)
    def load_frozen(self, DEBUG=False, feats_dict=None, points_dict=None):
        assert feats_dict is not None, "only the in-memory branch (net.py:840-844) exists without Caffe/pickles"
        self._feats_dict = feats_dict
        self._points_dict = points_dict
        dev = self.eng.device
        self._feats_dev = {}
        for k, v in feats_dict.items():
            v32 = np.asarray(v, dtype=np.float32)
            self._feats_dev[k] = torch.as_tensor(v32 if np.array_equal(v32.astype(np.float64), v) else np.asarray(v),
                                                 device=dev)

    def freeze(self, names=None):
        """freeze_images (net.py:749-800) without the pickle: sample points + features once."""
        names = names or self.convs
        feats_dict, points_dict = self.extract_features(names, save=1)
        self.load_frozen(feats_dict=feats_dict, points_dict=points_dict)
        return feats_dict, points_dict

    # ---- extract_XY, net.py:534-684
    def _extract_X_device(self, X, Y, relu):
        spec = self._spec[Y]
        pd = self._points_dict
        P, nBatches = pd["nPointsPerLayer"], pd["nBatches"]
        eng = self.eng
        out = None
        for batch in range(nBatches):
            blob = self.forward(batch)[X].contiguous()
            B, c = blob.shape[0], blob.shape[1]
            if out is None:
                out = eng.empty(nBatches * P * B, c * spec.kernel_size ** 2, dtype=torch.float32)
            rx = torch.as_tensor(np.asarray(pd[(batch, Y, "randx")], dtype=np.int32), device=eng.device)
            ry = torch.as_tensor(np.asarray(pd[(batch, Y, "randy")], dtype=np.int32), device=eng.device)
            eng.patch_gather(blob, rx, ry, B, P, spec.kernel_size, spec.pad, spec.stride, relu=relu,
                             out=out[batch * P * B:(batch + 1) * P * B])
        return out

    def extract_XY(self, X, Y, DEBUG=False, w1=None):
        """Returns the (N*k*k, c) float64 matrix of the reference (rows (sample, kh, kw))."""
        assert w1 is None, "the w1 branch (net.py:544-548) is not part of the c3 path"
        k = self._spec[Y].kernel_size
        Xd = self._extract_X_device(X, Y, relu=False)
        N = Xd.shape[0]
        c = Xd.shape[1] // (k * k)
        out = Xd.view(N, c, k * k).permute(0, 2, 1).reshape(N * k * k, c)
        return out.cpu().numpy().astype(np.float64)

    # ---- dictionary_kernel, net.py:1685-1735 (VGG branch: relu on X, resY = 0)
    def dictionary_kernel(self, X_name, weights, d_prime, Y_name, Y, DEBUG=0):
        if dcfgs.model in [cfgs.Models.xception, cfgs.Models.resnet] or dcfgs.res.short:
            raise NotImplementedError("ResNet/Xception residual branches (net.py:1716-1719) are not implemented")
        Xd = self._extract_X_device(X_name, Y_name, relu=True)  # :1698 + :1720
        W2 = self._w[Y_name]
        n, c, h = W2.shape[0], W2.shape[1], W2.shape[-1]
        feats = self._feats_dev[Y_name]
        bias = self._b[Y_name]
        y_bias = bias if feats.dtype == torch.float32 else None
        Yd = feats if y_bias is not None else feats - bias.to(torch.float64)
        idxs, Wd, bd = _dictionary_device(self.eng, Xd, W2.reshape(n, c * h * h), Yd, y_bias, c, h, d_prime)
        rank = int(idxs.sum())
        return idxs, Wd.cpu().numpy().reshape(n, rank, h, h), bd.cpu().numpy()

    # ---- R3, net.py:1292-1471 -- channel-pruning block (:1406-1459)
    def R3(self, alldic=None, pooldic=None, c_ratio=1.15):
        """Walks (conv, convnext) pairs like the reference; for conv in alldic|pooldic prunes
        convnext's input channels.  The VH / ITQ stages of the reference loop (:1351-1404) are 3C
        companions outside this path.  Returns (WPQ, new_pt) with new_pt a dict of pruned widths."""
        convs = self.convs
        self.WPQ = dict()
        self.selection = dict()
        end = 5
        if alldic is None:  # net.py:1307-1308 (VGG-16)
            alldic = ['conv%d_1' % i for i in range(1, end)] + ['conv%d_2' % i for i in range(3, end)]
        if pooldic is None:
            pooldic = ['conv1_2', 'conv2_2']
        new_pt = {}
        for conv, convnext in zip(convs[1:], convs[2:] + ['pool5']):
            if not (dcfgs.dic.vh and (conv in alldic or conv in pooldic) and (convnext in self.convs)):
                continue
            d_c = int(self.param_shape(conv)[0] / c_ratio)  # :1346
            X_name = self.bottom_names[convnext][0] if conv in pooldic else conv  # :1411-1414
            idxs, W2, B2 = self.dictionary_kernel(X_name, None, d_c, convnext, None)
            self.selection[convnext] = idxs
            it = torch.as_tensor(idxs, device=self.eng.device)
            self._w[convnext][:, ~it, ...] = 0  # :1446
            self._w[convnext][:, it, ...] = torch.as_tensor(W2, device=self.eng.device, dtype=torch.float32)
            self.set_param_b(convnext, B2)
            self.WPQ[(conv, 0)] = self._w[conv][it].cpu().numpy()  # producer rows, :1455-1456
            self.WPQ[(conv, 1)] = self._b[conv][it].cpu().numpy()
            self.WPQ[(convnext, 0)] = W2
            self.WPQ[(convnext, 1)] = B2
            new_pt[conv] = int(idxs.sum())
        return self.WPQ, new_pt
