"""Drop-in for the channel-pruning part of the reference's ``lib/decompose.py``.

Same names, argument meaning, return values and implicit state (``cfgs.alpha``, the
numpy global RNG) as the reference; the arithmetic runs on the B200 through libcpb200
(see include/cpb200.h).  numpy in -> numpy out, exactly like the reference; torch CUDA
tensors are accepted too and avoid the host<->device copies.

  relu, rel_error        lib/decompose.py:22-23, 31-32
  dictionary             lib/decompose.py:386-634  (c3 configuration)
  fc_kernel              lib/decompose.py:636-669  (default LinearRegression branch)
  VH_decompose, ITQ_decompose, nonlinear_fc
                         signatures kept (lib/decompose.py:85,163,671); the 3C companions
                         are SURVEY.md 8(f) "next" and raise NotImplementedError here.
"""
from __future__ import annotations

import numpy as np
import torch

from . import cfgs
from .cfgs import c as dcfgs
from ..engine import MAX_PROBES, RAND_R_MAX, get_engine


def relu(x):
    """lib/decompose.py:22-23"""
    if isinstance(x, torch.Tensor):
        return torch.clamp_min(x, 0.)
    return np.maximum(x, 0.)


def rel_error(A, B):
    """lib/decompose.py:31-32"""
    return np.mean((A - B) ** 2) ** .5 / np.mean(A ** 2) ** .5


def _dev_f32(a, eng):
    if isinstance(a, torch.Tensor):
        return a.to(device=eng.device, dtype=torch.float32).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=eng.device)


def _dev_y(Y, eng):
    """Targets are used exactly: fp32 when every value is fp32-representable, else fp64."""
    if isinstance(Y, torch.Tensor):
        if Y.dtype == torch.float32:
            return Y.to(eng.device).contiguous()
        Yd = Y.to(device=eng.device, dtype=torch.float64).contiguous()
        Y32 = Yd.to(torch.float32)
        return Y32 if bool((Y32.to(torch.float64) == Yd).all()) else Yd
    Y = np.ascontiguousarray(Y)
    if Y.dtype == np.float32:
        return torch.as_tensor(Y, device=eng.device)
    Y = Y.astype(np.float64, copy=False)
    Y32 = Y.astype(np.float32)
    if np.array_equal(Y32.astype(np.float64), Y):
        return torch.as_tensor(Y32, device=eng.device)
    return torch.as_tensor(Y, device=eng.device)


class DictionaryInfo:
    """Diagnostics of the last ``dictionary`` call (alpha probes, CD iterations)."""
    last = None


def dictionary(X, W2, Y, alpha=1e-4, rank=None, DEBUG=0, B2=None, rank_tol=.1, verbose=0):
    """LASSO channel selection + least-squares reconstruction, reference
    lib/decompose.py:386-634.

    X: (N, c, h, w)   W2: (n, c, h, w)   Y: (N, n)   rank: channels to keep
    returns (idxs bool[c], newW2 (n, c', h, w) float64, newB2 (n,) float64)
    or, with DEBUG, (newX, newW2, newB2) (decompose.py:629-632).

    Reference behaviour that is kept on purpose:
      * ``rank_tol`` argument ignored, dcfgs.dic.rank_tol used (:393); verbose forced off (:387)
      * rows for the LASSO drawn with replacement from the numpy GLOBAL RNG (:425), one
        further global draw per Lasso.fit for its coordinate order (sklearn _cd_fast)
      * alpha search starts at cfgs.alpha and stores the final alpha back (:491, :627);
        with rank == c the LASSO is skipped and cfgs.alpha becomes the *argument* (:487, :627)
      * square kernels assumed: w = h (:401-402)
    Deviation: the reference's unguarded ``while True`` loops (:502, :516) are capped at
    64 probes; hitting the cap raises RuntimeError instead of spinning forever.
    """
    if dcfgs.autodet or dcfgs.solver != cfgs.solvers.sk or dcfgs.ls != 'linear' or dcfgs.dic.alter or \
            dcfgs.dic.debug or dcfgs.fc_ridge or dcfgs.nonlinear_fc or dcfgs.nofc:
        raise NotImplementedError("only the `train.py -action c3` configuration of dictionary() is implemented")
    eng = get_engine()
    N, c, h = X.shape[0], X.shape[1], X.shape[2]
    w = h
    n = W2.shape[0]
    assert tuple(X.shape) == (N, c, h, w) and tuple(W2.shape) == (n, c, h, w) and tuple(Y.shape) == (N, n)
    Xd = _dev_f32(X, eng).reshape(N, c * h * w)
    W2m = _dev_f32(W2, eng).reshape(n, c * h * w)
    Yd = _dev_y(Y, eng)
    idxs, Wd, bd = _dictionary_device(eng, Xd, W2m, Yd, None, c, h, rank, alpha)
    rank = int(idxs.sum())
    newW2 = Wd.cpu().numpy().reshape((n, rank, h, w))
    newB2 = bd.cpu().numpy()
    if DEBUG:
        Xh = X.cpu().numpy() if isinstance(X, torch.Tensor) else np.asarray(X)
        return Xh[:, idxs, ...], newW2, newB2
    return idxs, newW2, newB2


def _dictionary_device(eng, Xd, W2m, Yd, y_bias, c, h, rank, alpha=1e-4):
    """Body of ``dictionary`` on device buffers: Xd (N, c*h*h) fp32 in (c,kh,kw) column order,
    W2m (n, c*h*h) fp32, Yd (N, n) fp32|fp64 with optional fp32 ``y_bias`` subtracted exactly.
    Returns (idxs numpy bool[c], W (n, K') fp64 device, b (n,) fp64 device)."""
    rank_tol = dcfgs.dic.rank_tol  # :393
    N = Xd.shape[0]
    k2 = h * h
    S = min(400, N // 20)
    samples = np.random.randint(0, N, S)  # :425 -- consumed even when rank == c, like the reference
    info = {"samples": samples, "probes": [], "alpha": alpha}
    if rank == c:  # :487-488
        idxs = np.array([True] * rank)
        g_full = eng.gram(Xd, Yd, y_bias=y_bias)
    else:
        state = np.random.get_state()
        seeds = np.random.randint(0, RAND_R_MAX, size=MAX_PROBES)
        samples_d = torch.as_tensor(samples.astype(np.int32), device=eng.device)
        g_full, res = eng.select_channels_async(Xd, W2m, Yd, y_bias, samples_d, c, k2, rank, rank_tol, cfgs.alpha,
                                                seeds)
        scal = res.scalars.cpu().numpy()  # synchronises
        nprobe, status = int(scal[1]), int(scal[2])
        np.random.set_state(state)
        if nprobe:
            np.random.randint(0, RAND_R_MAX, size=nprobe)  # the draws the reference's fits would have made
        plog = res.probe_log[:nprobe].cpu().numpy()
        info["probes"] = [(float(a), int(z)) for a, z, _, _ in plog]
        info["cd"] = [(int(it), float(gap)) for _, _, it, gap in plog]
        info["coef"] = res.coef.cpu().numpy()
        if status != 0:
            raise RuntimeError("alpha search hit the %d-probe cap (the reference would loop forever); probes=%r"
                               % (MAX_PROBES, info["probes"]))
        alpha = float(scal[0])
        idxs = res.idxs.cpu().numpy().astype(bool)
    Wd, bd = _solve_ls(eng, g_full, Xd, Yd, y_bias, idxs, k2, info)
    cfgs.alpha = alpha  # :626-627
    info["alpha"] = alpha
    DictionaryInfo.last = info
    return idxs, Wd, bd


def _solve_ls(eng, g_full, Xd, Yd, y_bias, idxs, k2, info=None):
    """LS on the surviving channels with the conditioning policy of engine.LS_RATIO_MIN: statistics from the
    tensor-core Gram are used while the Cholesky stays well conditioned, otherwise the layer is re-solved from
    exact-product fp64 statistics; a pivot below sklearn's rank cut-off (cond=1e-6, _base.py:752) raises."""
    Wd, bd, info_d, stat_d = eng.reconstruct_async(g_full, Xd, Yd, y_bias, idxs, k2)
    dual = not (g_full["N"] - 1 >= int(np.count_nonzero(idxs)) * k2)
    fail, ratio = int(info_d.cpu()[0]), float(stat_d.cpu()[0])
    verdict = eng.ls_verdict(fail, ratio, g_full["mode"], dual)
    if info is not None:
        info["ls"] = {"pivot_ratio": ratio, "verdict": verdict}
    if verdict == "redo":
        Wd, bd, info_d, stat_d = eng.reconstruct_exact_async(Xd, Yd, y_bias, idxs, k2)
        fail, ratio = int(info_d.cpu()[0]), float(stat_d.cpu()[0])
        verdict = "singular" if fail else "ok"
        if info is not None:
            info["ls"].update(pivot_ratio_exact=ratio, verdict="redo->" + verdict)
    if verdict == "singular":
        raise np.linalg.LinAlgError(
            "least-squares system is numerically rank deficient (pivot %d below 1e-12 of its diagonal: the "
            "reference's gelsd would truncate here)" % fail)
    return Wd, bd


def fc_kernel(X, Y, copy_X=True, W=None, B=None, ret_reg=False, fit_intercept=True):
    """Least squares with intercept, reference lib/decompose.py:636-669 (default branch:
    ``LinearRegression(fit_intercept=True).fit(X, Y)``).  Returns (coef (n, K), intercept (n,))."""
    assert copy_X == True  # noqa: E712   (decompose.py:640)
    assert len(X.shape) == 2  # decompose.py:641
    if dcfgs.ls != 'linear' or dcfgs.fc_ridge:
        raise NotImplementedError("only the default LinearRegression branch of fc_kernel is implemented")
    if ret_reg or not fit_intercept:
        raise NotImplementedError("ret_reg / fit_intercept=False are used by nonlinear_fc only (SURVEY.md 8f)")
    eng = get_engine()
    Xd = _dev_f32(X, eng)
    Yd = _dev_y(Y, eng)
    g = eng.gram(Xd, Yd)
    K = Xd.shape[1]
    Wd, bd = _solve_ls(eng, g, Xd, Yd, None, np.ones(K, dtype=bool), 1)
    return Wd.cpu().numpy(), bd.cpu().numpy()


def VH_decompose(weights, rank=None, DEBUG=0, X=None, Y=None):
    """Signature of reference lib/decompose.py:85.  Spatial (VH) decomposition is a 3C
    companion outside the pruning hot path."""
    raise NotImplementedError("VH_decompose: 3C companion, not part of the channel-pruning hot path yet")


def ITQ_decompose(feature, gt_feature, weight, rank, bias=None, DEBUG=False, Wr=None):
    """Signature of reference lib/decompose.py:163."""
    raise NotImplementedError("ITQ_decompose: 3C companion, not part of the channel-pruning hot path yet")


def nonlinear_fc(X, Y, copy_X=True, W=None, B=None):
    """Signature of reference lib/decompose.py:671."""
    raise NotImplementedError("nonlinear_fc: 3C companion, not part of the channel-pruning hot path yet")
