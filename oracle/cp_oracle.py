"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the channel-pruning hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module, and only as the *checker* or
as the timed CPU baseline.  The product (``channel-pruning_b200``) never does.

What it is
----------
A numpy (+ plain C for the coordinate-descent loop, ``oracle/cd_oracle.c``)
restatement of the reference's algorithm, function by function:

  ===========================  ==========================================
  here                         reference (ethanhe42/channel-pruning)
  ===========================  ==========================================
  relu, rel_error              lib/decompose.py:22-23, 31-32
  extract_features             lib/net.py:368-532   (point sampling / gather)
  extract_XY                   lib/net.py:534-684   (sparse-point im2col)
  dictionary_kernel            lib/net.py:1685-1735
  dictionary                   lib/decompose.py:386-634
  fc_kernel                    lib/decompose.py:636-669
  prune_block_R3               lib/net.py:1406-1459 (channel-pruning block of R3)
  solve_relu, svd, pinv        lib/decompose.py:51-59, 154-156, 149-152
  nonlinear_fc                 lib/decompose.py:671-685
  VH_decompose                 lib/decompose.py:85-147
  ITQ_decompose                lib/decompose.py:163-319
  NumpyNet + R3                lib/net.py:1292-1471 (the whole 3C walk: VH -> ITQ -> pruning, sequential)
  ===========================  ==========================================

Third-party arithmetic (absent from /root/reference; the reference pins no
version, README.md:46 -- we pin what is installed in this image):

  scikit-learn 1.9.0  Lasso.fit -> _cd_fast.enet_coordinate_descent  (restated
                      in cd_oracle.c), LinearRegression.fit -> centre + gelsd
  scipy 1.18.1        linalg.lstsq(gelsd) (called through numpy/scipy here)
  numpy 2.3.5         global RandomState draws (samples, CD seeds)

Pinning status
--------------
The reference has no golden vectors or tests for this path.  The
oracle is pinned instead against outputs of the reference's *own code run in the
build container*: ``oracle/make_golden.py`` imports the unmodified
``lib/decompose.py`` / ``lib/net.py`` (``oracle/ref_shims.py``) and writes
``tests/golden/*.npz``; ``tests/test_oracle.py`` checks this module against those
fixtures and against sklearn itself.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RAND_R_MAX = 2147483647


# --------------------------------------------------------------------------- C part
def build_c(force: bool = False) -> str:
    """Compile cd_oracle.c -> oracle/_build/libcporacle.so (gcc, a few ms)."""
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libcporacle.so")
    src = os.path.join(_HERE, "cd_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src, "-lm"])
    return so


def _clib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c())
        dp = ctypes.POINTER(ctypes.c_double)
        lib.cp_enet_cd_dense.restype = ctypes.c_int
        lib.cp_enet_cd_dense.argtypes = [dp, ctypes.c_double, ctypes.c_double, dp, dp, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_uint32,
                                         ctypes.c_int, ctypes.c_int, dp, dp]
        lib.cp_enet_cd_gram.restype = ctypes.c_int
        lib.cp_enet_cd_gram.argtypes = [dp, dp, ctypes.c_double, dp, ctypes.c_int, dp, ctypes.c_double,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_uint32,
                                        ctypes.c_int, ctypes.c_int, dp, dp]
        lib.cp_our_rand_r.restype = ctypes.c_uint32
        lib.cp_our_rand_r.argtypes = [ctypes.POINTER(ctypes.c_uint32)]
        _LIB = lib
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# --------------------------------------------------------------------------- small helpers
def relu(x):
    """lib/decompose.py:22-23"""
    return np.maximum(x, 0.)


def rel_error(A, B):
    """lib/decompose.py:31-32"""
    return np.mean((A - B) ** 2) ** .5 / np.mean(A ** 2) ** .5


# --------------------------------------------------------------------------- sklearn restatement
class LassoCD:
    """``sklearn.linear_model.Lasso(alpha, warm_start=True, selection='random')``
    as the reference constructs it (lib/decompose.py:449), restated.

    fit(): ElasticNet.fit with fit_intercept=True, precompute=False, tol=1e-4,
    max_iter=1000: centre X and y (sklearn _pre_fit/_preprocess_data), keep
    ``coef_`` from the previous fit (warm_start), call enet_coordinate_descent with
    l1_reg = alpha * n_samples (_coordinate_descent.py:781), one
    ``rng.randint(0, RAND_R_MAX)`` consumed from the numpy *global* RandomState
    per fit (_cd_fast.pyx:374; random_state=None -> check_random_state ->
    np.random.mtrand._rand).
    """

    def __init__(self, alpha, tol=1e-4, max_iter=1000, form="dense", rng=None, do_screening=True):
        self.alpha = alpha
        self.tol = tol
        self.max_iter = max_iter
        self.form = form
        self.rng = rng if rng is not None else np.random.mtrand._rand
        self.do_screening = do_screening
        self.coef_ = None
        self._Qw = None  # Gram form: Q @ coef_, carried between fits like the device kernel does
        self._prep = None
        self.history = []  # (alpha, seed, n_iter, gap, nnz)

    def _prepare(self, Z, y):
        # cached: the reference refits the same (Z, reY) for every alpha probe
        if self._prep is not None and self._prep[0] is Z and self._prep[1] is y:
            return self._prep[2:]
        Z64 = np.asarray(Z, dtype=np.float64)
        y64 = np.asarray(y, dtype=np.float64)
        zmean = Z64.mean(axis=0)
        ymean = y64.mean()
        Zc = np.asfortranarray(Z64 - zmean)
        yc = np.ascontiguousarray(y64 - ymean)
        extra = None
        if self.form == "gram":
            Q = np.ascontiguousarray(Zc.T @ Zc)
            q = np.ascontiguousarray(Zc.T @ yc)
            extra = (Q, q, float(yc @ yc))
        self._prep = (Z, y, Zc, yc, zmean, ymean, extra)
        return self._prep[2:]

    def fit(self, Z, y):
        Zc, yc, zmean, ymean, extra = self._prepare(Z, y)
        ns, nf = Zc.shape
        if self.coef_ is None:
            self.coef_ = np.zeros(nf)
        w = np.ascontiguousarray(self.coef_, dtype=np.float64).copy()
        seed = int(self.rng.randint(0, RAND_R_MAX))
        gap = ctypes.c_double()
        tol_s = ctypes.c_double()
        l1_reg = float(self.alpha) * ns
        if self.form == "dense":
            n_iter = _clib().cp_enet_cd_dense(_dp(w), l1_reg, 0.0, _dp(Zc), _dp(yc), ns, nf, self.max_iter,
                                              self.tol, seed, 1, int(self.do_screening),
                                              ctypes.byref(gap), ctypes.byref(tol_s))
        else:
            Q, q, yn2 = extra
            if self._Qw is None:
                self._Qw = np.ascontiguousarray(Q @ w)  # zeros on a cold start
            n_iter = _clib().cp_enet_cd_gram(_dp(w), _dp(self._Qw), l1_reg, _dp(Q), nf, _dp(q), yn2, nf,
                                             self.max_iter, self.tol, seed, 1, int(self.do_screening),
                                             ctypes.byref(gap), ctypes.byref(tol_s))
        self.coef_ = w
        self.intercept_ = ymean - zmean @ w
        self.n_iter_ = n_iter
        self.dual_gap_ = gap.value / ns
        self.history.append((float(self.alpha), seed, n_iter, gap.value, int(np.count_nonzero(w))))
        return self


def linear_regression(X, Y):
    """``LinearRegression(fit_intercept=True).fit(X, Y)`` restated
    (sklearn/linear_model/_base.py): centre X and Y by their column means, solve
    the centred problem with scipy.linalg.lstsq(Xc, Yc, cond=tol) with the estimator's
    default tol=1e-6 (sklearn 1.9.0 _base.py:752-753; LAPACK gelsd: singular values
    below 1e-6*sigma_max are dropped, minimum-norm solution when rank deficient),
    intercept = ybar - xbar @ coef.T.
    Returns (coef_ (n_targets, n_features), intercept_ (n_targets,))."""
    import scipy.linalg

    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    xm = X.mean(axis=0)
    ym = Y.mean(axis=0)
    Xc = X - xm
    Yc = Y - ym
    coef, _, _, _ = scipy.linalg.lstsq(Xc, Yc, cond=1e-6)
    coef = coef.T
    intercept = ym - xm @ coef.T
    return coef, intercept


def fc_kernel(X, Y, copy_X=True, W=None, B=None, ret_reg=False, fit_intercept=True):
    """lib/decompose.py:636-669, default branch (dcfgs.ls='linear', fc_ridge=0)."""
    assert copy_X == True  # noqa: E712  (decompose.py:640)
    assert len(X.shape) == 2  # decompose.py:641
    assert fit_intercept and not ret_reg
    return linear_regression(X, Y)


# --------------------------------------------------------------------------- 3C companions
def solve_relu(RU, Z, Lambda):
    """lib/decompose.py:51-59: argmin_U (relu(U) - Z)^2 + Lambda (U - RU)^2, elementwise."""
    U0 = np.minimum(RU, 0.)  # case 0: U <= 0
    Cost0 = Z ** 2 + Lambda * (U0 - RU) ** 2
    U1 = relu((Lambda * RU + Z) / (Lambda + 1.))  # case 1: U > 0
    Cost1 = (U1 - Z) ** 2 + Lambda * (U1 - RU) ** 2
    return (Cost0 <= Cost1) * U0 + (Cost0 > Cost1) * U1


def svd(x):
    """lib/decompose.py:154-156"""
    import scipy.linalg

    return scipy.linalg.svd(x, full_matrices=False, lapack_driver='gesvd')


def pinv(x):
    """lib/decompose.py:149-152 (scipy.linalg.pinv(x, 1e-6): the positional cond of the reference's scipy is
    today's rtol -- singular values below 1e-6 * sigma_max are dropped)."""
    import scipy.linalg

    return scipy.linalg.pinv(x, rtol=1e-6)


def nonlinear_fc(X, Y, copy_X=True, W=None, B=None):
    """lib/decompose.py:671-685: 30 + 20 alternations of {least squares of U on X, ReLU-aware update of U}."""
    assert len(X.shape) == 2 and copy_X == True and W is None and B is None  # noqa: E712
    U = Y.copy()
    Z = relu(Y)
    its = [30, 20]
    coef = icpt = None
    for epoch, l in enumerate([10 ** i for i in range(-1, 1)]):
        for _ in range(its[epoch]):
            coef, icpt = linear_regression(X, U)  # fc_kernel(X, U, ret_reg=True)
            RU = X @ coef.T + icpt  # reg.predict(X)
            U = solve_relu(RU, Z, l)
    return coef, icpt


def VH_decompose(weights, rank=None, DEBUG=0, X=None, Y=None):
    """lib/decompose.py:85-147: spatial decomposition W (n,c,h,w) ~ V (rank,c,h,1) then H (n,rank,1,w), H refitted
    on data by nonlinear_fc when X, Y are given.  Returns V, H, VHr (n,c,h,w)[, b]."""
    dim = weights.shape
    VH = np.transpose(weights, [1, 2, 0, 3]).reshape([dim[1] * dim[2], dim[0] * dim[3]])  # ch x nw  (:96-99)
    V, sigmaVH, H = svd(VH)
    if rank is None:
        rank = dim[1] * dim[2]
    V = V[:, :rank]
    H = np.diag(sigmaVH[:rank]).dot(H[:rank, :])  # (:105-111)
    VHr = (V.dot(H)).reshape([dim[1], dim[2], dim[0], dim[3]])
    H = np.transpose(H.reshape([rank, dim[0], dim[3], 1]), [1, 0, 3, 2])  # n rank 1 w  (:121-123)
    origV = V.copy()
    V = np.transpose(V.reshape((dim[1], 1, dim[2], rank)), [3, 0, 2, 1])  # rank c h 1  (:125-127)
    b = None
    if X is not None:
        Xv = np.transpose(np.tensordot(X, V, [[1, 2], [1, 2]]), [0, 2, 3, 1])  # (:130-131)
        N = Xv.shape[0]
        o = H.shape[0]
        H, b = nonlinear_fc(Xv.reshape([N, -1]), Y)
        H = H.reshape([o, rank, 1, 3])
        reH = np.transpose(H, [1, 0, 2, 3]).reshape([rank, -1])
        VHr = (origV.dot(reH)).reshape([dim[1], dim[2], dim[0], dim[3]])  # (:134-138)
    VHr = np.transpose(VHr, [2, 0, 1, 3])
    if X is not None:
        return V, H, VHr, b
    return V, H, VHr


def ITQ_decompose(feature, gt_feature, weight, rank, bias=None, DEBUG=False, Wr=None):
    """lib/decompose.py:163-319 on the branch R3 takes (weight (n, r_vh, 1, w) with dim[3] != n, `right = 1`)."""
    n_ins, n_filter_channels = feature.shape
    assert gt_feature.shape == feature.shape
    Y = feature.copy()
    Z = relu(gt_feature)
    Zsq = Z ** 2
    Y_mean = Y.mean(0)
    G = Y - Y_mean
    PG = pinv((G.T).dot(G))  # (:182-189)
    PGGt = PG.dot(G.T)
    UU = G.copy()
    U_mean = Y_mean.copy()
    lambdas = [0.1, 1]
    step_iters = [30, 20]
    T = None
    for step in range(len(lambdas)):
        Lambda = lambdas[step]
        for _ in range(step_iters[step]):
            X = G.dot(PGGt.dot(UU))  # (:213)
            L, sigma, R = svd(X)
            T = L[:, :rank].dot(np.diag(sigma[:rank])).dot(R[:rank, :])  # (:219)
            T = PGGt.dot(T)  # (:225)
            RU = G.dot(T)
            RU += U_mean
            U0 = np.minimum(RU, 0.)
            Cost0 = Zsq + Lambda * (U0 - RU) ** 2
            U1 = relu((Lambda * RU + Z) / (Lambda + 1.))
            Cost1 = (U1 - Z) ** 2 + Lambda * (U1 - RU) ** 2
            U = (Cost0 <= Cost1) * U0 + (Cost0 > Cost1) * U1  # (:231-240)
            U_mean = U.mean(0)
            UU = U - U_mean
    L, sigma, R = svd(T)  # (:250)
    L = L[:, :rank]
    R = np.diag(sigma[:rank]).dot(R[:rank, :])
    dim = weight.shape
    assert len(dim) == 4 and dim[3] != n_filter_channels and dim[0] == n_filter_channels
    weight = np.transpose(weight, [1, 2, 3, 0])
    W1 = weight.reshape([-1, n_filter_channels]).dot(L)  # (:265-266)
    if Wr is not None:
        Wr = np.transpose(Wr, [1, 2, 3, 0])
        W12 = Wr.reshape([-1, n_filter_channels]).dot(L)
    else:
        W12 = W1
    W1 = np.transpose(W1.reshape(weight.shape[:3] + (rank,)), [3, 0, 1, 2])  # (:283-284)
    W2 = R
    W12 = W12.dot(W2)
    W2 = W2.T.reshape([n_filter_channels, rank, 1, 1])
    W12 = np.transpose(W12.reshape((Wr.shape[:3] if Wr is not None else weight.shape[:3]) + (n_filter_channels,)),
                       [3, 0, 1, 2])  # (:301-302)
    B = - Y_mean.dot(T) + U_mean
    B = B.T + bias if bias is not None else B.T
    return W1, W2, B, W12


# --------------------------------------------------------------------------- dictionary
class DictState:
    """The implicit global state the reference keeps in ``lib.cfgs``:
    ``cfgs.alpha`` (cfgs.py:18, mutated at decompose.py:627) and
    ``dcfgs.dic.rank_tol`` (cfgs.py:84)."""

    def __init__(self, alpha=1e-3, rank_tol=.1):
        self.alpha = alpha
        self.rank_tol = rank_tol


class SeedFeeder:
    """Stands in for the numpy global RandomState where a test must hand the CD solver the SAME per-fit seeds the
    device path was given (``rng.randint(0, RAND_R_MAX)`` is the only call Lasso.fit makes, _cd_fast.pyx:374)."""

    def __init__(self, seeds):
        self.seeds = [int(v) for v in seeds]
        self.used = 0

    def randint(self, lo, hi=None, size=None):
        assert size is None
        v = self.seeds[self.used]
        self.used += 1
        return v


def dictionary(X, W2, Y, alpha=1e-4, rank=None, DEBUG=0, B2=None, rank_tol=.1, verbose=0, *, state=None,
               samples=None, form="dense", engine="restated", max_probes=200, info=None, rng=None):
    """lib/decompose.py:386-634 on the default configuration of ``train.py -action c3``
    (dcfgs.autodet=False, solver='sklearn', ls='linear', dic.alter=0, dic.debug=0,
    fc_ridge=0, nonlinear_fc=0, nofc=0).

    X: (N,c,h,w)  W2: (n,c,h,w)  Y: (N,n)  -> idxs bool[c], newW2 (n,c',h,w), newB2 (n,)

    ``state`` carries cfgs.alpha / dic.rank_tol; ``samples`` overrides the draw of
    decompose.py:425 (otherwise taken from the numpy global RNG exactly as there).
    ``form``: 'dense' = sklearn's data-form CD, 'gram' = same control flow in Gram
    arithmetic (model of the CUDA kernel).  ``engine='sklearn'`` calls the installed
    sklearn Lasso instead of the C restatement (cross-check).
    ``max_probes`` guards the reference's unguarded ``while True`` loops
    (decompose.py:502,516); hitting it raises.
    """
    import time as _time

    _t0 = _time.perf_counter()
    state = state if state is not None else DictState()
    rank_tol = state.rank_tol  # decompose.py:393 (argument ignored)
    X = np.asarray(X)
    N, c, h = X.shape[0], X.shape[1], X.shape[2]
    w = h  # decompose.py:401-402
    n = W2.shape[0]
    if samples is None:
        samples = np.random.randint(0, N, min(400, N // 20))  # decompose.py:425
    samples = np.asarray(samples)
    probes = []
    if rank == c:  # decompose.py:487-488
        idxs = np.array([True] * rank)
        tmp = rank
    else:
        reX = np.rollaxis(X.reshape((N, c, -1))[samples], 1, 0)  # c S hw   (:428)
        reW2 = np.transpose(np.asarray(W2).reshape((n, c, -1)), [1, 2, 0])  # c hw n (:430)
        Z = np.matmul(reX, reW2).reshape((c, -1)).T  # (S*n, c)  (:434)
        reY = np.asarray(Y)[samples].reshape(-1)  # (:437)
        if engine == "sklearn":
            from sklearn.linear_model import Lasso

            _solver = Lasso(alpha=alpha, warm_start=True, selection='random')  # (:449)
        else:
            _solver = LassoCD(alpha=alpha, form=form, rng=rng)

        def solve(a):  # decompose.py:453-466
            if len(probes) >= max_probes:
                raise RuntimeError("alpha search did not terminate (reference would loop forever)")
            _solver.alpha = a
            _solver.fit(Z, reY)
            idxs = _solver.coef_ != 0.
            tmp = int(sum(idxs))
            probes.append((float(a), tmp))
            return idxs, tmp

        left = 0
        right = state.alpha  # cfgs.alpha (:491)
        lbound = rank
        if rank_tol >= 1:
            rbound = rank + rank_tol
        else:
            rbound = rank + rank_tol * rank
            if rank_tol == .2:  # (:498-501)
                lbound = rank + 0.1 * rank
                rbound = rank + 0.2 * rank
        while True:  # (:502-515)
            _, tmp = solve(right)
            if tmp < rank:
                break
            else:
                right *= 2
        while True:  # (:516-525)
            alpha = (left + right) / 2
            idxs, tmp = solve(alpha)
            if tmp > rbound:
                left = alpha
            elif tmp < lbound:
                right = alpha
            else:
                break
        rank = tmp  # (:590)
        if info is not None:
            info["coef"] = np.array(_solver.coef_)
            if engine != "sklearn":
                info["cd_history"] = list(_solver.history)
    _t1 = _time.perf_counter()
    # least squares on the survivors (:621-623)
    newW2, newB2 = fc_kernel(X[:, idxs, ...].reshape((N, -1)), Y,
                             W=np.asarray(W2)[:, idxs, ...].reshape(n, -1), B=B2)
    if info is not None:
        info["t_lasso"] = _t1 - _t0
        info["t_ls"] = _time.perf_counter() - _t1
    newW2 = newW2.reshape((n, rank, h, w))
    state.alpha = alpha  # (:626-627); NB: with rank == c this is the *argument default*
    if info is not None:
        info["alpha"] = float(alpha)
        info["probes"] = probes
        info["samples"] = samples
    if DEBUG:
        return X[:, idxs, ...], newW2, newB2
    return idxs, newW2, newB2


# --------------------------------------------------------------------------- sampling / gather
class ConvSpec:
    """pad / kernel_size / stride of a conv layer (what net.py:542-553 reads from
    the prototxt) plus its bottom blob name."""

    def __init__(self, name, bottom, kernel_size=3, pad=1, stride=1):
        self.name, self.bottom = name, bottom
        self.kernel_size, self.pad, self.stride = kernel_size, pad, stride


def extract_features(forward, names, nBatches, nPointsPerLayer, points_dict=None, save=True, rng=None):
    """lib/net.py:368-532 for conv blobs (the ``inner``/FC and ResNet branches are
    out of scope).  ``forward(batch) -> {blob_name: ndarray (B, n, H, W)}`` stands in
    for ``self.forward()`` + ``self.blobs_data(name)``.

    Row order: idx + point*nPicsPerBatch + image  (net.py:509-513,519).
    feats dtype float64 (np.ndarray default, net.py:426).
    Points: randx = randint(0, H, P); randy = randint(0, W, P) (net.py:464-465), one
    pair of draws per (batch, name), shared by all images of the batch.
    """
    rng = rng if rng is not None else np.random
    if not isinstance(names, list):
        names = [names]
    if points_dict is None:
        frozen_points = False
        points_dict = dict()
        points_dict["nPointsPerLayer"] = nPointsPerLayer
        points_dict["nBatches"] = nBatches
    else:
        frozen_points = True
        nPointsPerLayer = points_dict["nPointsPerLayer"]
        nBatches = points_dict["nBatches"]
    feats_dict = dict()
    idx = 0
    nFeatsPerBatch = None
    for batch in range(nBatches):
        blobs = forward(batch)
        for name in names:
            feat = blobs[name]
            num, chs, H, W = feat.shape
            if nFeatsPerBatch is None:
                nFeatsPerBatch = nPointsPerLayer * num
            if name not in feats_dict:
                feats_dict[name] = np.ndarray(shape=(nFeatsPerBatch * nBatches, chs))
            if not frozen_points or (batch, name, "randx") not in points_dict:
                randx = rng.randint(0, H - 0, nPointsPerLayer)
                randy = rng.randint(0, W - 0, nPointsPerLayer)
                points_dict[(batch, name, "randx")] = randx.copy()
                points_dict[(batch, name, "randy")] = randy.copy()
            else:
                randx = points_dict[(batch, name, "randx")]
                randy = points_dict[(batch, name, "randy")]
            for point, x, y in zip(range(nPointsPerLayer), randx, randy):
                i_from = idx + point * num
                feats_dict[name][i_from:(i_from + num)] = feat[:, :, x, y].reshape((num, -1))
        idx += nFeatsPerBatch
    return feats_dict, points_dict


def extract_XY(forward, X_name, Y_spec, points_dict):
    """lib/net.py:534-684 (``w1 is None`` branch).  Returns the (N*k*k, c) float64
    matrix exactly as the reference does; callers reshape to (N,k,k,c) and rollaxis
    to (N,c,k,k) (net.py:1331,1702)."""
    pad, kernel_size, stride = Y_spec.pad, Y_spec.kernel_size, Y_spec.stride
    half = int(kernel_size / 2)
    Y = Y_spec.name
    nPointsPerLayer = points_dict["nPointsPerLayer"]
    nBatches = points_dict["nBatches"]
    feats = None
    idx = 0
    for batch in range(nBatches):
        blob = forward(batch)[X_name]
        B, c, H, W = blob.shape
        nPicsPerBatch = B * kernel_size * kernel_size
        nFeatsPerBatch = nPointsPerLayer * nPicsPerBatch
        if feats is None:
            feats = np.ndarray(shape=(nFeatsPerBatch * nBatches, c))
        feat = np.zeros((B, c, H + 2 * pad, W + 2 * pad), dtype=blob.dtype)  # net.py:631
        feat[:, :, pad:H + pad, pad:W + pad] = blob
        randx = points_dict[(batch, Y, "randx")]
        randy = points_dict[(batch, Y, "randy")]
        for point, x, y in zip(range(nPointsPerLayer), randx, randy):
            i_from = idx + point * nPicsPerBatch
            x0 = half + stride * x  # top2bottom, padded=1 (net.py:564-574)
            y0 = half + stride * y
            xs, xe = x0 - half, x0 + half + 1  # y2x (net.py:580-589)
            ys, ye = y0 - half, y0 + half + 1
            feats[i_from:(i_from + nPicsPerBatch)] = \
                np.moveaxis(feat[:, :, xs:xe, ys:ye], 1, -1).reshape((nPicsPerBatch, -1))  # :656-657
        idx += nFeatsPerBatch
    return feats


def dictionary_kernel(forward, X_name, Y_spec, W2, B2, feats_Y, points_dict, d_prime, *, state=None,
                      samples=None, form="dense", engine="restated", info=None, rng=None):
    """lib/net.py:1685-1735 for the VGG branch (relu on X, resY = 0)."""
    import time as _time

    _t0 = _time.perf_counter()
    X = extract_XY(forward, X_name, Y_spec, points_dict)  # :1698
    if info is not None:
        info["t_gather"] = _time.perf_counter() - _t0
    h = W2.shape[-1]
    w = h
    newX = np.rollaxis(X.reshape((-1, h, w, X.shape[1])), 3, 1).copy()  # :1702
    gtY = feats_Y - B2  # :1707
    Y = gtY
    newX = relu(newX)  # :1720
    if info is not None:
        info["rMSE"] = rel_error(newX.reshape((newX.shape[0], -1)).dot(W2.reshape((W2.shape[0], -1)).T), gtY)
    return dictionary(newX, W2, Y, rank=d_prime, B2=B2, state=state, samples=samples, form=form,
                      engine=engine, info=info, rng=rng)


def prune_block_R3(forward, conv_specs, weights, biases, feats_dict, points_dict, pairs, c_ratio=1.15, *,
                   state=None, samples_by_layer=None, form="dense", infos=None):
    """The channel-pruning block of Net.R3 (lib/net.py:1406-1459) over the given
    (conv, convnext) pairs, on a plain weight dict: for each pair select channels of
    ``conv``'s output / ``convnext``'s input, write back W(convnext) (pruned input
    channels zeroed, :1446-1447), bias, and prune the producer's rows (:1455-1456).
    ``d_c = int(n_out(conv) / c_ratio)`` (net.py:1346).  Mutates copies and returns
    (weights, biases, selection)."""
    weights = {k: np.array(v, dtype=np.float32) for k, v in weights.items()}
    biases = {k: np.array(v, dtype=np.float32) for k, v in biases.items()}
    selection = {}
    WPQ = {}
    state = state if state is not None else DictState()
    for conv, convnext in pairs:
        d_c = int(weights[conv].shape[0] / c_ratio)
        info = {} if infos is not None else None
        samples = None if samples_by_layer is None else samples_by_layer[convnext]
        idxs, W2, B2 = dictionary_kernel(forward, conv, conv_specs[convnext], weights[convnext],
                                         biases[convnext], feats_dict[convnext], points_dict, d_c,
                                         state=state, samples=samples, form=form, info=info)
        selection[convnext] = idxs
        weights[convnext][:, ~idxs, ...] = 0
        weights[convnext][:, idxs, ...] = W2.copy()
        biases[convnext] = B2.astype(np.float32)
        WPQ[(conv, 0)] = weights[conv][idxs]
        WPQ[(conv, 1)] = biases[conv][idxs]
        if infos is not None:
            infos[convnext] = info
    return weights, biases, selection, WPQ


# --------------------------------------------------------------------------- the whole R3 walk on a numpy net
RANKDIC = {'conv1_1': 17, 'conv1_2': 17, 'conv2_1': 37, 'conv2_2': 47, 'conv3_1': 83, 'conv3_2': 89, 'conv3_3': 106,
           'conv4_1': 175, 'conv4_2': 192, 'conv4_3': 227, 'conv5_1': 398, 'conv5_2': 390, 'conv5_3': 379}  # net.py:1309-1321


def conv2d_numpy(x, w, b, pad, stride):
    """fp32 direct convolution standing in for Caffe's forward."""
    B, c, H, W = x.shape
    n, _, k, _ = w.shape
    xp = np.zeros((B, c, H + 2 * pad, W + 2 * pad), dtype=np.float32)
    xp[:, :, pad:H + pad, pad:W + pad] = x
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    out = np.zeros((B, n, Ho, Wo), dtype=np.float32)
    for i in range(Ho):
        for j in range(Wo):
            patch = xp[:, :, i * stride:i * stride + k, j * stride:j * stride + k].reshape(B, -1)
            out[:, :, i, j] = patch @ w.reshape(n, -1).T + b
    return out


class NumpyNet:
    """What the reference's Net is on this path, without Caffe: conv / 2x2 max-pool specs (dicts: name, bottom, k, pad,
    stride | type='pool'), fp32 weights / biases (mutated by R3 like the live Caffe net), frozen images."""

    def __init__(self, specs, weights, biases):
        self.specs = specs
        self.convs = [s["name"] for s in specs if s.get("type") != "pool"]
        self.bottom_names = {s["name"]: [s["bottom"]] for s in specs}
        self.weights = {k: np.array(v, dtype=np.float32) for k, v in weights.items()}
        self.biases = {k: np.array(v, dtype=np.float32) for k, v in biases.items()}
        self.WPQ, self.selection = {}, {}
        self._feats_dict = self._points_dict = None

    def forward_blobs(self, data):
        blobs = {"data": data}
        for s in self.specs:
            if s.get("type") == "pool":
                x = blobs[s["bottom"]]
                B, c, H, W = x.shape
                blobs[s["name"]] = x[:, :, :H // 2 * 2, :W // 2 * 2].reshape(B, c, H // 2, 2, W // 2, 2).max((3, 5))
                continue
            y = conv2d_numpy(blobs[s["bottom"]], self.weights[s["name"]], self.biases[s["name"]], s["pad"], s["stride"])
            blobs[s["name"]] = y
            blobs[s["name"] + "_relu"] = np.maximum(y, 0)
        return blobs

    def spec(self, name):
        s = [q for q in self.specs if q["name"] == name][0]
        return ConvSpec(name, s["bottom"], s["k"], s["pad"], s["stride"])

    def frozen_forward(self):
        pd = self._points_dict
        return lambda batch: self.forward_blobs(pd[(batch, 0)])


def R3(net, keep=3., c_ratio=1.15, state=None, form="dense", infos=None, trace=None, checkpoint=None):
    """lib/net.py:1292-1471 on a NumpyNet whose frozen ``_feats_dict`` / ``_points_dict`` are set (points_dict carries
    the images under (batch, 0) like the reference's, net.py:441).  Mutates net.weights / net.biases; returns WPQ.
    checkpoint(stage): test hook called where the reference enters VH_decompose / ITQ_decompose / dictionary_kernel and
    at the end ('vh', 'itq', 'prune', 'final'); it may inspect and overwrite net.weights / net.biases (make_golden.py
    stores the reference's live state at exactly those points)."""
    state = state if state is not None else DictState()
    checkpoint = checkpoint or (lambda stage: None)
    convs = net.convs
    net.WPQ, net.selection = {}, {}
    end = 5
    alldic = ['conv%d_1' % i for i in range(1, end)] + ['conv%d_2' % i for i in range(3, end)]  # :1307
    pooldic = ['conv1_2', 'conv2_2']
    rankdic = dict(RANKDIC)
    for i in rankdic:
        if 'conv5' in i:
            continue
        rankdic[i] = int(rankdic[i] * 4. / keep)  # :1323-1326

    def getX(name):  # :1329-1331
        x = extract_XY(net.frozen_forward(), net.bottom_names[name][0], net.spec(name), net._points_dict)
        return np.rollaxis(x.reshape((-1, 3, 3, x.shape[1])), 3, 1).copy()

    def setConv(c, d):  # :1333-1337
        if c in net.selection:
            net.weights[c][:, net.selection[c], :, :] = d
        else:
            net.weights[c][...] = d

    for conv, convnext in zip(convs[1:], convs[2:] + ['pool5']):
        conv_V, conv_H, conv_P = conv + '_V', conv + '_H', conv + '_P'
        d_c = int(net.weights[conv].shape[0] / c_ratio)
        rank = rankdic[conv]
        d_prime = rank
        if d_c < rank:
            d_c = rank  # :1349
        # ---- spatial decomposition (:1351-1380)
        checkpoint("vh")
        weights = net.weights[conv]
        if conv in net.selection:
            weights = weights[:, net.selection[conv], :, :]
        Y = net._feats_dict[conv] - net.biases[conv]
        X = getX(conv)
        if conv in net.selection:
            X = X[:, net.selection[conv], :, :]
        V, H, VHr, b = VH_decompose(weights, rank=rank, X=X, Y=Y)
        net.biases[conv][...] = b
        net.WPQ[conv_V] = V
        setConv(conv, VHr)
        net.WPQ[(conv_H, 0)] = H
        net.WPQ[(conv_H, 1)] = net.biases[conv]
        if trace is not None:
            trace.append((conv, "vh", dict(VHr=VHr.copy(), b=np.asarray(b).copy(), X=X.copy(), Y=Y.copy())))
        # ---- channel decomposition (:1384-1404)
        checkpoint("itq")
        feats_new, _ = extract_features(net.frozen_forward(), [conv], None, None, points_dict=net._points_dict)
        W1, W2, B, W12 = ITQ_decompose(feats_new[conv], net._feats_dict[conv], H, d_prime, bias=net.biases[conv], Wr=VHr)
        setConv(conv, W12.copy())
        net.biases[conv][...] = B
        net.WPQ[(conv_H, 0)] = W1.reshape([d_prime, H.shape[1], H.shape[2], H.shape[3]])
        net.WPQ[(conv_H, 1)] = np.zeros(d_prime)
        net.WPQ[(conv_P, 0)] = W2.reshape([W2.shape[0], W2.shape[1], 1, 1])
        net.WPQ[(conv_P, 1)] = B
        if trace is not None:
            trace.append((conv, "itq", dict(W12=W12.copy(), B=np.asarray(B).copy(), Yf=feats_new[conv].copy())))
        # ---- channel pruning (:1406-1459)
        if (conv in alldic or conv in pooldic) and (convnext in net.convs):
            X_name = net.bottom_names[convnext][0] if conv in pooldic else conv
            info = {} if infos is not None else None
            checkpoint("prune")
            idxs, W2n, B2n = dictionary_kernel(net.frozen_forward(), X_name, net.spec(convnext), net.weights[convnext],
                                               net.biases[convnext], net._feats_dict[convnext], net._points_dict, d_c,
                                               state=state, form=form, info=info)
            net.selection[convnext] = idxs
            net.weights[convnext][:, ~idxs, ...] = 0
            net.weights[convnext][:, idxs, ...] = W2n.copy()
            net.biases[convnext][...] = B2n
            key = conv_P if (conv_P, 0) in net.WPQ else conv_H
            net.WPQ[(key, 0)] = net.WPQ[(key, 0)][idxs]
            net.WPQ[(key, 1)] = net.WPQ[(key, 1)][idxs]
            if trace is not None:
                trace.append((conv, "prune", dict(idxs=idxs.copy(), W2=W2n.copy(), B2=np.asarray(B2n).copy())))
            if infos is not None:
                infos[convnext] = info
    checkpoint("final")
    return net.WPQ
