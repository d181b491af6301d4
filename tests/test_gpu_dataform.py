"""GPU: the data-form coordinate-descent benchmark kernel (cp_lasso_dataform_build + cp_lasso_cd_dataform) against the
oracle's restatement of sklearn's data-form solver (cd_oracle.c:cp_enet_cd_dense, screening off) on the same Z, y,
alpha and per-fit seed: identical sweep count and support, coefficients to reduction-order rounding."""
import numpy as np
import pytest

import cases
import cp_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev(a, eng):
    return torch.as_tensor(np.ascontiguousarray(a), device=eng.device)


@pytest.mark.parametrize("c,n,N,S,alpha", [(24, 16, 600, 30, 0.004), (40, 24, 1000, 50, 0.01), (67, 33, 1500, 75, 0.002)])
def test_dataform_cd_matches_oracle_dense(engine, c, n, N, S, alpha):
    X, W2, Y = cases.dictionary_inputs(c=c, n=n, N=N, k=3, seed=200 + c)
    r = np.random.RandomState(c)
    samples = r.randint(0, N, S).astype(np.int32)
    Xd = _dev(X.reshape(N, -1), engine)
    Z, y = engine.lasso_dataform_build(Xd, _dev(W2.reshape(n, -1), engine), _dev(Y.astype(np.float32), engine), None,
                                       _dev(samples, engine), c, 9)
    # the reference's Z (lib/decompose.py:428-434), in float64
    reX = np.rollaxis(X.astype(np.float64).reshape((N, c, -1))[samples], 1, 0)
    reW2 = np.transpose(W2.astype(np.float64).reshape((n, c, -1)), [1, 2, 0])
    Zref = np.matmul(reX, reW2).reshape((c, -1)).T
    Zh = Z.cpu().numpy().T.astype(np.float64)
    assert Zh.shape == Zref.shape and np.abs(Zh - Zref).max() <= 2e-7 * np.abs(Zref).max()  # stored in fp32
    np.testing.assert_array_equal(y.cpu().numpy(), Y[samples].reshape(-1))
    seed = 987654 + c
    solver = O.LassoCD(alpha=alpha, form="dense", rng=O.SeedFeeder([seed, seed]), do_screening=False)
    solver.fit(Zh, y.cpu().numpy())  # the oracle on the SAME (fp32-rounded) design matrix
    w, out = engine.lasso_cd_dataform(Z, y, alpha, seed)
    o = out.cpu().numpy()
    wd = w.cpu().numpy()
    assert int(o[0]) == solver.n_iter_, (o, solver.history)
    assert np.array_equal(wd != 0, solver.coef_ != 0)
    assert np.abs(wd - solver.coef_).max() <= 1e-9 * max(1.0, np.abs(solver.coef_).max())
    # warm start at a second alpha (what the reference's alpha search does)
    solver.alpha = alpha * 2
    solver.fit(Zh, y.cpu().numpy())
    w2, out2 = engine.lasso_cd_dataform(Z, y, alpha * 2, seed, w=w)
    assert int(out2.cpu()[0]) == solver.n_iter_
    assert np.array_equal(w2.cpu().numpy() != 0, solver.coef_ != 0)
