"""-m gpu: seeded random sweep of the argument space of the six entry points (penalty mixes, groups with and
without the unpenalised group 0, zero / uneven penalty factors, user lambda lists, alpha / gamma / tau, the
standardize x intercept flags, ragged n and p, observation weights for xval) -- every draw is run through the C ABI
and compared with the CPU oracle at the BASELINE bar (max |delta beta| <= 1e-8, equal lambda sequences, d to 1e-9)."""
import numpy as np
import pytest

from cases import args_xy, assert_same_fit, binomial_problem, gaussian_problem, sparse_problem

pytestmark = pytest.mark.gpu

COORD = ["lasso", "ols", "elastic.net", "scad", "scad.net", "mcp", "mcp.net"]
GROUP = ["grp.lasso", "grp.lasso.net", "grp.mcp", "grp.scad", "grp.mcp.net", "grp.scad.net", "sparse.grp.lasso"]


def draw(rng, entry):
    n = int(rng.integers(150, 1400))
    p = int(rng.integers(2, 45))
    n = max(n, 6 * p + 20)
    k = int(rng.integers(1, 4))
    pool = COORD + GROUP if entry != "logistic" else [x for x in COORD + GROUP if x != "ols"]
    pens = [str(x) for x in rng.choice(pool, size=k, replace=False)]
    intercept = bool(rng.integers(0, 2))
    standardize = bool(rng.integers(0, 2))
    ngrp = int(rng.integers(1, max(2, p // 2) + 1))
    g = np.sort(rng.integers(1, ngrp + 1, size=p)).astype(np.int32)
    if rng.uniform() < 0.3:
        g[g == g[0]] = 0                                   # an unpenalised group of variables
    explicit = intercept and entry in ("big", "xval", "logistic", "sparse")
    groups = np.concatenate([[0], g]).astype(np.int32) if explicit else g
    pf = rng.uniform(0.3, 2.0, size=p)
    if rng.uniform() < 0.4:
        pf[rng.integers(0, p)] = 0.0
    kw = dict(groups=groups, unique_groups=np.unique(groups), nlambda=int(rng.integers(4, 14)),
              lmin_ratio=float(10 ** rng.uniform(-3, -1)), alpha=float(rng.choice([0.3, 0.6, 1.0])),
              gamma=float(rng.uniform(2.2, 5.0)), tau=float(rng.uniform(0.1, 0.9)), penalty_factor=pf,
              standardize=standardize, intercept=intercept,
              opts=dict(tol=1e-10, maxit=3000, irls_tol=1e-6, irls_maxit=60))
    if rng.uniform() < 0.3:
        gw = rng.uniform(0.5, 2.0, size=np.unique(groups).size)
        gw[np.unique(groups) == 0] = 0.0
        kw["group_weights"] = gw
    return n, p, pens, kw


def user_lambda(rng, ref, kw):
    """re-run with an explicit (unsorted lengths allowed) lambda list per penalty, derived from the generated one"""
    lams = []
    for lam in ref["lambda_"]:
        lam = np.asarray(lam)
        keep = np.sort(rng.choice(lam.size, size=max(1, lam.size // 2), replace=False))
        lams.append(lam[keep] * 0.97)
    kw = dict(kw)
    kw["lambda_"] = lams
    return kw


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_dense(lib, oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    n, p, pens, kw = draw(rng, "dense")
    X, y = gaussian_problem(2000 + seed, n, p, sd_x=float(rng.uniform(0.5, 3.0)), mean_x=float(rng.uniform(-1, 1)))
    kw["compute_loss"] = bool(rng.integers(0, 2))
    a = args_xy(X, y, "gaussian", pens, **kw)
    got, ref = lib.oem_fit_dense(*a), oracle.oem_fit_dense(*a)
    assert_same_fit(got, ref, check_niter=False)
    if kw["compute_loss"]:
        for pp in range(len(pens)):
            assert np.allclose(got["loss"][pp], ref["loss"][pp], rtol=1e-8)
    a = args_xy(X, y, "gaussian", pens, **user_lambda(rng, ref, kw))
    assert_same_fit(lib.oem_fit_dense(*a), oracle.oem_fit_dense(*a), check_niter=False)


@pytest.mark.parametrize("seed", range(30))
def test_fuzz_big(lib, oracle, seed):
    rng = np.random.default_rng(3000 + seed)
    n, p, pens, kw = draw(rng, "big")
    X, y = gaussian_problem(4000 + seed, n, p, mean_x=float(rng.uniform(-0.5, 0.5)))
    a = args_xy(X, y, "gaussian", pens, **kw)
    got, ref = lib.oem_fit_big(*a), oracle.oem_fit_big(*a)
    assert_same_fit(got, ref, check_niter=False)
    a = args_xy(X, y, "gaussian", pens, **user_lambda(rng, ref, kw))
    assert_same_fit(lib.oem_fit_big(*a), oracle.oem_fit_big(*a), check_niter=False)


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_sparse(lib, oracle, seed):
    rng = np.random.default_rng(13000 + seed)
    n, p, pens, kw = draw(rng, "sparse")
    n = max(n, 12 * p + 50)
    X, y = sparse_problem(13500 + seed, n, p, density=float(rng.choice([0.02, 0.1, 0.4])), shift_y=float(rng.uniform(-2, 2)),
                          empty_cols=(int(rng.integers(0, p)),) if rng.uniform() < 0.3 else ())
    kw["compute_loss"] = bool(rng.integers(0, 2))
    a = args_xy(X, y, "gaussian", pens, **kw)
    got, ref = lib.oem_fit_sparse(*a), oracle.oem_fit_sparse(*a)
    assert_same_fit(got, ref, check_niter=False)
    if kw["compute_loss"]:
        for pp in range(len(pens)):
            assert np.allclose(got["loss"][pp][:len(ref["loss"][pp])], ref["loss"][pp], rtol=1e-8)
    a = args_xy(X, y, "gaussian", pens, **user_lambda(rng, ref, kw))
    assert_same_fit(lib.oem_fit_sparse(*a), oracle.oem_fit_sparse(*a), check_niter=False)


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_logistic(lib, oracle, seed):
    rng = np.random.default_rng(5000 + seed)
    n, p, pens, kw = draw(rng, "logistic")
    n = max(n, 400)
    X, y = binomial_problem(6000 + seed, n, p)
    kw["lmin_ratio"] = max(kw["lmin_ratio"], 1e-2)
    kw["opts"]["hessian_type"] = str(rng.choice(["upper.bound", "full"]))
    a = args_xy(X, y, "binomial", pens, **kw)
    got, ref = lib.oem_fit_logistic_dense(*a), oracle.oem_fit_logistic_dense(*a)
    assert_same_fit(got, ref, tol=1e-7, check_niter=False)


@pytest.mark.parametrize("seed", range(30))
def test_fuzz_xval(lib, oracle, seed):
    rng = np.random.default_rng(7000 + seed)
    n, p, pens, kw = draw(rng, "xval")
    F = int(rng.integers(3, 8))
    n = max(n, (F + 1) * (2 * p + 10))
    X, y = gaussian_problem(8000 + seed, n, p, noise=float(rng.uniform(0.5, 3.0)))
    foldid = (1 + rng.permutation(n) % F).astype(np.int32)
    a = args_xy(X, y, "gaussian", pens, **kw)
    if rng.uniform() < 0.5:
        a[4] = rng.uniform(0.2, 3.0, size=n)
    measure = str(rng.choice(["mse", "mae"]))
    a = a[:17] + [F, foldid, False, measure, a[18]]
    got, ref = lib.oem_xval_dense(*a), oracle.oem_xval_dense(*a)
    assert_same_fit(got, ref, check_niter=False)
    for pp in range(len(pens)):
        assert np.allclose(got["cvm"][pp], ref["cvm"][pp], rtol=1e-8, atol=1e-12)
        assert np.allclose(got["cvsd"][pp], ref["cvsd"][pp], rtol=1e-7, atol=1e-12)


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_xtx(lib, oracle, seed):
    rng = np.random.default_rng(9000 + seed)
    n, p, pens, kw = draw(rng, "xtx")
    X, y = gaussian_problem(9500 + seed, n, p, sd_x=float(rng.uniform(0.5, 2.0)))
    xtx, xty = X.T @ X / n, X.T @ y / n
    sf = np.sqrt(np.diag(xtx)) if rng.uniform() < 0.5 else []
    args = ["gaussian", pens, kw["groups"], kw["unique_groups"], kw.get("group_weights", []), [], kw["nlambda"],
            kw["lmin_ratio"], kw["alpha"], kw["gamma"], kw["tau"], sf, kw["penalty_factor"], dict(maxit=3000, tol=1e-10)]
    got, ref = lib.oem_xtx(xtx, xty, *args), oracle.oem_xtx(xtx, xty, *args)
    assert_same_fit(got, ref, check_niter=False, lam_ulps=4)


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_wide_path_modes(lib, oracle, seed):
    """q = 100..900: the path kernel's cluster (DSMEM) and cooperative (global exchange) modes, fast (fused prox) and
    replicated (group penalty) routes mixed in one launch, with and without a scale factor."""
    rng = np.random.default_rng(11000 + seed)
    p = int(rng.integers(100, 900))
    n = 2 * p + 50
    X, y = gaussian_problem(11500 + seed, n, p, nnz=12)
    xtx, xty = X.T @ X / n, X.T @ y / n
    k = int(rng.integers(1, 5))
    pens = [str(x) for x in rng.choice(COORD + GROUP, size=k, replace=False)]
    g = np.sort(rng.integers(1, p // 4 + 1, size=p)).astype(np.int32)
    sf = np.sqrt(np.diag(xtx)) if seed % 2 else []
    args = ["gaussian", pens, g, np.unique(g), [], [], int(rng.integers(4, 9)), 0.05, float(rng.choice([0.5, 1.0])),
            float(rng.uniform(2.5, 4.0)), 0.5, sf, np.ones(p), dict(maxit=400, tol=1e-9)]
    got, ref = lib.oem_xtx(xtx, xty, *args), oracle.oem_xtx(xtx, xty, *args)
    assert_same_fit(got, ref, check_niter=False, lam_ulps=4)


@pytest.mark.parametrize("seed", range(4))
def test_fuzz_xval_wide(lib, oracle, seed):
    """several Grams (folds + full data) of q = 120..400 sharing one persistent launch"""
    rng = np.random.default_rng(12000 + seed)
    p = int(rng.integers(120, 400))
    F = int(rng.integers(3, 6))
    n = (F + 1) * (p + 40)
    X, y = gaussian_problem(12500 + seed, n, p, nnz=10, noise=2.0)
    foldid = (1 + rng.permutation(n) % F).astype(np.int32)
    pens = [str(x) for x in rng.choice(["lasso", "mcp", "scad", "grp.lasso", "elastic.net"], size=2, replace=False)]
    g = np.concatenate([[0], np.sort(rng.integers(1, p // 5 + 1, size=p))]).astype(np.int32)
    a = args_xy(X, y, "gaussian", pens, groups=g, unique_groups=np.unique(g), nlambda=6, lmin_ratio=0.05, alpha=0.8,
                opts=dict(tol=1e-9, maxit=400))
    a = a[:17] + [F, foldid, False, "mse", a[18]]
    got, ref = lib.oem_xval_dense(*a), oracle.oem_xval_dense(*a)
    assert_same_fit(got, ref, check_niter=False)
    for pp in range(2):
        assert np.allclose(got["cvm"][pp], ref["cvm"][pp], rtol=1e-8)
        assert np.allclose(got["cvsd"][pp], ref["cvsd"][pp], rtol=1e-7)
