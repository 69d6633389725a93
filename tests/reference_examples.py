"""The reference's own seeded documentation examples, re-run through a front-end namespace `fe` (oem_b200.frontend, with
either the CUDA library or -- CPU tests only -- the oracle behind it) and compared with the numbers the reference
printed when that documentation was rendered (tests/golden/reference_printed.json, made by
tools/extract_reference_outputs.py from docs/reference/*.html and vignettes/oem_vignette.html).

Inputs come from oracle/r_rng.py (R's set.seed / runif / rnorm stream restated), so nothing here reads /root/reference.
Each function returns {key: (got, printed_values, unit_of_last_printed_digit)}.
"""
import json
import os

import numpy as np

from oracle.r_rng import RStream

_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_printed.json")


def printed():
    with open(_GOLD) as f:
        return json.load(f)


def _pair(got, entry):
    return np.asarray(got, dtype=np.float64).ravel(), np.asarray(entry["values"]), float(entry["unit"])


def _predict_inputs():
    """man/predict.oem.Rd:43-55 (identical in predict.cv.oem.Rd / predict.xval.oem.Rd)."""
    r = RStream(123)
    n_obs, n_vars, n_test = 10000, 100, 1000
    true_beta = np.concatenate([r.runif(15, -0.5, 0.5), np.zeros(n_vars - 15)])
    x = r.matrix_rnorm(n_obs, n_vars)
    y = r.rnorm(n_obs, sd=3) + x @ true_beta
    x_test = r.matrix_rnorm(n_test, n_vars)
    y_test = r.rnorm(n_test, sd=3) + x_test @ true_beta
    return x, y, x_test, y_test


def _mse(y_test, preds):
    return np.mean((y_test[:, None] - preds) ** 2, axis=0)       # apply(preds, 2, function(x) mean((y.test - x) ^ 2))


def example_predict_oem(fe):
    """man/predict.oem.Rd:43-69 -> docs/reference/predict.oem.html; also the two cv.oem lines of predict.cv.oem.html,
    whose full-data fit is this oem() fit (they must be columns of the same path)."""
    g = printed()
    x, y, x_test, y_test = _predict_inputs()
    fit = fe.oem(x, y, penalty=["lasso", "grp.lasso"], groups=np.repeat(np.arange(1, 11), 10), nlambda=10)
    m1 = _mse(y_test, fe.predict(fit, newx=x_test, type="response", which_model=1))
    m2 = _mse(y_test, fe.predict(fit, newx=x_test, type="response", which_model=2))
    out = {"mse_lasso": _pair(m1, g["predict_oem"]["mse_lasso"]),
           "mse_grp_lasso": _pair(m2, g["predict_oem"]["mse_grp_lasso"])}
    cv = g["predict_cv_oem"]
    j = int(np.argmin(np.abs(m2 - cv["mse_grp_lasso"]["values"][0])))       # cv.oem's lambda.min (its folds came from sample())
    out["cv_oem_mse_grp_lasso"] = _pair(m2[j], cv["mse_grp_lasso"])
    out["cv_oem_mse_best"] = _pair(m2[j], cv["mse_best"])
    out["cv_oem_mse_lasso"] = _pair(m1[j], cv["mse_lasso"])
    return out


def example_predict_xval_oem(fe):
    """man/predict.xval.oem.Rd -> docs/reference/predict.xval.oem.html.  The reference drew its folds with sample(), which
    is not restated; the printed numbers only depend on the full-data fit and on WHICH lambda won, so any balanced fold
    assignment that selects the same lambda.min reproduces them."""
    g = printed()["predict_xval_oem"]
    x, y, x_test, y_test = _predict_inputs()
    foldid = 1 + (np.arange(x.shape[0]) % 10)
    fit = fe.xval_oem(x, y, penalty=["lasso", "grp.lasso"], groups=np.repeat(np.arange(1, 11), 10), nlambda=10, foldid=foldid)
    best = _mse(y_test, fe.predict(fit, newx=x_test, type="response", which_model="best.model"))    # s = lambda.min
    gl = _mse(y_test, fe.predict(fit, newx=x_test, type="response", which_model="grp.lasso"))
    la = _mse(y_test, fe.predict(fit, newx=x_test, type="response", which_model=1))
    return {"mse_best": _pair(best, g["mse_best"]), "mse_grp_lasso": _pair(gl, g["mse_grp_lasso"]),
            "mse_lasso": _pair(la, g["mse_lasso"])}


def example_predict_cv_oem(fe):
    """man/predict.cv.oem.Rd -> docs/reference/predict.cv.oem.html, the literal example: cv.oem() (ten refits + held-out
    predictions) and predict(fit, newx, which.model = "best.model" | "grp.lasso" | 1) at lambda.min.  The reference's
    folds came from sample(); any balanced assignment that selects the same lambda.min reproduces the printed values."""
    g = printed()["predict_cv_oem"]
    x, y, x_test, y_test = _predict_inputs()
    foldid = 1 + (np.arange(x.shape[0]) % 10)
    fit = fe.cv_oem(x, y, penalty=["lasso", "grp.lasso"], groups=np.repeat(np.arange(1, 11), 10), nlambda=10, foldid=foldid)
    best = _mse(y_test, fe.predict(fit, newx=x_test, type="response", which_model="best.model"))
    gl = _mse(y_test, fe.predict(fit, newx=x_test, type="response", which_model="grp.lasso"))
    la = _mse(y_test, fe.predict(fit, newx=x_test, type="response", which_model=1))
    return {"mse_best": _pair(best, g["mse_best"]), "mse_grp_lasso": _pair(gl, g["mse_grp_lasso"]),
            "mse_lasso": _pair(la, g["mse_lasso"])}, fit


def example_logLik(fe):
    """man/logLik.Rd:32-55 -> docs/reference/logLik.html: compute.loss = TRUE paths of oem() (100 and, through cv.oem's
    full-data fit, 25 lambdas) and xval.oem() (25 lambdas), lasso and mcp."""
    g = printed()["logLik"]
    r = RStream(123)
    n_obs, n_vars = 2000, 50
    true_beta = np.concatenate([r.runif(15, -0.25, 0.25), np.zeros(n_vars - 15)])
    x = r.matrix_rnorm(n_obs, n_vars)
    y = r.rnorm(n_obs, sd=3) + x @ true_beta
    out = {}
    fit = fe.oem(x, y, penalty=["lasso", "mcp"], compute_loss=True)
    out["oem_lasso"] = _pair(fe.logLik(fit), g["oem_lasso"])
    out["oem_mcp"] = _pair(fe.logLik(fit, which_model="mcp"), g["oem_mcp"])
    fit = fe.oem(x, y, penalty=["lasso", "mcp"], compute_loss=True, nlambda=25)        # cv.oem(...)$oem.fit
    out["cv_oem_lasso"] = _pair(fe.logLik(fit), g["cv_oem_lasso"])
    out["cv_oem_mcp"] = _pair(fe.logLik(fit, which_model="mcp"), g["cv_oem_mcp"])
    fit = fe.xval_oem(x, y, penalty=["lasso", "mcp"], compute_loss=True, nlambda=25, foldid=1 + (np.arange(n_obs) % 10))
    out["xval_oem_lasso"] = _pair(fe.logLik(fit), g["xval_oem_lasso"])
    out["xval_oem_mcp"] = _pair(fe.logLik(fit, which_model="mcp"), g["xval_oem_mcp"])
    return out


XTX_PENALTIES = ["lasso", "elastic.net", "ols", "mcp", "scad", "mcp.net", "scad.net", "grp.lasso", "grp.lasso.net",
                 "grp.mcp", "grp.scad", "sparse.grp.lasso"]


def example_oem_xtx(fe):
    """man/oem.xtx.Rd:113-145 -> docs/reference/oem.xtx.html: oem(standardize = FALSE, intercept = FALSE) against
    oem.xtx() on crossprod(x) / n, all twelve penalties in one call.  The reference printed 8.788848e-16 (rounding noise of
    its BLAS); what is checkable is that the two entry points agree to that order.  Returns max |diff| per penalty."""
    r = RStream(123)
    n_obs, n_vars = 10000, 100
    true_beta = np.concatenate([r.runif(15, -0.25, 0.25), np.zeros(n_vars - 15)])
    x = r.matrix_rnorm(n_obs, n_vars)
    y = r.rnorm(n_obs, sd=3) + x @ true_beta
    groups = np.repeat(np.arange(1, 21), 5)
    fit = fe.oem(x, y, penalty=XTX_PENALTIES, standardize=False, intercept=False, groups=groups)
    fit_xtx = fe.oem_xtx(x.T @ x / n_obs, x.T @ y / n_obs, penalty=XTX_PENALTIES, groups=groups)
    diffs = []
    for pen in XTX_PENALTIES:
        a, b = np.asarray(fit["beta"][pen]), np.asarray(fit_xtx["beta"][pen])
        a = a[1:] if a.ndim == 2 else a[1:]
        diffs.append(float(np.max(np.abs(a.reshape(b.shape) - b))))
    return np.array(diffs)


def example_vignette_bigmat(fe):
    """vignettes/oem_vignette.Rmd:398-428: big.oem() on a file-backed matrix against oem() on the same data in memory;
    the two entry points standardise differently, which is what the printed 1.534783e-05 is."""
    g = printed()["vignette_bigmat"]
    r = RStream(123)
    nrows, ncols = 50000, 100
    bigmat = np.empty((nrows, ncols), order="F")
    for i in range(ncols):
        bigmat[:, i] = r.rnorm(nrows) * (i + 1)
    yb = r.rnorm(nrows) + bigmat[:, 0] - bigmat[:, 1]
    groups = np.repeat(np.arange(1, 21), 5)
    fit = fe.big_oem(bigmat, yb, penalty=["lasso", "grp.lasso"], groups=groups)
    fit2 = fe.oem(bigmat, yb, penalty=["lasso", "grp.lasso"], groups=groups)
    d = np.max(np.abs(np.asarray(fit["beta"]["lasso"]) - np.asarray(fit2["beta"]["lasso"])))
    return {"maxdiff_big_vs_oem_lasso": _pair(d, g["maxdiff_big_vs_oem_lasso"])}


def example_logistic_dense_vs_sparse(fe):
    """man/oem.Rd:186-209 -> docs/reference/oem.html: the ONE number the reference prints for its binomial entries,
    max |oem(dense, family = "binomial") - oem(sparse, family = "binomial")| = 6.647085e-05 (intercept = FALSE, grp.lasso,
    10 lambdas, irls.tol 1e-3, tol 1e-8): oem_fit_logistic_dense (upper-bound Hessian) against oem_fit_logistic_sparse
    (Hessian rebuilt every IRLS pass) on the same data.  true.beta and x come from the seeded R stream; xs goes through
    Matrix::rsparsematrix and ys through rbinom, which are not restated, so a statistically equivalent stand-in (same
    shape, density and value distribution; ys drawn with the example's own recycled probabilities) is used and only the
    ORDER of the difference is comparable.  Returns (our max |diff|, printed value)."""
    import scipy.sparse as sps
    g = printed()["oem_rd"]["maxdiff_logistic_dense_vs_sparse_grp_lasso"]["values"][0]
    r = RStream(123)
    n_obs, n_vars = 10000, 50
    true_beta = np.concatenate([r.runif(15, -0.25, 0.25), np.zeros(n_vars - 15)])
    x = r.matrix_rnorm(n_obs, n_vars)
    rng = np.random.default_rng(123)
    xs = sps.random(2 * n_obs, n_vars, density=0.01, random_state=np.random.RandomState(123), format="csc",
                    data_rvs=rng.standard_normal)
    prob = 1.0 / (1.0 + np.exp(-(x @ true_beta)))
    ys = (rng.uniform(size=2 * n_obs) < np.tile(prob, 2)).astype(np.float64)      # rbinom(n.obs * 2, 1, prob) recycles prob
    groups = np.repeat(np.arange(1, 6), 10)
    kw = dict(family="binomial", penalty="grp.lasso", intercept=False, nlambda=10, groups=groups, irls_tol=1e-3, tol=1e-8)
    res_gr = fe.oem(xs.toarray(), ys, **kw)
    res_gr_s = fe.oem(xs, ys, **kw)
    d = float(np.max(np.abs(np.asarray(res_gr["beta"]["grp.lasso"]) - np.asarray(res_gr_s["beta"]["grp.lasso"]))))
    return d, g


def assert_printed(results, units=0.51):
    """|got - printed| <= `units` of the last printed digit (0.5 = correct rounding; + 2 % for the decimal -> binary
    conversion of the printed value itself)."""
    for key, (got, want, unit) in results.items():
        assert got.shape == want.shape, (key, got.shape, want.shape)
        err = np.abs(got - want) / unit
        assert np.all(err <= units), f"{key}: off by {err.max():.3f} units of the last printed digit at index {int(err.argmax())}: " \
                                     f"got {got[err.argmax()]!r}, reference printed {want[err.argmax()]!r}"
