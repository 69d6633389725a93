"""Seeded synthetic problems shaped like BASELINE.json's configs, at sizes the CPU
oracle finishes in seconds, and helpers that call the oracle and the CUDA path with IDENTICAL
reference-style argument lists."""
import numpy as np

DEFAULT_OPTS = dict(maxit=500, tol=1e-7, irls_maxit=100, irls_tol=1e-3, ncores=1,
                    hessian_type="upper.bound", accelerate=False)


def gaussian_problem(seed, n, p, nnz=25, sd_x=1.0, mean_x=0.0, noise=1.0, coef="unif"):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.normal(mean_x, sd_x, size=(n, p)))
    b = np.zeros(p)
    k = min(nnz, p)
    if coef == "unif":
        b[:k] = rng.uniform(-0.5, 0.5, size=k)
    else:
        b[:min(5, p)] = [0.5, 0.5, -0.5, -0.5, 1.0][:min(5, p)]
    y = X @ b + rng.normal(0.0, noise, size=n)
    return X, y


def binomial_problem(seed, n, p):
    rng = np.random.default_rng(seed)
    X = np.asfortranarray(rng.normal(size=(n, p)))
    b = np.zeros(p)
    b[:min(5, p)] = [0.15, 0.15, -0.15, -0.15, 0.25][:min(5, p)]
    pr = 1.0 / (1.0 + np.exp(-(X @ b)))
    y = (rng.uniform(size=n) < pr).astype(np.float64)
    return X, y


def args_xy(X, y, family, penalty, *, groups=None, unique_groups=None, group_weights=None, lambda_=None,
            nlambda=100, lmin_ratio=1e-4, alpha=1.0, gamma=3.0, tau=0.5, penalty_factor=None,
            standardize=True, intercept=True, compute_loss=False, opts=None):
    """Positional argument list of oem_fit_dense / oem_fit_big / oem_fit_logistic_dense."""
    p = X.shape[1]
    o = dict(DEFAULT_OPTS)
    o.update(opts or {})
    return [X, y, family, list(penalty), [], groups if groups is not None else [],
            unique_groups if unique_groups is not None else [], group_weights if group_weights is not None else [],
            lambda_ if lambda_ is not None else [], nlambda, lmin_ratio, alpha, gamma, tau,
            penalty_factor if penalty_factor is not None else np.ones(p), standardize, intercept, compute_loss, o]


def max_beta_diff(a, b):
    return max(float(np.max(np.abs(np.asarray(x) - np.asarray(y)))) if np.asarray(x).size else 0.0
               for x, y in zip(a["beta"], b["beta"]))


def assert_same_fit(got, ref, tol=1e-8, check_niter=True, lam_ulps=None, lam_rtol=1e-12):
    """The parity bar of BASELINE.json: max |delta beta| <= 1e-8 and equal lambda sequences.
    The grid itself (log-spaced from lmax, host libm) is bit-identical given the same lmax -- checked with
    lam_ulps=4 where lmax is an INPUT (oem_xtx).  Where lmax = max|X'y|/n is a sum over n rows, the two
    summation orders differ in the last bits, so the sequences are compared to 1e-12 relative."""
    assert len(got["beta"]) == len(ref["beta"])
    for pp in range(len(ref["beta"])):
        lr, lg = np.asarray(ref["lambda_"][pp]), np.asarray(got["lambda_"][pp])[:len(ref["lambda_"][pp])]
        if lam_ulps is not None:
            assert np.all(np.abs(lg - lr) <= lam_ulps * np.spacing(np.abs(lr))), f"lambda mismatch, penalty {pp}"
        assert np.all(np.abs(lg - lr) <= lam_rtol * np.abs(lr)), f"lambda mismatch, penalty {pp}"
        bg, br = np.asarray(got["beta"][pp]), np.asarray(ref["beta"][pp])
        assert bg.shape == br.shape, (bg.shape, br.shape)
        diff = float(np.max(np.abs(bg - br)))
        assert diff <= tol, f"penalty {pp}: max|dbeta| = {diff:.3e} > {tol}"
        if check_niter:
            ng, nr = np.asarray(got["niter"][pp]), np.asarray(ref["niter"][pp])
            # an iteration-count flip (1-ulp difference at the stop threshold) is tolerated on a few lambdas
            assert np.mean(ng != nr) <= 0.1 and np.max(np.abs(ng - nr)) <= 2, (ng, nr)
    assert abs(got["d"] - ref["d"]) <= 1e-9 * abs(ref["d"]), (got["d"], ref["d"])


def sparse_problem(seed, n, p, density=0.05, nnz=10, noise=0.5, shift_y=0.0, empty_cols=(), empty_rows=0):
    """A dgCMatrix-like design (scipy CSC, sorted indices) in the style of man/oem.Rd:104-112 (rsparsematrix + rnorm)."""
    import scipy.sparse as sps
    rng = np.random.default_rng(seed)
    X = sps.random(n, p, density=density, random_state=np.random.RandomState(seed), format="lil",
                   data_rvs=rng.standard_normal)
    for j in empty_cols:
        X[:, j] = 0.0
    if empty_rows:
        X[:empty_rows, :] = 0.0
    X = sps.csc_matrix(X)
    X.eliminate_zeros()
    X.sort_indices()
    b = np.zeros(p)
    b[:min(nnz, p)] = rng.uniform(-1.0, 1.0, size=min(nnz, p))
    y = X @ b + rng.normal(0.0, noise, size=n) + shift_y
    return X, y
