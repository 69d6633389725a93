"""-m gpu: the exchange modes and mat-vec variants of the path kernel at the coefficient counts where they switch
(csrc/path_kernel.cu): one CTA, a DSMEM cluster, and the L2 ("global") mode with its register-resident DFMA mat-vec
(RPT = ceil(q / 256) = 2..5 rows per thread, <= 4 chains per Gram, one 8-column atom per member) or the DMMA mat-vec
(more chains, q > 1184).  oem.xtx is the entry that reaches the path kernel with nothing in front of it: X'X / n and
X'y / n are inputs, so every case costs one small Gram on the host and the CPU oracle's own iterations."""
import numpy as np
import pytest

from cases import assert_same_fit, gaussian_problem

pytestmark = pytest.mark.gpu

PENS = ["lasso", "mcp", "scad", "elastic.net", "scad.net"]


def _xtx(seed, p, n=None):
    n = n or 2 * p + 100
    X, y = gaussian_problem(seed, n, p, nnz=12)
    return X.T @ X / n, X.T @ y / n


def _common(pens, p, group=5, nlambda=7, alpha=0.7, gamma=3.3):
    g = np.arange(p) // group + 1
    return ["gaussian", pens, g, np.unique(g), [], [], nlambda, 0.05, alpha, gamma, 0.4]


@pytest.mark.parametrize("p,nchains", [(150, 2), (300, 1), (440, 3),           # single CTA / cluster
                                       (460, 1), (512, 4), (700, 2), (768, 3),  # global, RPT = 2, 3
                                       (1001, 3), (1024, 1), (1150, 4),         # RPT = 4, 5
                                       (600, 5), (1300, 2)])                   # DMMA mat-vec: > 4 chains, two atoms per member
def test_path_modes_coordinate_penalties(lib, oracle, p, nchains):
    xtx, xty = _xtx(40 + p, p)
    c = _common(PENS[:nchains], p)
    o = dict(maxit=400, tol=1e-9)
    got = lib.oem_xtx(xtx, xty, *c, [], np.ones(p), o)
    ref = oracle.oem_xtx(xtx, xty, *c, [], np.ones(p), o)
    assert_same_fit(got, ref, lam_ulps=4)
    assert sum(int(np.count_nonzero(b[:, -1])) for b in ref["beta"]) > 10      # the paths do leave zero


@pytest.mark.parametrize("p", [480, 1001])
def test_path_global_mode_group_penalties_scale_factor_and_maxit(lib, oracle, p):
    # group penalties publish u and every member applies the prox (replicated chains next to coordinate-wise ones);
    # scale.factor rescales the iterate in place after every lambda (src/oem_xtx.h:576-581); maxit exhaustion -> maxit + 1
    xtx, xty = _xtx(7 + p, p)
    pf = np.ones(p); pf[3] = 0.0; pf[10] = 2.5
    c = _common(["grp.lasso", "lasso", "sparse.grp.lasso", "grp.mcp"], p, group=7)
    o = dict(maxit=300, tol=1e-9)
    sf = np.sqrt(np.diag(xtx))
    got = lib.oem_xtx(xtx, xty, *c, sf, pf, o)
    ref = oracle.oem_xtx(xtx, xty, *c, sf, pf, o)
    assert_same_fit(got, ref)
    o = dict(maxit=3, tol=1e-12)
    c = _common(["lasso", "mcp"], p)
    got = lib.oem_xtx(xtx, xty, *c, [], np.ones(p), o)
    ref = oracle.oem_xtx(xtx, xty, *c, [], np.ones(p), o)
    assert_same_fit(got, ref, lam_ulps=4)
    assert max(int(n.max()) for n in got["niter"]) == 4


@pytest.mark.parametrize("p", [500, 1001])
def test_path_global_mode_nesterov(lib, oracle, p):
    # accelerate = TRUE exists in oem_fit_dense only (src/oem_dense.h:633-651): every chain is a replicated one
    from cases import args_xy
    X, y = gaussian_problem(90 + p, 2 * p + 50, p, nnz=15)
    a = args_xy(X, y, "gaussian", ["lasso", "mcp"], nlambda=6, lmin_ratio=0.1, standardize=False, intercept=False,
                opts=dict(tol=1e-9, maxit=300, accelerate=True))
    assert_same_fit(lib.oem_fit_dense(*a), oracle.oem_fit_dense(*a))


def test_path_global_mode_reruns_are_bit_identical(lib):
    xtx, xty = _xtx(5, 1001)
    c = _common(["lasso", "scad", "mcp"], 1001)
    o = dict(maxit=300, tol=1e-8)
    a = lib.oem_xtx(xtx, xty, *c, [], np.ones(1001), o)
    b = lib.oem_xtx(xtx, xty, *c, [], np.ones(1001), o)
    for x, z in zip(a["beta"], b["beta"]):
        assert np.array_equal(x, z)
    assert all(np.array_equal(x, z) for x, z in zip(a["niter"], b["niter"])) and a["d"] == b["d"]
