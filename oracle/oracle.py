"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (pinned on the reference's printed outputs for the four
gaussian entry points; the logistic entry stays "parity unpinned").

CPU restatement (numpy + the plain-C iteration in oracle/oem_oracle.c) of the five C++
entry points of jaredhuling/oem 2.0.12 that make up the hot path:

    oem_fit_dense            src/oem_dense.cpp:30-309,   src/oem_dense.h, src/DataStd.h
    oem_xtx                  src/oem_xtx.cpp:29-219,     src/oem_xtx.h
    oem_xval_dense           src/oem_xval_dense.cpp:31-477, src/oem_xval_dense.h
    oem_fit_logistic_dense   src/oem_logistic_dense.cpp:29-313, src/oem_logistic_dense.h
    oem_fit_big              src/oem_big.cpp:30-258,     src/oem_big.h

Each function keeps the reference's argument order and returns the reference's named list
as a dict (beta / lambda / niter / loss / d [/ cvm / cvsd]).  Only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product never does.

The reference cannot be built here (no R / Rcpp / Eigen / Spectra) and ships no test suite.
What it does ship is rendered documentation (docs/reference/*.html, vignettes/oem_vignette.html)
with the values its own oem(), xval.oem(), oem.xtx() and big.oem() returned on seeded inputs.
oracle/r_rng.py restates R's set.seed / runif / rnorm stream, tests/reference_examples.py re-runs
those examples and tests/test_reference_pins.py requires this oracle (and, on the GPU, the CUDA
path) to round to every digit the reference printed: 28 test-set MSEs (6-7 digits), 300
log-likelihoods of compute.loss paths (7 digits), max |big.oem - oem| = 1.534783e-05 (7 digits)
and the oem == oem.xtx identity.  No seeded binomial example prints a value, so
oem_fit_logistic_dense is pinned only by identities (tests/test_oracle.py) -- "parity unpinned".

Third-party arithmetic restated (not vendored under /root/reference):
  * Eigen (RcppEigen, unpinned): SYRK/GEMV = numpy/OpenBLAS here; LinSpaced = linspace_eigen().
  * Spectra::SymEigsSolver (RSpectra >= 0.16-2), nev=1, tol 1e-10: the oracle defines the top
    eigenvalue as the converged one (numpy.linalg.eigvalsh).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PENALTIES = ["lasso", "ols", "elastic.net", "scad", "scad.net", "mcp", "mcp.net", "grp.lasso",
             "grp.lasso.net", "grp.mcp", "grp.scad", "grp.mcp.net", "grp.scad.net", "sparse.grp.lasso"]
PEN_ID = {name: i for i, name in enumerate(PENALTIES)}


class _Pen(ctypes.Structure):
    _fields_ = [("q", ctypes.c_int), ("penalty", ctypes.c_int),
                ("alpha", ctypes.c_double), ("gamma", ctypes.c_double), ("tau", ctypes.c_double),
                ("pen_fact", ctypes.c_void_p), ("ngroups", ctypes.c_int),
                ("unique_groups", ctypes.c_void_p), ("grp_ptr", ctypes.c_void_p),
                ("grp_idx", ctypes.c_void_p), ("group_weights", ctypes.c_void_p)]


def build(force=False):
    """Compile oracle/oem_oracle.c into oracle/_build/liboem_oracle.so (gcc, OpenMP)."""
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "liboem_oracle.so")
    src = os.path.join(_HERE, "oem_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-march=x86-64-v2", "-fopenmp", "-fPIC", "-shared", "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_solve.restype = ctypes.c_int
        _LIB.oracle_solve.argtypes = [ctypes.POINTER(_Pen), ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_double,
                                      ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_void_p]
        _LIB.oracle_stop_rule.restype = ctypes.c_int
        _LIB.oracle_stop_rule.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double]
        _LIB.oracle_xtx.restype = None
        _LIB.oracle_xtx.argtypes = [ctypes.c_long, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    return _LIB


def stop_rule(cur, prev, tol):
    """src/utils.cpp:537-549."""
    cur = np.ascontiguousarray(cur, dtype=np.float64)
    prev = np.ascontiguousarray(prev, dtype=np.float64)
    return bool(_lib().oracle_stop_rule(cur.size, cur.ctypes.data, prev.ctypes.data, float(tol)))


def xtx_port(X, ncores=1):
    """Row-sliced lower SYRK restated in plain C (src/oem_dense.h:318-361)."""
    X = np.asfortranarray(X, dtype=np.float64)
    n, p = X.shape
    G = np.zeros((p, p), order="F")
    _lib().oracle_xtx(n, p, X.ctypes.data, G.ctypes.data, int(ncores))
    return G


def linspace_eigen(N, lo, hi):
    """Eigen >= 3.3 LinSpaced (linspaced_op_impl, not vendored): see SURVEY.md A.1."""
    N = int(N)
    if N == 1:
        return np.array([lo], dtype=np.float64)
    step = (hi - lo) / float(N - 1)
    v = np.empty(N, dtype=np.float64)
    if abs(hi) < abs(lo):
        for i in range(N):
            v[i] = hi - float(N - 1 - i) * step
        v[0] = lo
    else:
        for i in range(N):
            v[i] = lo + float(i) * step
        v[N - 1] = hi
    return v


def lambda_base(lmax, nl, lmin_ratio):
    """src/oem_dense.cpp:179-186: exp(LinSpaced(nl, log(lmax), log(lmin_ratio*lmax)))."""
    import math
    lmin = lmin_ratio * lmax
    v = linspace_eigen(nl, math.log(lmax), math.log(lmin))
    return np.array([math.exp(t) for t in v], dtype=np.float64)


def top_eig(XX):
    """Converged largest algebraic eigenvalue (stands in for Spectra)."""
    return float(np.linalg.eigvalsh(XX)[-1])


class _Groups:
    """get_group_indexes (src/oem_dense.h:421-456): members of each unique group among the first
    `scan` entries of `groups`, default weights sqrt(|g|); logistic sets w=0 for group 0
    (src/oem_logistic_dense.h:421-437)."""

    def __init__(self, groups, unique_groups, group_weights, scan, zero_weight_for_group0=False):
        groups = np.asarray(groups, dtype=np.int32).ravel()
        self.unique = np.ascontiguousarray(np.asarray(unique_groups, dtype=np.int32).ravel())
        ptr, idx = [0], []
        for g in self.unique:
            members = [v for v in range(min(scan, groups.size)) if groups[v] == g]
            idx.extend(members)
            ptr.append(len(idx))
        self.ptr = np.ascontiguousarray(np.array(ptr, dtype=np.int32))
        self.idx = np.ascontiguousarray(np.array(idx if idx else [0], dtype=np.int32))
        gw = np.asarray(group_weights, dtype=np.float64).ravel()
        if gw.size < 1:
            gw = np.sqrt(np.diff(self.ptr).astype(np.float64))
            if zero_weight_for_group0:
                gw = np.where(self.unique == 0, 0.0, gw)
        self.weights = np.ascontiguousarray(gw)


class _Solver:
    """State shared by the five solvers: beta (warm start), Nesterov ak, penalty descriptor."""

    def __init__(self, q, pen_fact, grp, maxit, tol, accelerate=False):
        self.q = q
        self.pf = np.ascontiguousarray(pen_fact, dtype=np.float64)
        self.grp = grp
        self.maxit, self.tol, self.accelerate = int(maxit), float(tol), bool(accelerate)
        self.beta = np.zeros(q)
        self.ak = ctypes.c_double(1.0)

    def init(self, penalty, alpha, gamma, tau):
        # init(): src/oem_dense.h:723-746 -- beta = 0, ak = 1
        self.beta = np.zeros(self.q)
        self.ak = ctypes.c_double(1.0)
        g = self.grp
        self.pen = _Pen(self.q, PEN_ID[penalty], float(alpha), float(gamma), float(tau),
                        self.pf.ctypes.data, int(g.unique.size), g.unique.ctypes.data,
                        g.ptr.ctypes.data, g.idx.ctypes.data, g.weights.ctypes.data)

    def solve(self, A, XY, d, lam):
        A = np.asfortranarray(A)
        XY = np.ascontiguousarray(XY)
        return _lib().oracle_solve(ctypes.byref(self.pen), A.ctypes.data, XY.ctypes.data, float(d),
                                   float(lam), self.maxit, self.tol, int(self.accelerate),
                                   ctypes.byref(self.ak), self.beta.ctypes.data)


def _lambda_lists(lambda_, penalty, nlambda, lmin_ratio, lmax, alpha, logistic_fudge=None, gamma=None):
    """Per-penalty lambda vectors: src/oem_dense.cpp:179-227 (logistic .net mcp/scad fudge:
    src/oem_logistic_dense.cpp:210-225)."""
    provided = len(lambda_) > 0 and np.asarray(lambda_[0]).size >= 1
    base = None if provided else lambda_base(lmax, nlambda, lmin_ratio)
    out = []
    for pp, pen in enumerate(penalty):
        if provided:
            lam = np.asarray(lambda_[pp], dtype=np.float64).copy()
        elif ".net" in pen:
            lam = base / alpha
            if logistic_fudge and ("mcp" in pen or "scad" in pen):
                fact = 3.5 - min(3.5, gamma) * 5.71425 / 8.0
                lam = fact * base / (alpha ** 0.8)
        else:
            lam = base.copy()
        out.append(lam)
    return out


def _as_opts(opts):
    o = dict(maxit=500, tol=1e-7, irls_maxit=100, irls_tol=1e-3, ncores=1,
             hessian_type="upper.bound", accelerate=False, gigs=4.0)
    for k, v in (opts or {}).items():
        o[k.replace(".", "_")] = v
    return o


def _gamma_for(gamma, pp):
    """API extension shared with the product: gamma may be one value per penalty."""
    g = np.atleast_1d(np.asarray(gamma, dtype=np.float64))
    return float(g[pp] if g.size > 1 else g[0])


# ------------------------------------------------------------------------------------------
# oem_fit_dense
# ------------------------------------------------------------------------------------------
def standardize(X, Y, standardize_, intercept):
    """DataStd::standardize, no-weights branch (src/DataStd.h:94-267).  Returns
    (Xs, Ys, meanX, scaleX, meanY, scaleY); quirk: flag 2 falls through into flag 3 for y."""
    flag = int(bool(standardize_)) + 2 * int(bool(intercept))
    n, p = X.shape
    X = np.array(X, dtype=np.float64, order="F", copy=True)
    Y = np.array(Y, dtype=np.float64, copy=True)
    meanY, scaleY = 0.0, 1.0
    meanX, scaleX = np.zeros(p), np.ones(p)
    n_invsqrt = 1.0 / np.sqrt(float(n))
    if flag == 1:
        scaleY = float(np.linalg.norm(Y - Y.mean()) / np.sqrt(float(n)))
        Y /= scaleY
    elif flag in (2, 3):
        meanY = float(Y.mean())
        Y -= meanY
        scaleY = float(np.linalg.norm(Y) * n_invsqrt)
        Y /= scaleY
    for i in range(p):
        if flag == 1:
            col = X[:, i]
            s = float(np.linalg.norm(col - col.mean()) / np.sqrt(float(n)))
            scaleX[i] = 1.0 if s == 0.0 else s
            X[:, i] *= (1.0 / scaleX[i])
        elif flag == 2:
            meanX[i] = X[:, i].mean()
            X[:, i] -= meanX[i]
        elif flag == 3:
            meanX[i] = X[:, i].mean()
            X[:, i] -= meanX[i]
            s = float(np.linalg.norm(X[:, i]) * n_invsqrt)
            scaleX[i] = 1.0 if s == 0.0 else s
            X[:, i] /= scaleX[i]
    return X, Y, meanX, scaleX, meanY, scaleY, flag


def recover(flag, coef, meanX, scaleX, meanY, scaleY):
    """DataStd::recover (src/DataStd.h:269-293)."""
    coef = coef.copy()
    beta0 = 0.0
    if flag == 1:
        coef /= scaleX
        coef *= scaleY
    elif flag == 2:
        coef *= scaleY
        beta0 = meanY - float((coef * meanX).sum())
    elif flag == 3:
        coef /= scaleX
        coef *= scaleY
        beta0 = meanY - float((coef * meanX).sum())
    return beta0, coef


def oem_fit_dense(x, y, family, penalty, weights, groups, unique_groups, group_weights, lambda_,
                  nlambda, lmin_ratio, alpha, gamma, tau, penalty_factor, standardize_, intercept,
                  compute_loss, opts, xtx_fn=None):
    """src/oem_dense.cpp:30-309."""
    o = _as_opts(opts)
    if family != "gaussian":
        raise ValueError("binomial not available for oem_fit_dense, use oem_fit_logistic_dense")
    if np.asarray(weights).size:
        raise ValueError("weights not implemented yet.")   # R/oem.R:244
    x = np.asarray(x, dtype=np.float64)
    n, p = x.shape
    X, Y, meanX, scaleX, meanY, scaleY, flag = standardize(x, y, standardize_, intercept)
    # init_oem: src/oem_dense.h:693-712, compute_XtX_d_update_A :458-506
    XY = X.T @ Y / n
    XX = (xtx_fn(X) if xtx_fn else X.T @ X) / n
    if n > p:
        d = top_eig(XX) * 1.005
    else:
        # src/oem_dense.h:474-483: d comes from the n x n matrix XX'/n; next_u (:515-521) is then
        # X'(Y - X beta)/n + d beta, which is (dI - X'X/n) beta + XY term by term -- iterated below in that form
        d = top_eig(X @ X.T / n) * 1.005
    A = -XX
    A[np.diag_indices(p)] += d
    lmax = float(np.abs(XY).max()) * scaleY
    lams = _lambda_lists(lambda_, penalty, nlambda, lmin_ratio, lmax, alpha)
    grp = _Groups(groups, unique_groups, group_weights, scan=p)
    s = _Solver(p, penalty_factor, grp, o["maxit"], o["tol"], o["accelerate"])
    out = dict(beta=[], lambda_=[], niter=[], loss=[], d=d)
    for pp, pen in enumerate(penalty):
        lam = lams[pp]
        L = 1 if pen == "ols" else lam.size
        beta = np.zeros((p + 1, L), order="F")
        niter = np.zeros(L, dtype=np.int32)
        loss = np.full(L, 1e99)
        s.init(pen, alpha, _gamma_for(gamma, pp), tau)
        for i in range(L):
            niter[i] = s.solve(A, XY, d, lam[i] / scaleY)
            b0, res = recover(flag, s.beta, meanX, scaleX, meanY, scaleY)
            beta[0, i] = b0
            beta[1:, i] = res
            if compute_loss:
                loss[i] = float(((Y - X @ s.beta) ** 2).sum())
        out["beta"].append(beta)
        out["lambda_"].append(lam)
        out["niter"].append(niter)
        out["loss"].append(loss)
    return out


# ------------------------------------------------------------------------------------------
# oem_xtx
# ------------------------------------------------------------------------------------------
def oem_xtx(xtx, xty, family, penalty, groups, unique_groups, group_weights, lambda_, nlambda,
            lmin_ratio, alpha, gamma, tau, scale_factor, penalty_factor, opts):
    """src/oem_xtx.cpp:29-219, src/oem_xtx.h:347-381,520-581."""
    o = _as_opts(opts)
    if family != "gaussian":
        raise ValueError("binomial not available for oem_fit_dense, use oem_fit_logistic_dense")
    XX = np.array(xtx, dtype=np.float64, order="F")
    p = XX.shape[1]
    XY = np.array(xty, dtype=np.float64).ravel().copy()
    sf = np.asarray(scale_factor, dtype=np.float64).ravel()
    sinv = None
    if sf.size:
        sinv = 1.0 / sf
        XY = XY * sinv
        XX = sinv[:, None] * XX * sinv[None, :]
    d = top_eig(XX) * 1.005
    A = -XX
    A[np.diag_indices(p)] += d
    lmax = float(np.abs(XY).max())
    lams = _lambda_lists(lambda_, penalty, nlambda, lmin_ratio, lmax, alpha)
    grp = _Groups(groups, unique_groups, group_weights, scan=p)
    s = _Solver(p, penalty_factor, grp, o["maxit"], o["tol"])
    out = dict(beta=[], lambda_=[], niter=[], loss=[], d=d)
    for pp, pen in enumerate(penalty):
        lam = lams[pp]
        L = 1 if pen == "ols" else lam.size
        beta = np.zeros((p, L), order="F")
        niter = np.zeros(L, dtype=np.int32)
        s.init(pen, alpha, _gamma_for(gamma, pp), tau)
        for i in range(L):
            niter[i] = s.solve(A, XY, d, lam[i])
            if sinv is not None:
                s.beta *= sinv          # get_beta() mutates the iterate (src/oem_xtx.h:576-581)
            beta[:, i] = s.beta
        out["beta"].append(beta)
        out["lambda_"].append(lam)
        out["niter"].append(niter)
        out["loss"].append(np.full(L, 1e99))
    return out


# ------------------------------------------------------------------------------------------
# oem_xval_dense
# ------------------------------------------------------------------------------------------
def fold_parts(X, Y, foldid, nfolds, intercept, weights=None):
    """XtX_xval / XtX_xval_int (src/oem_xval_dense.h:358-484): per-fold Gram pieces; with observation
    weights the twins XtWX_xval / XtWX_xval_int (:489-627): G = sub' W sub, b = sub' (y*w), border
    sum_i w_i x_i, corner sum w, b(0) = sum y*w, while nobs stays the row COUNT and colsq stays the
    UNWEIGHTED sum of squares (:533-535 "we do not standardize with respect to weights").
    The reference's intercept twin writes rankUpdate(sqrt(W).asDiagonal() * sub.adjoint()) (:596-597),
    a numelem x numelem diagonal times an nvars x numelem matrix -- dimensionally invalid (undefined
    behaviour under NDEBUG); its no-intercept twin (:528-529) and every other weighted term define the
    intent sub' W sub, which is what is restated here for both."""
    p = X.shape[1]
    parts = []
    for k in range(1, nfolds + 1):
        idx = np.nonzero(foldid == k)[0]
        sub, sub_y = X[idx, :], Y[idx]
        if weights is None:
            G = sub.T @ sub
            b = sub.T @ sub_y
            cs, corner, b0 = sub.sum(axis=0), float(idx.size), sub_y.sum()
        else:
            w = weights[idx]
            sw = np.sqrt(w)[:, None] * sub
            G = sw.T @ sw
            yw = sub_y * w
            b = sub.T @ yw
            cs, corner, b0 = (w[:, None] * sub).sum(axis=0), w.sum(), yw.sum()
        if intercept:
            Gi = np.zeros((p + 1, p + 1))
            Gi[1:, 1:] = G
            Gi[0, 1:] = cs
            Gi[1:, 0] = cs
            Gi[0, 0] = corner
            G = Gi
            b = np.concatenate([[b0], b])
        parts.append(dict(xtx=G, xty=b, nobs=idx.size, colsq=(sub ** 2).sum(axis=0)))
    return parts


def assemble(parts, skip_fold, p, standardize_, intercept):
    """compute_/update_XtX_d_update_A (src/oem_xval_dense.h:731-788, 791-853): sum the parts
    except fold `skip_fold` (1-based; 0 = none), uncentred scaling, /n, top eigenvalue."""
    q = p + int(intercept)
    XX, XY, colsq, nobs = np.zeros((q, q)), np.zeros(q), np.zeros(p), 0
    for k, part in enumerate(parts, start=1):
        if k != skip_fold:
            XX += part["xtx"]
            XY += part["xty"]
            nobs += part["nobs"]
            colsq += part["colsq"]
    colsq = colsq / (float(nobs) - 1.0)
    colsq = np.where(colsq == 0.0, 1.0, colsq)
    colsq_inv = 1.0 / np.sqrt(colsq)
    if standardize_:
        if intercept:
            XX[1:, 1:] = colsq_inv[:, None] * XX[1:, 1:] * colsq_inv[None, :]
            XX[0, 1:] *= colsq_inv
            XX[1:, 0] *= colsq_inv
            XY[1:] *= colsq_inv
        else:
            XX = colsq_inv[:, None] * XX * colsq_inv[None, :]
            XY *= colsq_inv
    XX /= nobs
    XY /= nobs
    d = top_eig(XX) * 1.005
    if not nobs > p:
        raise ValueError("dimension of x larger than number of observations")
    A = -XX
    A[np.diag_indices(q)] += d
    return A, XY, d, colsq_inv, nobs


def oem_xval_dense(x, y, family, penalty, weights, groups, unique_groups, group_weights, lambda_,
                   nlambda, lmin_ratio, alpha, gamma, tau, penalty_factor, standardize_, intercept,
                   nfolds, foldid, compute_loss, type_measure, opts):
    """src/oem_xval_dense.cpp:31-477, ncores=1 semantics; observation weights
    (reachable through xval.oem(weights=), R/oem_xval.R:215-222) enter the fold Grams (fold_parts)
    and the CV score t*w_i (oem_xval_dense.cpp:389-392, 398-401); the loss stays unweighted
    (oem_xval_dense.h:1122-1145)."""
    o = _as_opts(opts)
    if family != "gaussian":
        raise ValueError("binomial not available for oem_xval_dense, use oem_xval_logistic_dense")
    X = np.asarray(x, dtype=np.float64)
    Y = np.asarray(y, dtype=np.float64).ravel()
    n, p = X.shape
    W = np.asarray(weights, dtype=np.float64).ravel()
    if W.size == 0:
        W = None
    elif W.size != n:
        raise ValueError("length of weights not same as number of observations in x")   # R/oem_xval.R:218-221
    foldid = np.asarray(foldid, dtype=np.int32).ravel()
    q = p + int(intercept)
    pf = np.asarray(penalty_factor, dtype=np.float64).ravel()
    if intercept:
        pf = np.concatenate([[0.0], pf])
    if not n > p:
        raise ValueError("dimension of x larger than number of observations")
    parts = fold_parts(X, Y, foldid, nfolds, intercept, W)
    grp = _Groups(groups, unique_groups, group_weights, scan=q)
    s = _Solver(q, pf, grp, o["maxit"], o["tol"])
    out = dict(beta=[None] * len(penalty), lambda_=[None] * len(penalty), niter=[None] * len(penalty),
               loss=[None] * len(penalty), cvm=[], cvsd=[], d=None)
    beta_folds = [[None] * nfolds for _ in penalty]
    lams = None
    for ff in range(nfolds + 1):
        A, XY, d, colsq_inv, nobs = assemble(parts, ff, p, standardize_, intercept)
        if ff == 0:
            out["d"] = d
            lmax = float(np.abs(XY[1:] if intercept else XY).max())
            lams = _lambda_lists(lambda_, penalty, nlambda, lmin_ratio, lmax, alpha)
        for pp, pen in enumerate(penalty):
            lam = lams[pp]
            L = 1 if pen == "ols" else lam.size
            beta = np.zeros((p + 1, L), order="F")
            niter = np.zeros(L, dtype=np.int32)
            loss = np.full(L, 1e99)
            s.init(pen, alpha, _gamma_for(gamma, pp), tau)
            for i in range(L):
                niter[i] = s.solve(A, XY, d, lam[i])
                res = s.beta.copy()               # get_beta(): src/oem_xval_dense.h:1102-1120
                if standardize_:
                    if intercept:
                        res[1:] *= colsq_inv
                    else:
                        res *= colsq_inv
                if intercept:
                    beta[:, i] = res
                else:
                    beta[1:, i] = res
                if compute_loss and ff == 0:
                    loss[i] = float(((Y - X @ beta[1:, i] - beta[0, i]) ** 2).sum())
            if ff == 0:
                out["beta"][pp], out["lambda_"][pp] = beta, lam
                out["niter"][pp], out["loss"][pp] = niter, loss
            else:
                beta_folds[pp][ff - 1] = beta
    # CV scoring: src/oem_xval_dense.cpp:345-464.  Welford over rows in original order gives the
    # mean and M2 = sum (t - mean)^2; evaluated here per fold block with the same definitions.
    for pp, pen in enumerate(penalty):
        L = beta_folds[pp][0].shape[1]
        T = np.empty((n, L))
        for k in range(1, nfolds + 1):
            idx = np.nonzero(foldid == k)[0]
            B = beta_folds[pp][k - 1]
            r = Y[idx, None] - (X[idx, :] @ B[1:, :] + B[0:1, :])
            T[idx, :] = r ** 2 if type_measure == "mse" else np.abs(r)
            if W is not None:
                T[idx, :] *= W[idx, None]
        m = T.mean(axis=0)
        M2 = ((T - m[None, :]) ** 2).sum(axis=0)
        out["cvm"].append(m)
        out["cvsd"].append(np.sqrt(M2 / float(n - 1)) / np.sqrt(float(n)))
    return out


def welford_rows(T):
    """Literal row-order Welford of src/oem_xval_dense.cpp:407-412 (small n only; used by tests
    to pin the blocked evaluation above)."""
    n, L = T.shape
    m, ss = np.zeros(L), np.zeros(L)
    for i in range(n):
        delta = T[i] - m
        m = m + delta / (i + 1)
        ss = ss + delta * (T[i] - m)
    return m, ss


# ------------------------------------------------------------------------------------------
# oem_fit_big
# ------------------------------------------------------------------------------------------
def oem_fit_big(x, y, family, penalty, weights, groups, unique_groups, group_weights, lambda_,
                nlambda, lmin_ratio, alpha, gamma, tau, penalty_factor, standardize_, intercept,
                compute_loss, opts, xtx_fn=None):
    """src/oem_big.cpp:30-258, src/oem_big.h:469-566,731-842; unweighted."""
    o = _as_opts(opts)
    if family != "gaussian":
        raise ValueError("binomial not available for oem_fit_dense, use oem_fit_logistic_dense")
    if np.asarray(weights).size:
        raise ValueError("weights not implemented yet.")   # R/big_oem.R:178
    X = np.asarray(x, dtype=np.float64)
    Y = np.asarray(y, dtype=np.float64).ravel()
    n, p = X.shape
    q = p + int(intercept)
    pf = np.asarray(penalty_factor, dtype=np.float64).ravel()
    if intercept:
        pf = np.concatenate([[0.0], pf])
    # nslices row slices for X'X (src/oem_big.h:319-361, 738-741): floor(n / nslices) rows each, the last takes the rest
    nslices = max(1, int(np.ceil(8.0 * float(n) * float(p) / 1e9 / float(o.get("gigs", 4.0)))))
    colsq_inv = np.ones(p)
    if nslices <= 1:
        if standardize_:
            colsq = (X ** 2).sum(axis=0) / (float(n) - 1.0)
            colsq = np.where(colsq == 0.0, 1.0, colsq)
            colsq_inv = 1.0 / np.sqrt(colsq)
        xty = X.T @ Y
        G = xtx_fn(X) if xtx_fn else X.T @ X
        colsums = X.sum(axis=0) if intercept else None
    else:
        # the same sums without n x p temporaries: column-wise sweeps like the reference's loops (src/oem_big.h:743-837)
        # in blocks of columns, X'X slice by slice
        colsq, colsums = np.zeros(p), np.zeros(p)
        xty = X.T @ Y                                    # X.col(j).dot(Y) for every j: one GEMV, no temporaries
        if intercept:
            colsums = np.ones(n) @ X                     # X.col(j).sum()
        if standardize_:
            for j in range(p):
                colsq[j] = np.dot(X[:, j], X[:, j])      # X.col(j).squaredNorm()
        if standardize_:
            colsq = colsq / (float(n) - 1.0)
            colsq = np.where(colsq == 0.0, 1.0, colsq)
            colsq_inv = 1.0 / np.sqrt(colsq)
        first = n // nslices
        G = np.zeros((p, p))
        for ff in range(nslices):
            Xs = X[ff * first:(n if ff + 1 == nslices else (ff + 1) * first)]
            G += xtx_fn(Xs) if xtx_fn else Xs.T @ Xs
    XY = np.zeros(q)
    XY[q - p:] = xty
    if intercept:
        XY[0] = Y.sum()
    if standardize_:
        XY[q - p:] *= colsq_inv
    XY /= n
    if standardize_:
        G = colsq_inv[:, None] * G * colsq_inv[None, :]
    XX = np.zeros((q, q))
    XX[q - p:, q - p:] = G
    if intercept:
        if standardize_:
            colsums = colsums * colsq_inv
        XX[0, 1:] = colsums
        XX[1:, 0] = colsums
        XX[0, 0] = n
    XX /= n
    d = top_eig(XX) * 1.005
    A = -XX
    A[np.diag_indices(q)] += d
    lmax = float(np.abs(XY).max())          # includes the intercept entry (src/oem_big.h:844-848)
    lams = _lambda_lists(lambda_, penalty, nlambda, lmin_ratio, lmax, alpha)
    grp = _Groups(groups, unique_groups, group_weights, scan=p)   # v < nvars quirk (src/oem_big.h:445)
    s = _Solver(q, pf, grp, o["maxit"], o["tol"])
    out = dict(beta=[], lambda_=[], niter=[], loss=[], d=d)
    for pp, pen in enumerate(penalty):
        lam = lams[pp]
        L = 1 if pen == "ols" else lam.size
        beta = np.zeros((p + 1, L), order="F")
        niter = np.zeros(L, dtype=np.int32)
        s.init(pen, alpha, _gamma_for(gamma, pp), tau)
        for i in range(L):
            niter[i] = s.solve(A, XY, d, lam[i])
            res = s.beta.copy()
            res[q - p:] *= colsq_inv if standardize_ else 1.0
            beta[1 - int(intercept):, i] = res
        out["beta"].append(beta)
        out["lambda_"].append(lam)
        out["niter"].append(niter)
        out["loss"].append(np.full(L, 1e99))
    return out


# ------------------------------------------------------------------------------------------
# oem_fit_sparse
# ------------------------------------------------------------------------------------------
def oem_fit_sparse(x, y, family, penalty, weights, groups, unique_groups, group_weights, lambda_,
                   nlambda, lmin_ratio, alpha, gamma, tau, penalty_factor, standardize_, intercept,
                   compute_loss, opts):
    """src/oem_sparse.cpp:30-264, src/oem_sparse.h:490-636,791-943: dgCMatrix X,
    unweighted; n > p, and n <= p without an intercept.  Like oem_fit_big it scales by the uncentred colsq/(n-1) and carries an
    explicit intercept column -- but that column is the CONSTANT intval = sqrt(mean(diag(X-block))/n)
    (:577-594), lambda_max skips the intercept entry (:851-862), and get_beta() multiplies the member
    beta(0) by intval IN PLACE (:895-900), so the next lambda warm-starts from the rescaled intercept and
    get_loss (:918-943, called after get_beta) sees the real one."""
    import scipy.sparse as sp
    o = _as_opts(opts)
    if family != "gaussian":
        raise ValueError("binomial not available for oem_fit_sparse, use oem_fit_logistic_sparse")
    if np.asarray(weights).size:
        raise ValueError("weights not implemented yet.")   # R/oem.R:244
    X = sp.csc_matrix(x, dtype=np.float64)
    Y = np.asarray(y, dtype=np.float64).ravel()
    n, p = X.shape
    wide = not n > p
    if wide and intercept:
        raise NotImplementedError("n <= p with an intercept: XY has min(n, p) entries and beta p + 1 in the reference "
                                  "(src/oem_sparse.h:782-784, 829-841), dimensionally inconsistent")
    q = p + int(intercept)
    pf = np.asarray(penalty_factor, dtype=np.float64).ravel()
    if intercept:
        pf = np.concatenate([[0.0], pf])
    colsq_inv = np.ones(p)
    if standardize_:
        colsq = np.asarray(X.multiply(X).sum(axis=0)).ravel() / (float(n) - 1.0)
        colsq = np.where(colsq == 0.0, 1.0, colsq)
        colsq_inv = 1.0 / np.sqrt(colsq)
    G = np.asarray((X.T @ X).todense())
    if standardize_:
        G = colsq_inv[:, None] * G * colsq_inv[None, :]
    XX = np.zeros((q, q))
    XX[q - p:, q - p:] = G
    intval = 1.0
    if intercept:
        xxdiag = float(np.diag(G).mean())
        intval = np.sqrt(xxdiag / n)
        colsums = np.asarray(X.sum(axis=0)).ravel() * intval
        if standardize_:
            colsums = colsums * colsq_inv
        XX[0, 1:] = colsums
        XX[1:, 0] = colsums
        XX[0, 0] = xxdiag
    XX /= n
    XY = np.zeros(q)
    XY[q - p:] = X.T @ Y
    if intercept:
        XY[0] = Y.sum() * intval
    if standardize_:
        XY[q - p:] *= colsq_inv
    XY /= n
    lmax = float(np.abs(XY[q - p:]).max())          # compute_lambda_zero: tail only
    if wide:
        # n <= p, no intercept (:609-616, 630-640): d from the n x n matrix XX'/n of the RAW X, and
        # next_u = X'(Y - X beta)/n + d beta never looks at XY or colsq_inv -- (dI - X'X/n) beta + X'y/n term by term.
        # `standardize` survives only in lambda_max (above, from the scaled XY) and in get_beta().
        d = top_eig(np.asarray((X @ X.T).todense()) / n) * 1.005
        XX = np.asarray((X.T @ X).todense()) / n
        XY = np.asarray(X.T @ Y).ravel() / n
    else:
        d = top_eig(XX) * 1.005
    A = -XX
    A[np.diag_indices(q)] += d
    lams = _lambda_lists(lambda_, penalty, nlambda, lmin_ratio, lmax, alpha)
    grp = _Groups(groups, unique_groups, group_weights, scan=q)      # scans groups.size() (:466)
    s = _Solver(q, pf, grp, o["maxit"], o["tol"])
    out = dict(beta=[], lambda_=[], niter=[], loss=[], d=d)
    for pp, pen in enumerate(penalty):
        lam = lams[pp]
        L = 1 if pen == "ols" else lam.size
        beta = np.zeros((p + 1, L), order="F")
        niter = np.zeros(L, dtype=np.int32)
        loss = np.full(L, 1e99)
        s.init(pen, alpha, _gamma_for(gamma, pp), tau)
        for i in range(L):
            niter[i] = s.solve(A, XY, d, lam[i])
            if intercept:
                s.beta[0] *= intval                  # in place: the warm start of the next lambda inherits it
            res = s.beta.copy()
            res[q - p:] *= colsq_inv if standardize_ else 1.0
            beta[1 - int(intercept):, i] = res
            if compute_loss:
                r = Y - X @ res[q - p:]
                if intercept:
                    r = r - res[0]
                loss[i] = float((r ** 2).sum())
        out["beta"].append(beta)
        out["lambda_"].append(lam)
        out["niter"].append(niter)
        out["loss"].append(loss)
    return out



# ------------------------------------------------------------------------------------------
# n <= p + intercept branches of the two logistic solvers (shared: they differ in three quirks only)
# ------------------------------------------------------------------------------------------
def logistic_wide_next_u(X, Y, W, beta_prev, d, colsq_inv, standardize_, intercept):
    """next_u of the n <= p + intercept branch, LITERALLY as the reference writes it
    (src/oem_logistic_dense.h:530-568, src/oem_logistic_sparse.h:534-580): the intercept variants use the unweighted
    residual Y - X beta, the no-intercept variants weight it by W.  Used by the tests to tie the Gram form iterated
    below -- (dI - G) beta + b -- to the reference's own expression."""
    n, p = X.shape
    cinv = colsq_inv if standardize_ else np.ones(p)
    if intercept:
        resid = Y - X @ (beta_prev[1:] * cinv)
        resid = resid - beta_prev[0]
        resid = resid / float(n)
        u = np.empty(p + 1)
        u[1:] = cinv * (X.T @ resid) + d * beta_prev[1:]
        u[0] = resid.sum() + d * beta_prev[0]
        return u
    return cinv * (X.T @ ((W * (Y - X @ (beta_prev * cinv))) / float(n))) + d * beta_prev


def logistic_wide_gram_form(X, Y, W, colsq_inv, standardize_, intercept):
    """The same iteration as (G, b) with u = (dI - G) beta + b: intercept variants G = [1, X D]'[1, X D] / n (unweighted,
    fixed for the whole fit), b = [sum y, D X'y] / n; no-intercept variants G = D X'WX D / n, b = D X'(W o y) / n."""
    n, p = X.shape
    cinv = colsq_inv if standardize_ else np.ones(p)
    Xs = X * cinv[None, :]
    if intercept:
        Xa = np.hstack([np.ones((n, 1)), Xs])
        return Xa.T @ Xa / float(n), Xa.T @ Y / float(n)
    return (Xs * W[:, None]).T @ Xs / float(n), Xs.T @ (W * Y) / float(n)


def _logistic_wide(X, Y, penalty, groups, unique_groups, group_weights, lambda_, nlambda, lmin_ratio, alpha, gamma, tau,
                   penalty_factor, standardize_, intercept, compute_loss, o, sparse_quirks):
    """n <= p + intercept (`nobs > nvars + int(intercept)` false): src/oem_logistic_dense.h:478-483, 530-568, 848-1036 and
    the sparse twin src/oem_logistic_sparse.h:503-509, 534-580, 851-1030.  What the reference does there is NOT a
    logistic fit: d = 1.0005 lambda_max((sqrt(W) X X' sqrt(W) + 1 [intercept]) / n) on the RAW X (the standardised
    variant is commented out), no gradient / XY update (the `nobs > nvars + intercept` guard at :965), and next_u
    (logistic_wide_next_u) is a least-squares step on the 0/1 response -- unweighted with an intercept, W-weighted
    without one.  Restated because the arithmetic is deterministic; iterated in Gram form (logistic_wide_gram_form).
    Quirks of the sparse solver: compute_XtX_d_update_A() on every data pass (no hessian.type test, :958-961);
    get_beta() multiplies beta(0) by `intval`, which this branch never sets (0), whenever nobs > nvars, i.e. for
    n = p + 1 with an intercept (:1040-1043)."""
    n, p = X.shape
    q = p + int(intercept)
    pf = np.asarray(penalty_factor, dtype=np.float64).ravel()
    if intercept:
        pf = np.concatenate([[0.0], pf])
    colsq_inv = np.ones(p)
    if standardize_:
        colsq = (X ** 2).sum(axis=0) / (float(n) - 1.0)
        colsq = np.where(colsq == 0.0, 1.0, colsq)
        colsq_inv = 1.0 / np.sqrt(colsq)
    XY0 = np.zeros(q)
    XY0[q - p:] = X.T @ Y
    if intercept:
        XY0[0] = Y.sum()
    if standardize_:
        XY0[q - p:] *= colsq_inv
    XY0 /= n
    lmax = float(np.abs(XY0[q - p:]).max())
    lams = _lambda_lists(lambda_, penalty, nlambda, lmin_ratio, lmax, alpha,
                         logistic_fudge=True, gamma=_gamma_for(gamma, 0))
    grp = _Groups(groups, unique_groups, group_weights, scan=q, zero_weight_for_group0=True)
    s = _Solver(q, pf, grp, o["maxit"], o["tol"])
    every_pass = sparse_quirks or o["hessian_type"] == "full"
    G, b, d = None, None, None
    W = np.zeros(n)
    prob = np.zeros(n)
    out = dict(beta=[], lambda_=[], niter=[], loss=[], d=None)
    for pp, pen in enumerate(penalty):
        lam = lams[pp]
        L = 1 if pen == "ols" else lam.size
        beta = np.zeros((p + 1, L), order="F")
        niter = np.zeros(L, dtype=np.int32)
        loss = np.full(L, 1e99)
        s.init(pen, alpha, _gamma_for(gamma, pp), tau)
        for i in range(L):
            on_lam_1 = (i == 0)
            it = 0
            for it in range(o["irls_maxit"]):
                beta_prev_irls = s.beta.copy()
                if not (it == 0 and not on_lam_1):
                    bx = s.beta[q - p:] * colsq_inv if standardize_ else s.beta[q - p:]
                    eta = X @ bx + (s.beta[0] if intercept else 0.0)
                    prob = 1.0 / (1.0 + np.exp(-eta))
                    W = prob * (1.0 - prob)
                    if it < n and W[it] < 1e-5:         # sic: indexed by the IRLS counter
                        W[it] = 1e-5
                    if (it == 0 and on_lam_1) or every_pass:
                        sw = np.sqrt(W)
                        M = (X * sw[:, None]) @ (X * sw[:, None]).T          # XWXt(): n x n, raw X
                        if intercept:
                            M = M + 1.0
                        d = top_eig(M / float(n)) * 1.0005
                    if not intercept or G is None:
                        G, b = logistic_wide_gram_form(X, Y, W, colsq_inv, standardize_, intercept)
                A = -G
                A[np.diag_indices(q)] += d
                s.solve(A, b, d, lam[i])
                if stop_rule(s.beta, beta_prev_irls, o["irls_tol"]):
                    break
            else:
                it = o["irls_maxit"]
            niter[i] = it + 1
            if sparse_quirks and intercept and n > p:
                s.beta[0] *= 0.0                        # get_beta(): beta(0) *= intval with intval never set
            res = s.beta.copy()
            if standardize_:
                res[q - p:] *= colsq_inv
            beta[1 - int(intercept):, i] = res
            if compute_loss:
                ok = np.where(Y == 1, prob > 1e-5, prob <= 1.0 - 1e-5)
                pr = np.where(Y == 1, prob, 1.0 - prob)
                loss[i] = float(np.where(ok, np.log(1.0 / np.where(ok, pr, 1.0)), np.log(1.0 / 1e-5)).sum())
        out["beta"].append(beta)
        out["lambda_"].append(lam)
        out["niter"].append(niter)
        out["loss"].append(loss)
    out["d"] = d
    return out


# ------------------------------------------------------------------------------------------
# oem_fit_logistic_dense
# ------------------------------------------------------------------------------------------
def oem_fit_logistic_dense(x, y, family, penalty, weights, groups, unique_groups, group_weights,
                           lambda_, nlambda, lmin_ratio, alpha, gamma, tau, penalty_factor,
                           standardize_, intercept, compute_loss, opts):
    """src/oem_logistic_dense.cpp:29-313, src/oem_logistic_dense.h:721-1036,
    unweighted, ncores=1, including the quirks of Appendix B item 5 (n <= p + intercept: _logistic_wide)."""
    o = _as_opts(opts)
    if np.asarray(weights).size:
        raise ValueError("weights not implemented yet.")
    X = np.asarray(x, dtype=np.float64)
    Y = np.asarray(y, dtype=np.float64).ravel()
    n, p = X.shape
    q = p + int(intercept)
    if not n > q:
        return _logistic_wide(X, Y, penalty, groups, unique_groups, group_weights, lambda_, nlambda, lmin_ratio, alpha,
                              gamma, tau, penalty_factor, standardize_, intercept, compute_loss, o, sparse_quirks=False)
    pf = np.asarray(penalty_factor, dtype=np.float64).ravel()
    if intercept:
        pf = np.concatenate([[0.0], pf])
    colsq_inv = np.ones(p)
    if standardize_:
        colsq = (X ** 2).sum(axis=0) / (float(n) - 1.0)
        colsq = np.where(colsq == 0.0, 1.0, colsq)
        colsq_inv = 1.0 / np.sqrt(colsq)
    XY = np.zeros(q)
    XY[q - p:] = X.T @ Y
    if intercept:
        XY[0] = Y.sum()
    if standardize_:
        XY[q - p:] *= colsq_inv
    XY /= n
    lmax = float(np.abs(XY[q - p:]).max())
    lams = _lambda_lists(lambda_, penalty, nlambda, lmin_ratio, lmax, alpha,
                         logistic_fudge=True, gamma=_gamma_for(gamma, 0))
    grp = _Groups(groups, unique_groups, group_weights, scan=q, zero_weight_for_group0=True)
    s = _Solver(q, pf, grp, o["maxit"], o["tol"])
    full_hessian = o["hessian_type"] == "full"
    XX, A, d = None, None, None
    prob = np.zeros(n)
    out = dict(beta=[], lambda_=[], niter=[], loss=[], d=None)
    for pp, pen in enumerate(penalty):
        lam = lams[pp]
        L = 1 if pen == "ols" else lam.size
        beta = np.zeros((p + 1, L), order="F")
        niter = np.zeros(L, dtype=np.int32)
        loss = np.full(L, 1e99)
        s.init(pen, alpha, _gamma_for(gamma, pp), tau)
        for i in range(L):
            on_lam_1 = (i == 0)
            it = 0
            for it in range(o["irls_maxit"]):
                beta_prev_irls = s.beta.copy()
                if not (it == 0 and not on_lam_1):
                    bx = s.beta[q - p:] * colsq_inv if standardize_ else s.beta[q - p:]
                    eta = X @ bx + (s.beta[0] if intercept else 0.0)
                    prob = 1.0 / (1.0 + np.exp(-eta))
                    W = prob * (1.0 - prob)
                    if W[it] < 1e-5:                    # sic: indexed by the IRLS counter
                        W[it] = 1e-5
                    if (it == 0 and on_lam_1) or full_hessian:
                        XtWX = (X * W[:, None]).T @ X
                        XX = np.zeros((q, q))
                        if standardize_:
                            XtWX = colsq_inv[:, None] * XtWX * colsq_inv[None, :]
                        XX[q - p:, q - p:] = XtWX
                        if intercept:
                            cs = W @ X
                            if standardize_:
                                cs = cs * colsq_inv
                            XX[0, 1:] = cs
                            XX[1:, 0] = cs
                            XX[0, 0] = W.sum()
                        XX /= n
                        d = top_eig(XX) * 1.0005
                        A = -XX
                        A[np.diag_indices(q)] += d
                    presid = Y - prob
                    grad = np.zeros(q)
                    grad[q - p:] = (X.T @ presid) / float(n)
                    if intercept:
                        grad[0] = presid.sum() / float(n)
                    if standardize_:
                        grad[q - p:] *= colsq_inv
                    XY = XX @ s.beta + grad
                s.solve(A, XY, d, lam[i])
                if stop_rule(s.beta, beta_prev_irls, o["irls_tol"]):
                    break
            else:
                it = o["irls_maxit"]
            niter[i] = it + 1
            res = s.beta.copy()
            if standardize_:
                res[q - p:] *= colsq_inv
            beta[1 - int(intercept):, i] = res
            if compute_loss:        # get_loss(): src/oem_logistic_dense.h:1057-1088 (stale prob)
                ok = np.where(Y == 1, prob > 1e-5, prob <= 1.0 - 1e-5)
                pr = np.where(Y == 1, prob, 1.0 - prob)
                loss[i] = float(np.where(ok, np.log(1.0 / np.where(ok, pr, 1.0)), np.log(1.0 / 1e-5)).sum())
        out["beta"].append(beta)
        out["lambda_"].append(lam)
        out["niter"].append(niter)
        out["loss"].append(loss)
    out["d"] = d
    return out


# ------------------------------------------------------------------------------------------
# oem_fit_logistic_sparse
# ------------------------------------------------------------------------------------------
def oem_fit_logistic_sparse(x, y, family, penalty, weights, groups, unique_groups, group_weights,
                            lambda_, nlambda, lmin_ratio, alpha, gamma, tau, penalty_factor,
                            standardize_, intercept, compute_loss, opts):
    """src/oem_logistic_sparse.cpp:30-330, src/oem_logistic_sparse.h:458-545 (compute_XtX_d_update_A), 727-1100
    (init_oem / solve / get_beta / get_loss): dgCMatrix X, n > p branch, unweighted, the `ncores <= 1` code path.

    Differences from the dense logistic solver that are restated here because they change results:
      * compute_XtX_d_update_A() is called on EVERY IRLS data pass (no hessian.type test, :958-961), so X'WX and d are
        rebuilt each time;
      * the intercept column is the constant `intval`, fixed at the first pass: xxdiag = mean(diag(X-block)) of the
        first X'WX, intval = sqrt((xxdiag / sum W) / n); XX(0,0) = xxdiag, border = (X'W o colsq_inv) * intval (:470-491);
      * the linear predictor of the (standardize, intercept) branch adds beta(0) WITHOUT intval (:875-876) and
        grad(0) = sum(y - prob) / n is not scaled either (:972-973);
      * get_beta() multiplies the member beta(0) by intval in place (:1040-1043), so the next lambda warm-starts from it;
      * lambda_max skips the intercept entry (:808-818).
    intercept && !standardize multiplies beta by a colsq_inv that was never initialised (:880 vs :737-751): undefined in
    the reference, rejected here."""
    import scipy.sparse as sp
    o = _as_opts(opts)
    if np.asarray(weights).size:
        raise ValueError("weights not implemented yet.")
    if intercept and not standardize_:
        raise NotImplementedError("oem_fit_logistic_sparse with intercept = TRUE, standardize = FALSE reads an "
                                  "uninitialised vector in the reference (src/oem_logistic_sparse.h:880)")
    X = sp.csc_matrix(x, dtype=np.float64)
    Xr = sp.csr_matrix(X)
    Y = np.asarray(y, dtype=np.float64).ravel()
    n, p = X.shape
    q = p + int(intercept)
    if not n > q:
        return _logistic_wide(np.asarray(X.todense()), Y, penalty, groups, unique_groups, group_weights, lambda_, nlambda,
                              lmin_ratio, alpha, gamma, tau, penalty_factor, standardize_, intercept, compute_loss, o,
                              sparse_quirks=True)
    pf = np.asarray(penalty_factor, dtype=np.float64).ravel()
    if intercept:
        pf = np.concatenate([[0.0], pf])
    colsq_inv = np.ones(p)
    if standardize_:
        colsq = np.asarray(X.multiply(X).sum(axis=0)).ravel() / (float(n) - 1.0)
        colsq = np.where(colsq == 0.0, 1.0, colsq)
        colsq_inv = 1.0 / np.sqrt(colsq)
    XY = np.zeros(q)
    XY[q - p:] = X.T @ Y                  # XY(0) = sum(Y) * intval with intval still 0 (:768): irrelevant, XY is rebuilt
    if standardize_:
        XY[q - p:] *= colsq_inv
    XY /= n
    lmax = float(np.abs(XY[q - p:]).max())
    lams = _lambda_lists(lambda_, penalty, nlambda, lmin_ratio, lmax, alpha,
                         logistic_fudge=True, gamma=_gamma_for(gamma, 0))
    grp = _Groups(groups, unique_groups, group_weights, scan=q, zero_weight_for_group0=True)
    s = _Solver(q, pf, grp, o["maxit"], o["tol"])
    XX, A, d = None, None, None
    xxdiag, intval = 0.0, 0.0
    prob = np.zeros(n)
    out = dict(beta=[], lambda_=[], niter=[], loss=[], d=None)
    for pp, pen in enumerate(penalty):
        lam = lams[pp]
        L = 1 if pen == "ols" else lam.size
        beta = np.zeros((p + 1, L), order="F")
        niter = np.zeros(L, dtype=np.int32)
        loss = np.full(L, 1e99)
        s.init(pen, alpha, _gamma_for(gamma, pp), tau)
        for i in range(L):
            on_lam_1 = (i == 0)
            it = 0
            for it in range(o["irls_maxit"]):
                beta_prev_irls = s.beta.copy()
                if not (it == 0 and not on_lam_1):
                    bx = s.beta[q - p:] * colsq_inv if standardize_ else s.beta[q - p:]
                    eta = Xr @ bx + (s.beta[0] if intercept else 0.0)
                    prob = 1.0 / (1.0 + np.exp(-eta))
                    W = prob * (1.0 - prob)
                    if W[it] < 1e-5:                    # sic: indexed by the IRLS counter (:948-954)
                        W[it] = 1e-5
                    XtWX = np.asarray((X.T @ sp.diags(W) @ X).todense())
                    if standardize_:
                        XtWX = colsq_inv[:, None] * XtWX * colsq_inv[None, :]
                    XX = np.zeros((q, q))
                    XX[q - p:, q - p:] = XtWX
                    if intercept:
                        cs = np.asarray(X.T @ W).ravel() * colsq_inv
                        if xxdiag <= 0:
                            xxdiag = float(np.diag(XtWX).mean())
                            intval = np.sqrt((xxdiag / W.sum()) / float(n))
                        cs = cs * intval
                        XX[0, 1:] = cs
                        XX[1:, 0] = cs
                        XX[0, 0] = xxdiag
                    XX /= n
                    d = top_eig(XX) * 1.0005
                    A = -XX
                    A[np.diag_indices(q)] += d
                    presid = Y - prob
                    grad = np.zeros(q)
                    grad[q - p:] = np.asarray(X.T @ presid).ravel() / float(n)
                    if intercept:
                        grad[0] = presid.sum() / float(n)
                    if standardize_:
                        grad[q - p:] *= colsq_inv
                    XY = XX @ s.beta + grad
                s.solve(A, XY, d, lam[i])
                if stop_rule(s.beta, beta_prev_irls, o["irls_tol"]):
                    break
            else:
                it = o["irls_maxit"]
            niter[i] = it + 1
            if intercept:
                s.beta[0] *= intval                     # get_beta(): in place (:1040-1043)
            res = s.beta.copy()
            if standardize_:
                res[q - p:] *= colsq_inv
            beta[1 - int(intercept):, i] = res
            if compute_loss:
                ok = np.where(Y == 1, prob > 1e-5, prob <= 1.0 - 1e-5)
                pr = np.where(Y == 1, prob, 1.0 - prob)
                loss[i] = float(np.where(ok, np.log(1.0 / np.where(ok, pr, 1.0)), np.log(1.0 / 1e-5)).sum())
        out["beta"].append(beta)
        out["lambda_"].append(lam)
        out["niter"].append(niter)
        out["loss"].append(loss)
    out["d"] = d
    return out
