"""An INDEPENDENT second implementation of the binomial lasso, used only as a checker for the (reference-unpinned)
logistic entry: plain cyclic coordinate descent on

    f(b0, b) = (1/n) sum_i [ log(1 + exp(eta_i)) - y_i eta_i ] + lambda * sum_j pf_j |b_j|,   eta = b0 + X b,

the problem the reference's vignette checks oem(family = "binomial") against glmnet on (README.md:53-92 does the same
for the gaussian lasso: agreement with glmnet at the 1e-7 level).  Nothing is shared with the oracle or the CUDA path:
no Gram matrix, no OEM majorisation with d = 1.0005 * lambda_max, no IRLS outer loop -- every coordinate takes a
soft-thresholded Newton-bound step (second derivative <= sum x_ij^2 / 4n), eta is updated incrementally, and the loop
runs until the largest coefficient change of a full sweep is below `tol`."""
import numpy as np


def binomial_lasso_cd(X, y, lam, pf=None, b0=0.0, b=None, tol=1e-11, max_sweeps=200000):
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n, p = X.shape
    pf = np.ones(p) if pf is None else np.asarray(pf, dtype=np.float64)
    b = np.zeros(p) if b is None else np.array(b, dtype=np.float64)
    eta = b0 + X @ b
    h = (X * X).sum(axis=0) / (4.0 * n)            # curvature bound of every coordinate
    for sweep in range(max_sweeps):
        delta = 0.0
        # intercept: unpenalised Newton-bound step (curvature bound 1/4)
        pr = 1.0 / (1.0 + np.exp(-eta))
        s = 4.0 * np.mean(y - pr)
        b0 += s
        eta += s
        delta = max(delta, abs(s))
        for j in range(p):
            if h[j] == 0.0:
                continue
            pr = 1.0 / (1.0 + np.exp(-eta))
            g = X[:, j] @ (y - pr) / n
            u = b[j] + g / h[j]
            t = lam * pf[j] / h[j]
            new = np.sign(u) * max(abs(u) - t, 0.0)
            if new != b[j]:
                eta += X[:, j] * (new - b[j])
                delta = max(delta, abs(new - b[j]))
                b[j] = new
        if delta < tol:
            return b0, b, sweep + 1
    raise RuntimeError("coordinate descent did not converge")
