"""Generates tests/golden/golden_oracle.json: small regression fixtures of the five entry points.

The reference (R + RcppEigen) cannot be built or run in this environment and ships no tests or golden
vectors, so these fixtures are produced by the CPU oracle itself on seeded inputs:
they pin the oracle (and, through tests/test_gpu_entries.py::test_golden_on_gpu, the CUDA path) against
silent drift, they do NOT pin it to the reference -- "parity unpinned".
Run:  python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cases import args_xy, binomial_problem, gaussian_problem   # noqa: E402

CASES = ["dense_c1", "dense_c2", "big_c5", "logistic_c4", "xtx", "xval_c3", "xval_weighted"]


def inputs(name):
    """-> (entry name, positional args)"""
    if name == "dense_c1":      # README lasso shape: elastic.net alpha=1, intercept, no standardize
        X, y = gaussian_problem(101, 4000, 50, sd_x=3.0)
        return "oem_fit_dense", args_xy(X, y, "gaussian", ["elastic.net"], standardize=False, nlambda=40, opts=dict(tol=1e-10))
    if name == "dense_c2":      # MCP gamma=2 + SCAD gamma=4 batched
        X, y = gaussian_problem(102, 2000, 60, sd_x=3.0)
        return "oem_fit_dense", args_xy(X, y, "gaussian", ["mcp", "scad"], gamma=[2.0, 4.0], nlambda=50, opts=dict(tol=1e-10))
    if name == "big_c5":
        X, y = gaussian_problem(105, 5000, 80)
        return "oem_fit_big", args_xy(X, y, "gaussian", ["lasso", "scad", "mcp"], gamma=[3.0, 3.7, 3.0], nlambda=40)
    if name == "logistic_c4":
        X, y = binomial_problem(104, 3000, 30)
        return "oem_fit_logistic_dense", args_xy(X, y, "binomial", ["lasso"], nlambda=12, lmin_ratio=1e-2)
    if name == "xtx":
        X, y = gaussian_problem(3, 2000, 25)
        n = X.shape[0]
        return "oem_xtx", [X.T @ X / n, X.T @ y / n, "gaussian", ["lasso", "mcp"], [], [], [], [], 30, 1e-3, 1.0, 3.0, 0.5,
                           [], np.ones(25), dict(maxit=500, tol=1e-10)]
    if name == "xval_c3":
        X, y = gaussian_problem(103, 3000, 40, coef="vignette", noise=4.0)
        rng = np.random.default_rng(103)
        foldid = 1 + rng.permutation(3000) % 5
        groups = np.concatenate([[0], np.repeat(np.arange(1, 9), 5)])
        a = args_xy(X, y, "gaussian", ["lasso", "grp.lasso", "mcp"], groups=groups, unique_groups=np.unique(groups), nlambda=25)
        return "oem_xval_dense", a[:17] + [5, foldid, False, "mse", a[18]]
    if name == "xval_weighted":     # xval.oem(weights=): X'WX fold Grams, weighted mae score
        X, y = gaussian_problem(77, 2000, 20, noise=2.0)
        rng = np.random.default_rng(77)
        foldid = 1 + rng.permutation(2000) % 4
        a = args_xy(X, y, "gaussian", ["lasso", "scad"], nlambda=20)
        a[4] = rng.uniform(0.25, 3.0, size=2000)
        return "oem_xval_dense", a[:17] + [4, foldid, False, "mae", a[18]]
    raise KeyError(name)


def summarize(out):
    w = None
    s = dict(d=float(out["d"]), beta_checksum=[], lambda_first_last=[], niter_sum=[])
    for pp in range(len(out["beta"])):
        B = np.asarray(out["beta"][pp])
        if w is None or w.shape != B.shape:
            w = np.cos(np.arange(B.size)).reshape(B.shape)        # fixed weights: a checksum sensitive to every entry
        s["beta_checksum"].append([float(np.sum(B * w)), float(np.sum(np.abs(B))), float(np.max(np.abs(B)))])
        lam = np.asarray(out["lambda_"][pp])
        s["lambda_first_last"].append([float(lam[0]), float(lam[-1])])
        s["niter_sum"].append(int(np.sum(out["niter"][pp])))
    if "cvm" in out:
        s["cvm_checksum"] = [float(np.sum(c)) for c in out["cvm"]]
        s["cvsd_checksum"] = [float(np.sum(c)) for c in out["cvsd"]]
    return s


def run_case(impl, name):
    fn, a = inputs(name)
    return summarize(getattr(impl, fn)(*a))


if __name__ == "__main__":
    from oracle import oracle as orc
    orc.build()
    gold = {name: run_case(orc, name) for name in CASES}
    with open(os.path.join(HERE, "golden_oracle.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", len(gold), "cases")
