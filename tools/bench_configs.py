"""Secondary measurements: BASELINE.json configs[0..3] at full size on one B200, device-resident inputs,
one JSON line per config with the library's own CUDA-event phase timings and the roofline figures of the
dominant kernels (Gram TFLOP/s, logistic GEMV GB/s, CV-scoring TFLOP/s).  Not the headline bench (bench.py)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oem_b200  # noqa: E402
from oem_b200 import api  # noqa: E402

dev = torch.device("cuda", 0)


def gen(n, p, seed, sd=1.0, coef=None, noise=1.0, binomial=False):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    ld = n + (n & 1)
    Xt = torch.empty((p, ld), dtype=torch.float64, device=dev)
    b = torch.zeros(p, dtype=torch.float64, device=dev)
    b[:len(coef)] = torch.tensor(coef, dtype=torch.float64, device=dev)
    eta = torch.zeros(n, dtype=torch.float64, device=dev)
    step = max(1, min(p, (1 << 28) // ld))
    for j in range(0, p, step):
        blk = Xt[j:j + step]
        blk.normal_(0.0, sd, generator=g)
        eta += blk[:, :n].t() @ b[j:j + step]
    if binomial:
        y = (torch.rand(n, generator=g, dtype=torch.float64, device=dev) < torch.sigmoid(eta)).double()
    else:
        y = eta + noise * torch.randn(n, generator=g, dtype=torch.float64, device=dev)
    return Xt.t()[:n], y


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    outs, ts = [], []
    for _ in range(reps):
        t0 = time.perf_counter()
        outs.append(fn())
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts), outs[-1]


def line(name, wall, out, extra):
    st = out["stats"]
    d = {"config": name, "wall_s": wall, "phases_ms": {k: round(v, 3) for k, v in st.items() if k.startswith("ms_")},
         "oem_iterations": st["total_oem_iters"], "kernel_launches": st["kernel_launches"]}
    if st["gram_launches"]:
        d["gram_tflops"] = st["gram_flops"] / (max(st["ms_gram"], 1e-9) / 1e3) / 1e12
    d.update(extra(st) if extra else {})
    print(json.dumps(d), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--scale", type=float, default=1.0, help="scale n (debug)")
    ap.add_argument("--nlambda", type=int, default=100, help="lambdas per penalty for the xval config (experiments on the CV GEMM's column count)")
    a = ap.parse_args()
    cfgs = [int(c) for c in a.configs.split(",")]
    opts = dict(maxit=500, tol=1e-7)
    if 1 in cfgs:      # README.md:45-73: published 1.617 s (CPU, single thread, unstated hardware)
        n, p = int(1e6 * a.scale), 100
        rng = np.random.default_rng(101)
        X, y = gen(n, p, 101, sd=3.0, coef=list(rng.uniform(0, 1, 25)))
        args = [X, y, "gaussian", ["elastic.net"], [], [], [], [], [], 100, 1e-4, 1.0, 3.0, 0.5, np.ones(p), False, True, False,
                dict(opts, tol=1e-10)]
        w, out = timed(lambda: oem_b200.oem_fit_dense(*args), a.reps)
        line("configs[0] README lasso n=1e6 p=100 (published CPU 1.617 s)", w, out, None)
        del X, y
    if 2 in cfgs:      # README.md:100-157: published MCP 105.8 ms, SCAD 78.8 ms (separate calls)
        n, p = 5000, 200
        rng = np.random.default_rng(102)
        X, y = gen(n, p, 102, sd=3.0, coef=list(rng.uniform(-0.5, 0.5, 25)))
        args = [X, y, "gaussian", ["mcp", "scad"], [], [], [], [], [], 200, 1e-4, 1.0, [2.0, 4.0], 0.5, np.ones(p), True, True,
                False, dict(opts, tol=1e-10)]
        w, out = timed(lambda: oem_b200.oem_fit_dense(*args), a.reps)
        line("configs[1] README MCP g=2 + SCAD g=4 n=5000 p=200 L=200 batched (published CPU 105.8 + 78.8 ms)", w, out, None)
        del X, y
    if 3 in cfgs:
        n, p, F = int(1e7 * a.scale), 500, 10
        X, y = gen(n, p, 103, coef=[.5, .5, -.5, -.5, 1.0], noise=4.0)
        rng = np.random.default_rng(103)
        foldid = (1 + rng.permutation(n) % F).astype(np.int32)
        groups = np.concatenate([[0], np.repeat(np.arange(1, 51), 10)])
        args = [X, y, "gaussian", ["lasso", "grp.lasso", "mcp"], [], groups, np.unique(groups), [], [], a.nlambda, 1e-4, 1.0, 3.0, 0.5,
                np.ones(p), True, True, F, foldid, False, "mse", dict(opts)]
        w, out = timed(lambda: oem_b200.oem_xval_dense(*args), a.reps)
        line("configs[2] xval.oem 10-fold lasso+grp.lasso+mcp n=1e7 p=500", w, out,
             lambda st: {"cvscore_tflops": 2.0 * n * p * 3 * a.nlambda / (st["ms_cvscore"] / 1e3) / 1e12,
                         "cvm_min_lasso": float(np.min(out["cvm"][0]))})
        del X, y
    if 4 in cfgs:
        n, p = int(2e6 * a.scale), 1000
        X, y = gen(n, p, 104, coef=[.15, .15, -.15, -.15, .25], binomial=True)
        args = [X, y, "binomial", ["lasso"], [], [], [], [], [], 100, 1e-4, 1.0, 3.0, 0.5, np.ones(p), True, True, False, dict(opts)]
        w, out = timed(lambda: oem_b200.oem_fit_logistic_dense(*args), max(1, a.reps - 1))
        line("configs[3] logistic lasso n=2e6 p=1000", w, out,
             lambda st: {"data_pass_gbs": st["gemv_bytes"] / (max(st["ms_irls_xb"] + st["ms_irls_xtr"], 1e-9) / 1e3) / 1e9,
                         "ms_per_irls_data_pass": (st["ms_irls_xb"] + st["ms_irls_xtr"]) / max(1, st["xb_launches"]),
                         "irls_iterations": int(np.sum(out["niter"][0])), "xb_launches": st["xb_launches"]})
        del X, y
    if 5 in cfgs:      # next row: predict.oem on device, newx 4e6 x 500, 100 lambdas, output stays on the device
        n, p, L = int(4e6 * a.scale), 500, 100
        X, _ = gen(n, p, 106, coef=[.5])
        rng = np.random.default_rng(106)
        B = rng.normal(size=(p + 1, L)) * (rng.uniform(size=(p + 1, L)) < 0.1)
        out_t = torch.empty((L, n), dtype=torch.float64, device=dev).t()
        w, (_, st) = timed(lambda: api.predict_matrix(X, B, out=out_t, return_stats=True), a.reps)
        print(json.dumps({"config": "predict.oem newx n=4e6 p=500 L=100 (device in, device out)", "wall_s": w,
                          "ms_gemm": st["ms_cvscore"], "gemm_tflops": 2.0 * n * p * L / (st["ms_cvscore"] / 1e3) / 1e12,
                          "out_gb": n * L * 8 / 1e9}), flush=True)
        del X, out_t
    if 6 in cfgs:      # vignettes/oem_vignette.html section 5.1.2: the reference's own logistic timings (n = 5e4, p = 100, 100 lambdas):
        # 2.64 s for grp.lasso alone, 10.83 s for five penalties in one call (CPU, unstated hardware)
        n, p = 50000, 100
        X, y = gen(n, p, 107, coef=[.15, .15, -.15, -.15, .25], binomial=True)
        groups = np.concatenate([[0], np.repeat(np.arange(1, 21), 5)])
        for pens, pub in ((["grp.lasso"], "2.64 s"), (["grp.lasso", "lasso", "mcp", "scad", "elastic.net"], "10.83 s")):
            args = [X, y, "binomial", pens, [], groups, np.unique(groups), [], [], 100, 1e-4, 1.0, 3.0, 0.5, np.ones(p), True, True,
                    False, dict(opts)]
            w, out = timed(lambda: oem_b200.oem_fit_logistic_dense(*args), a.reps)
            line(f"vignette 5.1.2 logistic n=5e4 p=100 {'+'.join(pens)} 100 lambdas (published CPU {pub})", w, out,
                 lambda st: {"irls_iterations": int(sum(np.sum(nn) for nn in out["niter"]))})
        del X, y
    torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
