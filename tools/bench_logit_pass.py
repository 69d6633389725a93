"""Logistic IRLS data pass at BASELINE configs[3] size (n = 2e6 x p = 1000, 16 GB) on one B200: the fused single-sweep
slab kernel (logit_slab.cu) against the two HBM sweeps (xb_kernel + colstats_kernel).  One JSON line."""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oem_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2_000_000)
    ap.add_argument("--p", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    n, p = a.n, a.p
    L = api.load()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(104)
    Xt = torch.empty((p, n), dtype=torch.float64, device=dev)
    for j in range(0, p, 50):
        Xt[j:j + 50].normal_(generator=g)
    X = Xt.t()
    b = torch.randn(p, generator=g, dtype=torch.float64, device=dev) * 0.05
    y = (torch.rand(n, generator=g, dtype=torch.float64, device=dev) < 0.5).double()
    prob = torch.empty(n, dtype=torch.float64, device=dev)
    w = torch.empty_like(prob)
    res = torch.empty_like(prob)
    grad = torch.empty(p + 1, dtype=torch.float64, device=dev)
    out3 = torch.empty(3 * p, dtype=torch.float64, device=dev)
    ms, msr = ctypes.c_double(), ctypes.c_double()
    api._check(L.oemb200_logit_slab_pass(X.data_ptr(), n, p, n, b.data_ptr(), 0.1, y.data_ptr(), prob.data_ptr(), w.data_ptr(),
                                         grad.data_ptr(), 3, None, ctypes.byref(ms), ctypes.byref(msr)))
    api._check(L.oemb200_logit_slab_pass(X.data_ptr(), n, p, n, b.data_ptr(), 0.1, y.data_ptr(), prob.data_ptr(), w.data_ptr(),
                                         grad.data_ptr(), a.reps, None, ctypes.byref(ms), ctypes.byref(msr)))
    t_xb, t_cs = [], []
    for _ in range(5):
        m1, m2 = ctypes.c_double(), ctypes.c_double()
        api._check(L.oemb200_xb_logistic(X.data_ptr(), n, p, n, b.data_ptr(), 0.1, y.data_ptr(), prob.data_ptr(), res.data_ptr(),
                                         w.data_ptr(), None, ctypes.byref(m1)))
        api._check(L.oemb200_colstats(X.data_ptr(), n, p, n, res.data_ptr(), None, out3.data_ptr(), None, ctypes.byref(m2)))
        t_xb.append(m1.value); t_cs.append(m2.value)
    err = float((grad[1:] - out3[:p]).abs().max())
    gb = 8.0 * n * p / 1e9
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    print(json.dumps({"n": n, "p": p, "x_gb": gb, "slab_pass_ms": ms.value, "slab_relayout_ms": msr.value,
                      "slab_gbs_algorithmic": gb / (ms.value / 1e3), "two_sweeps_ms": min(t_xb) + min(t_cs),
                      "xb_ms": min(t_xb), "colstats_ms": min(t_cs), "hbm_peak_gbs": peak,
                      "slab_frac_of_peak": gb / (ms.value / 1e3) / peak if peak else None,
                      "max_abs_grad_diff_vs_two_sweeps": err}))


if __name__ == "__main__":
    main()
