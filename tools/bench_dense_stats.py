"""A/B of oem_fit_dense's column statistics: fused into the Gram launch (default for p >= 256, n * p >= 2^28) against
the separate HBM sweep (OEMB200_SEPARATE_COLSTATS=1).  Device-resident inputs; prints the library's phase timings."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import oem_b200  # noqa: E402
from bench_configs import gen  # noqa: E402

n, p = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 512
X, y = gen(n, p, 77, coef=[0.5, 0.5, -0.5, -0.5, 1.0])
args = [X, y, "gaussian", ["lasso", "mcp"], [], [], [], [], [], 100, 1e-4, 1.0, 3.0, 0.5, np.ones(p), True, True, False,
        dict(maxit=500, tol=1e-7)]
for mode in ("fused", "separate"):
    if mode == "separate":
        os.environ["OEMB200_SEPARATE_COLSTATS"] = "1"
    best, out = None, None
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = oem_b200.oem_fit_dense(*args)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    st = out["stats"]
    print(json.dumps({"config": f"oem_fit_dense n={n} p={p} standardize+intercept, column statistics {mode}", "wall_s": best,
                      "phases_ms": {k: round(v, 3) for k, v in st.items() if k.startswith("ms_")},
                      "gram_tflops": st["gram_flops"] / (st["ms_gram"] / 1e3) / 1e12}))
