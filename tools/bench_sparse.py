"""Secondary measurement: oem_fit_sparse on an rsparsematrix-like design (man/oem.Rd:104-112 scaled
up), device-resident dgCMatrix slots; prints the library's CUDA-event phase timings.  Not the headline bench."""
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

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2_000_000)
ap.add_argument("--p", type=int, default=1000)
ap.add_argument("--density", type=float, default=0.01)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(7)
per_col = max(1, int(a.n * a.density))
# every column draws `per_col` distinct rows (sorted), like a uniform-density dgCMatrix
rows = torch.empty((a.p, per_col), dtype=torch.int32, device=dev)
for j in range(a.p):
    rows[j] = torch.randperm(a.n, generator=g, device=dev)[:per_col].sort().values.int()
col_ptr = (torch.arange(a.p + 1, device=dev) * per_col).int()
vals = torch.randn(a.p * per_col, generator=g, dtype=torch.float64, device=dev)
beta = torch.zeros(a.p, dtype=torch.float64, device=dev)
beta[:25] = torch.rand(25, generator=g, dtype=torch.float64, device=dev) - 0.5
y = torch.randn(a.n, generator=g, dtype=torch.float64, device=dev)
cols = torch.arange(a.p, device=dev).repeat_interleave(per_col)
y.index_add_(0, rows.view(-1).long(), vals * beta[cols])
slots = (rows.view(-1), col_ptr, vals, (a.n, a.p))
args = [slots, y, "gaussian", ["lasso", "scad", "mcp"], [], [], [], [], [], 100, 1e-4, 1.0, [3.0, 3.7, 3.0], 0.5, np.ones(a.p),
        True, True, False, dict(maxit=500, tol=1e-7)]
best = None
for _ in range(a.reps + 1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = oem_b200.oem_fit_sparse(*args)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
st = out["stats"]
nnz = a.p * per_col
print(json.dumps({"config": f"oem_fit_sparse n={a.n} p={a.p} density={a.density} nnz={nnz} lasso+scad+mcp 100 lambdas",
                  "wall_s": best, "phases_ms": {k: round(v, 3) for k, v in st.items() if k.startswith("ms_")},
                  "row_pair_fma": float(per_col) * a.p * (a.density * a.p), "kernel_launches": st["kernel_launches"],
                  "oem_iterations": st["total_oem_iters"]}))
