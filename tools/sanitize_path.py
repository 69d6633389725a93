"""compute-sanitizer target: one small oem.xtx fit that runs the path kernel's global mode (register mat-vec, q = 480) and one
in the DMMA variant (5 chains) -- few lambdas, few iterations, so memcheck / racecheck finish in a minute."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oem_b200  # noqa: E402
from cases import gaussian_problem  # noqa: E402

p = 480
X, y = gaussian_problem(3, 2 * p, p, nnz=12)
xtx, xty = X.T @ X / X.shape[0], X.T @ y / X.shape[0]
g = np.arange(p) // 6 + 1
for pens in (["lasso", "mcp", "grp.lasso"], ["lasso", "mcp", "scad", "elastic.net", "scad.net"]):
    r = oem_b200.oem_xtx(xtx, xty, "gaussian", pens, g, np.unique(g), [], [], 3, 0.3, 0.7, 3.0, 0.4, np.sqrt(np.diag(xtx)),
                         np.ones(p), dict(maxit=25, tol=1e-7))
    print(pens, [int(n.max()) for n in r["niter"]], float(r["d"]))
