"""CPU tests of the boundary: the C-ABI library builds (nvcc cross-compiles sm_100a without a GPU), loads,
exports every symbol include/oem_b200.h declares, has no dependency on the oracle, and its compute entries
fail loudly (OEMB200_ENODEVICE) instead of falling back to a CPU path when no device is present."""
import os
import re
import subprocess

import numpy as np
import pytest

from cases import args_xy, gaussian_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    h = open(os.path.join(ROOT, "include", "oem_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(oemb200_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol(lib):
    L = lib.load()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/oem_b200.h but not exported"
    assert set(lib.EXPORTS) == set(syms)
    nm = subprocess.run(["nm", "-D", "--defined-only", lib.lib_path()], capture_output=True, text=True).stdout
    for s in syms:
        assert re.search(rf"\bT {s}\b", nm), s


def test_product_does_not_touch_the_oracle(lib):
    # the product library and package must not link, load or import anything under oracle/
    ldd = subprocess.run(["ldd", lib.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in ldd
    for dirpath, _, files in os.walk(os.path.join(ROOT, "oem_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "liboem_oracle" not in src and "oem_oracle.c" not in src, f


def test_sass_has_fp64_tensor_and_tma(lib):
    # the Gram kernel must be the FP64 tensor-core path staged by TMA: DMMA + UTMALDG in the sm_100a SASS
    out = subprocess.run(["cuobjdump", "-sass", lib.lib_path()], capture_output=True, text=True).stdout
    assert "DMMA.8x8x4" in out and "UTMALDG" in out
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", lib.lib_path()], capture_output=True, text=True).stdout


def test_host_helpers_match_oracle_bit_for_bit(lib, oracle):
    from oem_b200 import api
    for lmax, nl, r in [(0.731, 100, 1e-4), (12.5, 7, 1e-2), (3e-3, 200, 1e-4), (1.0, 1, 0.5)]:
        assert np.array_equal(api.lambda_grid(lmax, nl, r), oracle.lambda_base(lmax, nl, r))
    rng = np.random.default_rng(1)
    for _ in range(200):
        a = rng.normal(size=6) * (rng.uniform(size=6) > 0.3)
        b = a * (1 + rng.normal(size=6) * 10 ** rng.uniform(-9, -5)) * (rng.uniform(size=6) > 0.05)
        tol = 10 ** rng.uniform(-9, -5)
        assert api.stop_rule(a, b, tol) == oracle.stop_rule(a, b, tol)
    L = lib.load()
    names = ["lasso", "ols", "elastic.net", "scad", "scad.net", "mcp", "mcp.net", "grp.lasso", "grp.lasso.net",
             "grp.mcp", "grp.scad", "grp.mcp.net", "grp.scad.net", "sparse.grp.lasso"]
    assert [L.oemb200_penalty_id(n.encode()) for n in names] == [oracle.PEN_ID[n] for n in names]
    assert L.oemb200_penalty_id(b"ridge") == -1


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device behaviour is checked on the CPU box")
    X, y = gaussian_problem(1, 200, 5)
    a = args_xy(X, y, "gaussian", ["lasso"])
    for fn in (lib.oem_fit_dense, lib.oem_fit_big):
        with pytest.raises(lib.OemB200Error) as ei:
            fn(*a)
        assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)
    with pytest.raises(lib.OemB200Error) as ei:
        lib.oem_xtx(np.eye(5), np.ones(5), "gaussian", ["lasso"], [], [], [], [], 10, 1e-2, 1.0, 3.0, 0.5, [], np.ones(5), {})
    assert ei.value.code == 2


def test_argument_validation_precedes_device_use(lib):
    from oem_b200 import api
    X, y = gaussian_problem(1, 50, 4)
    with pytest.raises(ValueError):
        lib.oem_fit_dense(*args_xy(X, y[:-1], "gaussian", ["lasso"]))
    with pytest.raises(ValueError):
        api.make_opts({"no_such_option": 1})
    o = api.make_opts({"maxit": 10, "tol": 1e-9, "hessian.type": "full", "irls.maxit": 7})
    assert (o.maxit, o.tol, o.hessian_full, o.irls_maxit) == (10, 1e-9, 1, 7)


def test_header_is_plain_c_and_callable_from_c(lib, tmp_path):
    """The boundary is a C ABI: include/oem_b200.h compiles as C99 (no C++ / torch types in the signatures) and a C
    program linked against liboem_b200.so can call it -- host-only helpers here, the compute entries need a GPU."""
    inc = os.path.join(ROOT, "include")
    src = tmp_path / "abi_probe.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "oem_b200.h"
int main(void) {
    oemb200_opts o; oemb200_spec s; oemb200_result r; oemb200_stats st;
    double grid[5];
    memset(&s, 0, sizeof s); memset(&r, 0, sizeof r); memset(&st, 0, sizeof st);
    oemb200_default_opts(&o);
    if (o.maxit != 500 || o.irls_maxit != 100) return 2;
    if (oemb200_penalty_id("grp.lasso") < 0 || oemb200_penalty_id("nope") != -1) return 3;
    if (oemb200_lambda_grid(2.0, 5, 1e-2, grid) != OEMB200_OK) return 4;
    if (!(grid[0] == 2.0 && grid[4] < grid[3] && grid[3] < grid[0])) return 5;
    /* a compute entry with NULL arguments must come back as a status code, never crash or throw */
    if (oemb200_fit_sparse(0, 0, 0, 10, 2, 0, 0, 0, 0) == OEMB200_OK) return 6;
    if (strlen(oemb200_last_error()) == 0) return 7;
    printf("%s\n", oemb200_version());
    return 0;
}
''')
    exe = tmp_path / "abi_probe"
    libdir = os.path.dirname(lib.lib_path())
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, str(src), "-o", str(exe),
                           "-L", libdir, "-loem_b200", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "sm_100a" in out.stdout


def test_r_shim_matches_the_c_abi():
    """integration/r_shim/oem_b200_shim.cpp is the reference-side binding.  R / Rcpp are not in this image,
    so it is syntax-checked against stand-in headers (tests/stubs): every oemb200_* call in it must match the prototypes
    of include/oem_b200.h, and it must define each .Call symbol the R code of the reference looks up for this path."""
    shim = os.path.join(ROOT, "integration", "r_shim", "oem_b200_shim.cpp")
    out = subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-Wall", "-Werror", "-DOEM_B200_WITH_BIGMEMORY",
                          "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"), shim],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    src = open(shim).read()
    for sym in ("oem_fit_dense", "oem_fit_logistic_dense", "oem_fit_sparse", "oem_fit_logistic_sparse", "oem_fit_big",
                "oem_fit_fb_big", "oem_xtx", "oem_xval_dense", "oem_b200_matrix", "oem_fit_dense_h"):
        assert re.search(rf"RcppExport SEXP {sym}\(", src), sym
