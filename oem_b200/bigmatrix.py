"""bigmemory file-backed matrices: the ingest format in front of big.oem().

R's `bigmemory::filebacked.big.matrix(..., type = "double", backingfile = "x.bk", descriptorfile = "x.desc")`
(used at R/big_oem.R:86-90 and vignettes/oem_vignette.Rmd:401-405) writes
  * `x.bk`   the raw column-major doubles (nrow * ncol * 8 bytes, no header) -- exactly the region the reference
             wraps with `Map<MatrixXd>` at src/oem_big.cpp:64, and
  * `x.desc` a dput() of the descriptor (totalRows, totalCols, type, filename, ...).
`attach()` maps the .bk file read-only; the resulting np.memmap goes straight into big_oem(), whose C entry
(oemb200_fit_big) streams it to the GPU in `gigs`-sized row chunks -- the file never has to fit in host or
device memory.  `write()` produces the same pair of files (for tests and for data prepared from Python)."""
import os
import re

import numpy as np

_TYPES = {"double": np.float64}


def read_descriptor(desc_path):
    txt = open(desc_path).read()

    def num(key):
        m = re.search(key + r"\s*=\s*([0-9.eE+]+)L?", txt)
        if not m:
            raise ValueError(f"{desc_path}: no '{key}' field")
        return int(float(m.group(1)))

    def string(key):
        m = re.search(key + r'\s*=\s*"([^"]*)"', txt)
        return m.group(1) if m else None

    d = dict(nrow=num("totalRows"), ncol=num("totalCols"), type=string("type") or "double",
             filename=string("filename"), dirname=string("dirname"))
    if re.search(r"separated\s*=\s*TRUE", txt):
        raise ValueError("separated (one file per column) big.matrix objects are not supported")
    return d


def attach(desc_path, backingpath=None):
    """-> read-only column-major np.memmap of shape (nrow, ncol), dtype float64."""
    d = read_descriptor(desc_path)
    if d["type"] != "double":
        raise ValueError("big.matrix type must be double")            # src/oem_big.cpp:57-62
    base = backingpath or os.path.dirname(os.path.abspath(desc_path))
    bk = os.path.join(base, d["filename"])
    expect = d["nrow"] * d["ncol"] * 8
    if os.path.getsize(bk) < expect:
        raise ValueError(f"{bk}: {os.path.getsize(bk)} bytes, descriptor needs {expect}")
    return np.memmap(bk, dtype=np.float64, mode="r", shape=(d["nrow"], d["ncol"]), order="F")


def write(x, bk_path, desc_path):
    """Write x (n x p) as a file-backed big.matrix pair (.bk + .desc)."""
    x = np.asarray(x, dtype=np.float64)
    n, p = x.shape
    mm = np.memmap(bk_path, dtype=np.float64, mode="w+", shape=(n, p), order="F")
    mm[:] = x
    mm.flush()
    del mm
    write_descriptor(desc_path, bk_path, n, p)


def write_descriptor(desc_path, bk_path, n, p):
    """The .desc file of an existing .bk backing file (n x p column-major doubles)."""
    with open(desc_path, "w") as f:
        f.write('new("big.matrix.descriptor", description = list(sharedType = "FileBacked", '
                f'filename = "{os.path.basename(bk_path)}", dirname = "{os.path.dirname(os.path.abspath(bk_path))}/", '
                f'totalRows = {n}L, totalCols = {p}L, rowOffset = c(0, {n}), colOffset = c(0, {p}), '
                f'nrow = {n}, ncol = {p}, rowNames = NULL, colNames = NULL, type = "double", separated = FALSE))\n')
