"""Collect the numbers the reference itself printed when its documentation was rendered and write them to
tests/golden/reference_printed.json.  Run in the build container only (reads /root/reference; the GPU box has
neither that tree nor any need for it -- the fixture travels).

Sources (rendered by pkgdown / knitr from the seeded examples in man/*.Rd and the vignette):
  docs/reference/predict.oem.html       test-set MSE of oem() for 10 lambdas x {lasso, grp.lasso}
  docs/reference/predict.xval.oem.html  test-set MSE of xval.oem()'s full-data fit at lambda.min
  docs/reference/logLik.html            logLik of oem() (100 lambdas), cv.oem()'s oem fit (25), xval.oem() (25); lasso + mcp
  docs/reference/oem.xtx.html           max |oem - oem.xtx| (rounding-level identity)
  vignettes/oem_vignette.html           max |big.oem - oem| on the seeded bigmemory example
  docs/reference/oem.html               max |dense - sparse| for the gaussian and the BINOMIAL entries (orders of magnitude)
Only printed OUTPUT values are copied (they are data, like golden vectors), never source code.
"""
import html
import json
import os
import re
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_printed.json")

_NUM = r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?"


def example_text(path):
    s = open(path).read()
    m = re.search(r'<pre class="examples">(.*?)</pre>', s, flags=re.S)
    return html.unescape(re.sub(r"<[^>]+>", "", m.group(1)))


def printed_after(text, call, occurrence=0):
    """The `#> [k] v v v` lines that directly follow the `occurrence`-th appearance of `call`."""
    pos = -1
    for _ in range(occurrence + 1):
        pos = text.index(call, pos + 1)
    rest = text[pos + len(call):]
    vals = []
    for m in re.finditer(r"#>\s*\[\d+\]((?:\s+" + _NUM + r")+)\s*", rest):
        if m.start() != 0 and not vals:
            break
        if vals and rest[last_end:m.start()].strip():
            break
        vals += m.group(1).split()
        last_end = m.end()
    return _entry(vals)


def _unit(tok):
    """One unit of the last printed digit of an R-formatted number."""
    mant, _, ex = tok.lower().partition("e")
    dec = len(mant.split(".")[1]) if "." in mant else 0
    return 10.0 ** (-dec + (int(ex) if ex else 0))


def _entry(tokens):
    return {"values": [float(t) for t in tokens], "unit": max(_unit(t) for t in tokens) if tokens else None}


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container only)")
    out = {"_comment": "values printed by the reference's own rendered documentation; made by tools/extract_reference_outputs.py"}
    t = example_text(f"{REF}/docs/reference/predict.oem.html")
    out["predict_oem"] = {
        "source": "docs/reference/predict.oem.html (man/predict.oem.Rd:43-63)",
        "mse_lasso": printed_after(t, "apply(preds.lasso,     2, function(x) mean((y.test - x) ^ 2))"),
        "mse_grp_lasso": printed_after(t, "apply(preds.grp.lasso, 2, function(x) mean((y.test - x) ^ 2))"),
    }
    t = example_text(f"{REF}/docs/reference/predict.xval.oem.html")
    out["predict_xval_oem"] = {
        "source": "docs/reference/predict.xval.oem.html",
        "mse_best": printed_after(t, "apply(preds.best, 2, function(x) mean((y.test - x) ^ 2))"),
        "mse_grp_lasso": printed_after(t, "apply(preds.gl, 2, function(x) mean((y.test - x) ^ 2))"),
        "mse_lasso": printed_after(t, "apply(preds.l, 2, function(x) mean((y.test - x) ^ 2))"),
    }
    t = example_text(f"{REF}/docs/reference/predict.cv.oem.html")
    out["predict_cv_oem"] = {
        "source": "docs/reference/predict.cv.oem.html (cv.oem's full-data fit is oem())",
        "mse_best": printed_after(t, "apply(preds.best, 2, function(x) mean((y.test - x) ^ 2))"),
        "mse_grp_lasso": printed_after(t, "apply(preds.gl, 2, function(x) mean((y.test - x) ^ 2))"),
        "mse_lasso": printed_after(t, "apply(preds.l, 2, function(x) mean((y.test - x) ^ 2))"),
    }
    t = example_text(f"{REF}/docs/reference/logLik.html")
    out["logLik"] = {
        "source": "docs/reference/logLik.html (man/logLik.Rd:32-55)",
        "oem_lasso": printed_after(t, "logLik(fit)", 0),
        "oem_mcp": printed_after(t, 'logLik(fit, which.model = "mcp")', 0),
        "cv_oem_lasso": printed_after(t, "logLik(fit)", 1),
        "cv_oem_mcp": printed_after(t, 'logLik(fit, which.model = "mcp")', 1),
        "xval_oem_lasso": printed_after(t, "logLik(fit)", 2),
        "xval_oem_mcp": printed_after(t, 'logLik(fit, which.model = "mcp")', 2),
    }
    t = example_text(f"{REF}/docs/reference/oem.xtx.html")
    out["oem_xtx"] = {
        "source": "docs/reference/oem.xtx.html (man/oem.xtx.Rd:113-145)",
        "maxdiff_lasso": printed_after(t, "max(abs(fit$beta[[1]][-1,] - fit.xtx$beta[[1]]))"),
        "maxdiff_enet": printed_after(t, "max(abs(fit$beta[[2]][-1,] - fit.xtx$beta[[2]]))"),
    }
    s = open(f"{REF}/vignettes/oem_vignette.html").read()
    s = html.unescape(re.sub(r"<[^>]+>", "", s))
    i = s.index("max(abs(fit$beta[[1]] - fit2$beta[[1]]))")
    m = re.search(r"##\s*\[1\]\s+(" + _NUM + ")", s[i:])
    out["vignette_bigmat"] = {
        "source": "vignettes/oem_vignette.html (vignettes/oem_vignette.Rmd:398-428)",
        "maxdiff_big_vs_oem_lasso": _entry([m.group(1)]),
    }
    t = example_text(f"{REF}/docs/reference/oem.html")
    out["oem_rd"] = {
        "source": "docs/reference/oem.html (man/oem.Rd:138-211): rsparsematrix inputs are not reproducible, orders of magnitude only",
        "maxdiff_dense_vs_sparse_lasso": printed_after(t, "max(abs(fit$beta[[1]] - fits$beta[[1]]))"),
        "maxdiff_dense_vs_sparse_grp_lasso": printed_after(t, "max(abs(fit$beta[[2]] - fits$beta[[2]]))"),
        # the one number the reference prints for its binomial entries: oem_fit_logistic_dense against
        # oem_fit_logistic_sparse on the same data (intercept = FALSE, grp.lasso, 10 lambdas, irls.tol 1e-3, tol 1e-8)
        "maxdiff_logistic_dense_vs_sparse_grp_lasso": printed_after(t, "max(abs(res.gr$beta[[1]] - res.gr.s$beta[[1]]))"),
    }
    for k, v in out.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                if kk != "source" and not vv["values"]:
                    sys.exit(f"nothing extracted for {k}.{kk}")
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT, {k: {kk: (len(vv["values"]), vv["unit"]) for kk, vv in v.items() if kk != "source"} for k, v in out.items() if isinstance(v, dict)})


if __name__ == "__main__":
    main()
