"""Parity pinned on the reference's own outputs.

The reference has no test suite, but its rendered documentation (docs/reference/*.html, vignettes/oem_vignette.html)
shows what `oem()`, `xval.oem()`, `oem.xtx()` and `big.oem()` returned on seeded inputs (`set.seed(123)`).  oracle/r_rng.py
regenerates those inputs, tests/reference_examples.py re-runs the examples, and the results must round to the digits
the reference printed (|got - printed| <= half a unit of the last printed digit): 20 + 5 test-set MSEs of
predict.oem / predict.cv.oem (the latter also through the cv.oem mirror itself), 3 of predict.xval.oem, 300 log-likelihoods (compute.loss paths of oem and xval.oem,
lasso + mcp), the oem <-> oem.xtx identity over twelve penalties and max |big.oem - oem| of the bigmemory example.

CPU leg: the oracle (through the front-end mirrors) against the printed values -- this is what pins the oracle.
GPU leg: the CUDA library through the same front-ends and the C ABI against the same printed values.
"""
import numpy as np
import pytest

import reference_examples as ex


def test_r_stream_known_answers():
    """Any R session: set.seed(123); runif(3) / rnorm(3); set.seed(1); rnorm(2); set.seed(42); runif(2)."""
    from oracle.r_rng import RStream
    assert np.allclose(RStream(123).runif(3), [0.2875775, 0.7883051, 0.4089769], rtol=0, atol=5e-8)
    assert np.allclose(RStream(123).rnorm(3), [-0.56047565, -0.23017749, 1.55870831], rtol=0, atol=5e-9)
    assert np.allclose(RStream(1).rnorm(2), [-0.6264538, 0.1836433], rtol=0, atol=5e-8)
    assert np.allclose(RStream(42).runif(2), [0.9148060, 0.9370754], rtol=0, atol=5e-8)
    m = RStream(7).matrix_rnorm(3, 2)
    assert m.flags["F_CONTIGUOUS"] and np.array_equal(m.ravel(order="F"), RStream(7).rnorm(6))      # column-major fill


def test_fixture_shape():
    g = ex.printed()
    assert len(g["predict_oem"]["mse_lasso"]["values"]) == 10 and g["predict_oem"]["mse_lasso"]["unit"] == 1e-6
    assert [len(g["logLik"][k]["values"]) for k in ("oem_lasso", "oem_mcp", "cv_oem_lasso", "cv_oem_mcp",
                                                    "xval_oem_lasso", "xval_oem_mcp")] == [100, 100, 25, 25, 25, 25]
    assert g["logLik"]["oem_lasso"]["unit"] == 1e-3
    assert g["vignette_bigmat"]["maxdiff_big_vs_oem_lasso"]["values"] == [1.534783e-05]
    assert g["oem_rd"]["maxdiff_logistic_dense_vs_sparse_grp_lasso"]["values"] == [6.647085e-05]


class _OracleApi:
    """Stands in for oem_b200.api behind the front-end mirrors in the CPU leg ONLY (tests may call the oracle; the
    product never does).  Same entry-point names and argument order; `comm` is dropped."""

    def __init__(self, orc):
        self._o = orc

    def __getattr__(self, name):
        fn = getattr(self._o, name)

        def call(*a, comm=None, **k):
            return fn(*a, **k)
        return call

    @staticmethod
    def predict_matrix(newx, beta, response=False, opts=None):
        b = np.asarray(beta, dtype=np.float64)
        eta = b[0][None, :] + np.asarray(newx) @ b[1:]
        return 1.0 / (1.0 + np.exp(-eta)) if response else eta


@pytest.fixture()
def fe_oracle(oracle, monkeypatch):
    from oem_b200 import frontend as fe
    monkeypatch.setattr(fe, "api", _OracleApi(oracle))
    return fe


@pytest.fixture()
def fe_gpu(lib):
    from oem_b200 import frontend as fe
    return fe


# ------------------------------------------------------------ CPU: the oracle is pinned on the reference's printed outputs
def test_oracle_predict_oem(fe_oracle):
    ex.assert_printed(ex.example_predict_oem(fe_oracle))


def test_oracle_predict_xval_oem(fe_oracle):
    ex.assert_printed(ex.example_predict_xval_oem(fe_oracle))


def test_oracle_predict_cv_oem(fe_oracle):
    res, fit = ex.example_predict_cv_oem(fe_oracle)
    ex.assert_printed(res)
    assert fit["best_model"] == "grp.lasso" and fit["name"] == "Mean-Squared Error"
    # each fold generates its own lambda sequence; full-fit lambdas below a fold's smallest one are not extrapolated and
    # drop out as NA (R/cv_oem.R:360-368, 192-206): here the last of the ten
    assert len(fit["cvm"]) == 2 and fit["cvm"][0].shape == (9,) and fit["lambda"][0].shape == (9,)
    assert np.all(fit["cvsd"][1] > 0) and np.all(fit["cvup"][0] > fit["cvlo"][0])
    assert fit["lambda_min"] == fit["lambda"][1][int(np.argmin(fit["cvm"][1]))]


def test_oracle_logLik(fe_oracle):
    ex.assert_printed(ex.example_logLik(fe_oracle))


def test_oracle_oem_xtx_identity(fe_oracle):
    assert np.all(ex.example_oem_xtx(fe_oracle) <= 1e-13)      # reference printed 8.788848e-16


def test_oracle_vignette_bigmat(fe_oracle):
    ex.assert_printed(ex.example_vignette_bigmat(fe_oracle))


def test_oracle_logistic_dense_vs_sparse_order(fe_oracle):
    # the reference's only printed binomial number (6.647085e-05): the two logistic entries differ at the IRLS tolerance
    # level, not at rounding level and not grossly -- same order for the restated pair (inputs are a stand-in, see the example)
    d, printed_value = ex.example_logistic_dense_vs_sparse(fe_oracle)
    assert printed_value == 6.647085e-05
    assert printed_value / 50 < d < printed_value * 50, d


# ------------------------------------------------------------ GPU: the CUDA path against the same printed outputs
@pytest.mark.gpu
def test_gpu_logistic_dense_vs_sparse_order(fe_gpu):
    d, printed_value = ex.example_logistic_dense_vs_sparse(fe_gpu)
    assert printed_value / 50 < d < printed_value * 50, d


@pytest.mark.gpu
def test_gpu_predict_oem(fe_gpu):
    ex.assert_printed(ex.example_predict_oem(fe_gpu))


@pytest.mark.gpu
def test_gpu_predict_xval_oem(fe_gpu):
    ex.assert_printed(ex.example_predict_xval_oem(fe_gpu))


@pytest.mark.gpu
def test_gpu_predict_cv_oem(fe_gpu):
    res, fit = ex.example_predict_cv_oem(fe_gpu)
    ex.assert_printed(res)
    assert fit["best_model"] == "grp.lasso"


@pytest.mark.gpu
def test_gpu_logLik(fe_gpu):
    ex.assert_printed(ex.example_logLik(fe_gpu))


@pytest.mark.gpu
def test_gpu_oem_xtx_identity(fe_gpu):
    assert np.all(ex.example_oem_xtx(fe_gpu) <= 1e-12)


@pytest.mark.gpu
def test_gpu_vignette_bigmat(fe_gpu):
    # the printed value is a DIFFERENCE of two coefficient paths (1.5e-05 to 7 digits, unit 1e-11): one full unit allows
    # for the ~1e-12 summation-order differences between the GPU Gram and the reference's BLAS
    ex.assert_printed(ex.example_vignette_bigmat(fe_gpu), units=1.0)


# ------------------------------------------------------------ cv.oem mirror, binomial measures (host logic, oracle behind it)
def test_cv_oem_binomial_measures(fe_oracle):
    from cases import binomial_problem
    X, y = binomial_problem(9, 1500, 8)
    foldid = 1 + (np.arange(1500) % 5)
    dev = fe_oracle.cv_oem(X, y, family="binomial", penalty="lasso", nlambda=6, lambda_min_ratio=0.05, foldid=foldid)
    cls = fe_oracle.cv_oem(X, y, family="binomial", penalty="lasso", nlambda=6, lambda_min_ratio=0.05, foldid=foldid,
                           type_measure="class")
    assert dev["name"] == "Binomial Deviance" and cls["name"] == "Misclassification Error"
    # the first lambda of the full fit is >= every fold's lambda_max up to sampling noise: there the held-out prediction is
    # (nearly) the training folds' mean of y, so the deviance is the null deviance and the error rate min(ybar, 1 - ybar)
    null_dev, null_err = [], []
    for k in range(1, 6):
        p = y[foldid != k].mean()
        yk = y[foldid == k]
        null_dev.append(-2 * np.mean(yk * np.log(p) + (1 - yk) * np.log(1 - p)))
        null_err.append(np.mean(yk == (p <= 0.5)))
    assert abs(dev["cvm"][0][0] - np.mean(null_dev)) < 5e-3
    assert abs(cls["cvm"][0][0] - np.mean(null_err)) < 2e-2
    assert dev["cvm"][0].min() < dev["cvm"][0][0] and np.all(dev["cvsd"][0] > 0)
    assert dev["lambda_1se"] >= dev["lambda_min"] and cls["lambda_min"] > 0
    auc = fe_oracle.cv_oem(X, y, family="binomial", penalty="lasso", nlambda=6, lambda_min_ratio=0.05, foldid=foldid,
                           type_measure="auc")
    assert auc["name"] == "AUC" and abs(auc["cvm"][0][0] - 0.5) < 0.05 and auc["cvm"][0].max() > 0.55
    assert auc["lambda_min"] == auc["lambda"][0][int(np.argmax(auc["cvm"][0]))]        # AUC is maximised
    with pytest.raises(ValueError, match="nfolds must be bigger than 3"):
        fe_oracle.cv_oem(X, y, family="binomial", penalty="lasso", nfolds=2)
