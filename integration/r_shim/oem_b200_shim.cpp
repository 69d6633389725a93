// oem_b200_shim.cpp -- the reference-side binding: drop this file into the R package's src/ in place of oem_dense.cpp,
// oem_xtx.cpp, oem_xval_dense.cpp, oem_logistic_dense.cpp, oem_big.cpp, oem_fb_big.cpp and oem_sparse.cpp (and use the
// Makevars next to it).  The R code is unchanged: every `.Call("<symbol>", ..., PACKAGE = "oem")` (R/oem.R:534-575,
// 605-646, R/oem_xtx.R:389-420, R/oem_xval.R:525-548, R/big_oem.R:449-490) keeps its symbol, argument order and the
// named list it gets back; the bodies only unmarshal SEXPs into the plain structs of include/oem_b200.h and call the C ABI.
//
// Not compiled in this repository's image (no R / Rcpp); tests/test_abi.py syntax-checks it against a stub Rcpp.h.
#include <Rcpp.h>
#include <string>
#include <vector>
#include "oem_b200.h"
using namespace Rcpp;

namespace {

// The 17 arguments every fitting entry shares (family_ .. opts_), unmarshalled once.
struct Call {
    std::string family;
    std::vector<std::string> pen;
    std::vector<const char *> pen_c;
    NumericVector weights, group_weights, penalty_factor, gamma;
    IntegerVector groups, unique_groups;
    std::vector<NumericVector> lam;
    std::vector<const double *> lam_p;
    std::vector<int> lam_n;
    oemb200_spec s;
    oemb200_opts o;

    Call(SEXP family_, SEXP penalty_, SEXP weights_, SEXP groups_, SEXP unique_groups_, SEXP group_weights_,
         SEXP lambda_, SEXP nlambda_, SEXP lmin_ratio_, SEXP alpha_, SEXP gamma_, SEXP tau_, SEXP penalty_factor_,
         SEXP standardize_, SEXP intercept_, SEXP compute_loss_, SEXP opts_)
        : family(as<std::string>(as<CharacterVector>(family_)[0])), pen(as<std::vector<std::string> >(penalty_)),
          weights(weights_), group_weights(group_weights_), penalty_factor(penalty_factor_), gamma(gamma_),
          groups(groups_), unique_groups(unique_groups_) {
        for (size_t i = 0; i < pen.size(); ++i) pen_c.push_back(pen[i].c_str());
        List lambda(lambda_);
        for (int i = 0; i < lambda.size(); ++i) {
            lam.push_back(NumericVector((SEXP)lambda[i]));
            lam_p.push_back(lam.back().begin());
            lam_n.push_back((int)lam.back().size());
        }
        oemb200_spec z = {};
        s = z;
        s.family = family.c_str();
        s.n_penalty = (int)pen_c.size();
        s.penalty = pen_c.data();
        s.weights = weights.begin();             s.n_weights = weights.size();
        s.groups = groups.begin();               s.n_groups = (int)groups.size();
        s.unique_groups = unique_groups.begin(); s.n_unique_groups = (int)unique_groups.size();
        s.group_weights = group_weights.begin(); s.n_group_weights = (int)group_weights.size();
        s.lambda = lam_p.data();                 s.n_lambda = lam_n.data();
        s.nlambda = as<int>(nlambda_);           s.lambda_min_ratio = as<double>(lmin_ratio_);
        s.alpha = as<double>(alpha_);            s.gamma = gamma.begin();  s.n_gamma = (int)gamma.size();
        s.tau = as<double>(tau_);                s.penalty_factor = penalty_factor.begin();
        s.standardize = as<bool>(standardize_);  s.intercept = as<bool>(intercept_);
        s.compute_loss = as<bool>(compute_loss_);
        List opts(opts_);                        // src/oem_dense.cpp:91-102, oem_big.cpp:96-104
        oemb200_default_opts(&o);
        o.maxit = as<int>(opts["maxit"]);        o.tol = as<double>(opts["tol"]);
        if (opts.containsElementNamed("irls_maxit")) o.irls_maxit = as<int>(opts["irls_maxit"]);
        if (opts.containsElementNamed("irls_tol")) o.irls_tol = as<double>(opts["irls_tol"]);
        if (opts.containsElementNamed("ncores")) o.ncores = as<int>(opts["ncores"]);
        if (opts.containsElementNamed("accelerate")) o.accelerate = as<bool>(opts["accelerate"]);
        if (opts.containsElementNamed("hessian.type")) o.hessian_full = as<std::string>(opts["hessian.type"]) == "full";
        if (opts.containsElementNamed("gigs")) o.gigs = as<double>(opts["gigs"]);
    }
};

// Caller-allocated result buffers and the reference's return list (src/oem_dense.cpp:280-307, oem_xval_dense.cpp:469-476).
struct Result {
    int P, L, rows;
    std::vector<double> beta, lambda, loss, cvm, cvsd;
    std::vector<int> niter, nlam;
    double d;
    oemb200_result r;

    Result(const Call &c, int beta_rows, bool xval) : P(c.s.n_penalty), L(oemb200_nlambda_max(&c.s)), rows(beta_rows), d(0.0) {
        const size_t PL = (size_t)P * L;
        beta.assign(PL * rows, 0.0); lambda.assign(PL, 0.0); loss.assign(PL, 1e99); niter.assign(PL, 0); nlam.assign(P, 0);
        if (xval) { cvm.assign(PL, 0.0); cvsd.assign(PL, 0.0); }
        oemb200_result z = {};
        r = z;
        r.beta = beta.data(); r.lambda = lambda.data(); r.niter = niter.data(); r.loss = loss.data(); r.d = &d;
        r.cvm = xval ? cvm.data() : 0; r.cvsd = xval ? cvsd.data() : 0; r.nlam_out = nlam.data();
    }

    List pack(const Call &c, bool xval) const {
        List beta_list(P), lambda_list(P), iter_list(P), loss_list(P), cvm_list(P), cvsd_list(P);
        for (int pp = 0; pp < P; ++pp) {
            const int k = nlam[pp];
            const size_t off = (size_t)pp * L;
            NumericMatrix B(rows, k);
            std::copy(beta.begin() + off * rows, beta.begin() + off * rows + (size_t)rows * k, B.begin());
            if (c.pen[pp] == "ols") beta_list[pp] = NumericVector(B.begin(), B.begin() + rows);   // ols: a vector
            else beta_list[pp] = B;
            lambda_list[pp] = NumericVector(lambda.begin() + off, lambda.begin() + off + L);
            iter_list[pp] = IntegerVector(niter.begin() + off, niter.begin() + off + k);
            loss_list[pp] = NumericVector(loss.begin() + off, loss.begin() + off + k);
            if (xval) {
                cvm_list[pp] = NumericVector(cvm.begin() + off, cvm.begin() + off + k);
                cvsd_list[pp] = NumericVector(cvsd.begin() + off, cvsd.begin() + off + k);
            }
        }
        if (xval)
            return List::create(Named("beta") = beta_list, Named("lambda") = lambda_list, Named("niter") = iter_list,
                                Named("loss") = loss_list, Named("cvm") = cvm_list, Named("cvsd") = cvsd_list, Named("d") = d);
        return List::create(Named("beta") = beta_list, Named("lambda") = lambda_list, Named("niter") = iter_list,
                            Named("loss") = loss_list, Named("d") = d);
    }
};

inline void check(int rc) {
    if (rc != OEMB200_OK) Rcpp::stop(oemb200_last_error());     // surfaces as an R condition, like BEGIN_RCPP / END_RCPP today
}

}  // namespace

#define OEM_COMMON_SEXPS                                                                                                 \
    SEXP family_, SEXP penalty_, SEXP weights_, SEXP groups_, SEXP unique_groups_, SEXP group_weights_, SEXP lambda_,    \
    SEXP nlambda_, SEXP lmin_ratio_, SEXP alpha_, SEXP gamma_, SEXP tau_, SEXP penalty_factor_, SEXP standardize_,       \
    SEXP intercept_, SEXP compute_loss_, SEXP opts_
#define OEM_COMMON_ARGS                                                                                                  \
    family_, penalty_, weights_, groups_, unique_groups_, group_weights_, lambda_, nlambda_, lmin_ratio_, alpha_, gamma_,\
    tau_, penalty_factor_, standardize_, intercept_, compute_loss_, opts_

// src/oem_dense.cpp:30-48
RcppExport SEXP oem_fit_dense(SEXP x_, SEXP y_, OEM_COMMON_SEXPS) {
    BEGIN_RCPP
    NumericMatrix x(x_); NumericVector y(y_);                       // column-major REALSXP, borrowed, not modified
    Call c(OEM_COMMON_ARGS);
    Result res(c, x.ncol() + 1, false);
    check(oemb200_fit_dense(x.begin(), x.nrow(), x.ncol(), x.nrow(), y.begin(), &c.s, &c.o, &res.r));
    return res.pack(c, false);
    END_RCPP
}

// src/oem_logistic_dense.cpp:29-47
RcppExport SEXP oem_fit_logistic_dense(SEXP x_, SEXP y_, OEM_COMMON_SEXPS) {
    BEGIN_RCPP
    NumericMatrix x(x_); NumericVector y(y_);
    Call c(OEM_COMMON_ARGS);
    Result res(c, x.ncol() + 1, false);
    check(oemb200_fit_logistic_dense(x.begin(), x.nrow(), x.ncol(), x.nrow(), y.begin(), &c.s, &c.o, &res.r));
    return res.pack(c, false);
    END_RCPP
}

// src/oem_sparse.cpp:30-48 -- x is a Matrix::dgCMatrix
RcppExport SEXP oem_fit_sparse(SEXP x_, SEXP y_, OEM_COMMON_SEXPS) {
    BEGIN_RCPP
    S4 x(x_);
    IntegerVector xi = x.slot("i"), xp = x.slot("p"), dim = x.slot("Dim");
    NumericVector xx = x.slot("x"), y(y_);
    Call c(OEM_COMMON_ARGS);
    Result res(c, dim[1] + 1, false);
    check(oemb200_fit_sparse(xi.begin(), xp.begin(), xx.begin(), dim[0], dim[1], y.begin(), &c.s, &c.o, &res.r));
    return res.pack(c, false);
    END_RCPP
}

// src/oem_logistic_sparse.cpp:30-48 -- binomial family on a dgCMatrix (R/oem.R:605-625)
RcppExport SEXP oem_fit_logistic_sparse(SEXP x_, SEXP y_, OEM_COMMON_SEXPS) {
    BEGIN_RCPP
    S4 x(x_);
    IntegerVector xi = x.slot("i"), xp = x.slot("p"), dim = x.slot("Dim");
    NumericVector xx = x.slot("x"), y(y_);
    Call c(OEM_COMMON_ARGS);
    Result res(c, dim[1] + 1, false);
    check(oemb200_fit_logistic_sparse(xi.begin(), xp.begin(), xx.begin(), dim[0], dim[1], y.begin(), &c.s, &c.o, &res.r));
    return res.pack(c, false);
    END_RCPP
}

// Device-resident design matrix (include/oem_b200.h: oemb200_matrix_*): `oem_b200_matrix(x)` uploads an R matrix once and
// returns an external pointer whose finalizer frees the device copy -- the same mechanism as the `x@address` pointer of a
// big.matrix (src/oem_big.cpp:52-64).  oem_fit_dense / oem_fit_logistic_dense / oem_xval_dense below accept either a
// plain matrix or such a pointer in `x_`, so `oem(xd, y, ...)` refits without re-uploading x.
static void matrix_finalizer(oemb200_matrix *m) { oemb200_matrix_destroy(m); }
typedef XPtr<oemb200_matrix, PreserveStorage, matrix_finalizer> MatrixPtr;
RcppExport SEXP oem_b200_matrix(SEXP x_) {
    BEGIN_RCPP
    NumericMatrix x(x_);
    oemb200_opts o;
    oemb200_default_opts(&o);
    oemb200_matrix *m = NULL;
    check(oemb200_matrix_create(x.begin(), x.nrow(), x.ncol(), x.nrow(), &o, &m));
    return MatrixPtr(m, true);
    END_RCPP
}
RcppExport SEXP oem_fit_dense_h(SEXP x_, SEXP y_, OEM_COMMON_SEXPS) {
    BEGIN_RCPP
    MatrixPtr x(x_);
    NumericVector y(y_);
    int64_t n = 0; int p = 0;
    check(oemb200_matrix_info(x.get(), &n, &p, NULL, NULL, NULL, NULL));
    Call c(OEM_COMMON_ARGS);
    Result res(c, p + 1, false);
    if (c.family == "binomial") check(oemb200_fit_logistic_dense_h(x.get(), y.begin(), &c.s, &c.o, &res.r));
    else check(oemb200_fit_dense_h(x.get(), y.begin(), &c.s, &c.o, &res.r));
    return res.pack(c, false);
    END_RCPP
}

// src/oem_big.cpp:30-64 and src/oem_fb_big.cpp -- x is big.matrix@address (an external pointer to a BigMatrix); its
// matrix() is the (possibly memory-mapped) n x p column-major payload the reference wraps at oem_big.cpp:64.
#ifdef OEM_B200_WITH_BIGMEMORY
#include <bigmemory/BigMatrix.h>
static SEXP fit_big_impl(SEXP x_, SEXP y_, OEM_COMMON_SEXPS) {
    XPtr<BigMatrix> bm(x_);
    if (bm->matrix_type() != 8) throw Rcpp::exception("big.matrix type must be double");       // oem_big.cpp:57-62
    NumericVector y(y_);
    Call c(OEM_COMMON_ARGS);
    Result res(c, (int)bm->ncol() + 1, false);
    check(oemb200_fit_big((const double *)bm->matrix(), bm->nrow(), (int)bm->ncol(), bm->nrow(), y.begin(), &c.s, &c.o, &res.r));
    return res.pack(c, false);
}
RcppExport SEXP oem_fit_big(SEXP x_, SEXP y_, OEM_COMMON_SEXPS) {
    BEGIN_RCPP
    return fit_big_impl(x_, y_, OEM_COMMON_ARGS);
    END_RCPP
}
RcppExport SEXP oem_fit_fb_big(SEXP x_, SEXP y_, OEM_COMMON_SEXPS) {
    BEGIN_RCPP
    return fit_big_impl(x_, y_, OEM_COMMON_ARGS);
    END_RCPP
}
#endif

// src/oem_xtx.cpp:29-44 -- (xtx, xty, family, penalty, groups, unique_groups, group_weights, lambda, nlambda, lmin_ratio,
// alpha, gamma, tau, scale_factor, penalty_factor, opts): no weights / standardize / intercept / compute_loss
RcppExport SEXP oem_xtx(SEXP xtx_, SEXP xty_, SEXP family_, SEXP penalty_, SEXP groups_, SEXP unique_groups_,
                        SEXP group_weights_, SEXP lambda_, SEXP nlambda_, SEXP lmin_ratio_, SEXP alpha_, SEXP gamma_,
                        SEXP tau_, SEXP scale_factor_, SEXP penalty_factor_, SEXP opts_) {
    BEGIN_RCPP
    NumericMatrix xtx(xtx_); NumericVector xty(xty_), sf(scale_factor_);
    NumericVector none(0);
    LogicalVector f(1);                                              // FALSE
    Call c(family_, penalty_, none, groups_, unique_groups_, group_weights_, lambda_, nlambda_, lmin_ratio_, alpha_, gamma_,
           tau_, penalty_factor_, f, f, f, opts_);
    Result res(c, xtx.ncol(), false);                                // beta is p x L (src/oem_xtx.cpp:129)
    check(oemb200_xtx(xtx.begin(), xty.begin(), xtx.ncol(), &c.s, sf.size() ? sf.begin() : 0, (int)sf.size(), &c.o, &res.r));
    return res.pack(c, false);
    END_RCPP
}

// src/oem_xval_dense.cpp:31-52 -- (x, y, family, penalty, weights, groups, unique_groups, group_weights, lambda, nlambda,
// lmin_ratio, alpha, gamma, tau, penalty_factor, standardize, intercept, nfolds, foldid, compute_loss, type_measure, opts)
RcppExport SEXP oem_xval_dense(SEXP x_, SEXP y_, SEXP family_, SEXP penalty_, SEXP weights_, SEXP groups_,
                               SEXP unique_groups_, SEXP group_weights_, SEXP lambda_, SEXP nlambda_, SEXP lmin_ratio_,
                               SEXP alpha_, SEXP gamma_, SEXP tau_, SEXP penalty_factor_, SEXP standardize_,
                               SEXP intercept_, SEXP nfolds_, SEXP foldid_, SEXP compute_loss_, SEXP type_measure_,
                               SEXP opts_) {
    BEGIN_RCPP
    NumericMatrix x(x_); NumericVector y(y_);
    IntegerVector foldid(foldid_);                                    // 1-based fold ids
    std::string measure = as<std::string>(as<CharacterVector>(type_measure_)[0]);
    Call c(OEM_COMMON_ARGS);
    Result res(c, x.ncol() + 1, true);
    check(oemb200_xval_dense(x.begin(), x.nrow(), x.ncol(), x.nrow(), y.begin(), &c.s, as<int>(nfolds_), foldid.begin(),
                             measure.c_str(), &c.o, &res.r));
    return res.pack(c, true);
    END_RCPP
}

// predict.oem's last line, as.matrix(newx %*% nbeta) (R/methods.R:113-118; binomial response :355-358):
// .Call("oem_predict_b200", newx, nbeta, response) with a dense matrix or a dgCMatrix newx
RcppExport SEXP oem_predict_b200(SEXP newx_, SEXP nbeta_, SEXP response_) {
    BEGIN_RCPP
    NumericMatrix nbeta(nbeta_);                                      // (p + 1) x L with the intercept row, or p x L (oem.xtx)
    oemb200_opts o;
    oemb200_default_opts(&o);
    const int type = as<bool>(response_) ? 1 : 0;
    if (Rf_isS4(newx_)) {
        S4 x(newx_);
        IntegerVector xi = x.slot("i"), xp = x.slot("p"), dim = x.slot("Dim");
        NumericVector xx = x.slot("x");
        NumericMatrix out(dim[0], nbeta.ncol());
        check(oemb200_predict_sparse(xi.begin(), xp.begin(), xx.begin(), dim[0], dim[1], nbeta.begin(), nbeta.nrow(),
                                     nbeta.ncol(), type, out.begin(), dim[0], &o, 0));
        return out;
    }
    NumericMatrix newx(newx_);
    NumericMatrix out(newx.nrow(), nbeta.ncol());
    check(oemb200_predict(newx.begin(), newx.nrow(), newx.ncol(), newx.nrow(), nbeta.begin(), nbeta.nrow(), nbeta.ncol(), type,
                          out.begin(), newx.nrow(), &o, 0));
    return out;
    END_RCPP
}
