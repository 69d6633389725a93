/*
 * oracle/oem_oracle.c -- TEST INFRASTRUCTURE ONLY (see DESIGN.md section 2 for what pins it).
 *
 * Plain-C CPU restatement of the OEM iteration of jaredhuling/oem 2.0.12.  It is
 * the checker for the CUDA path: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * library (oem_b200/lib/liboem_b200.so) never links, loads or calls it.
 *
 * The reference cannot be compiled here (needs R, Rcpp, Eigen, Spectra) and
 * ships no test suite; this iteration is pinned through oracle/oracle.py on
 * the values the reference's rendered documentation prints for seeded inputs
 * (tests/test_reference_pins.py: gaussian entry points; the logistic entry
 * remains "parity unpinned").  Every function below cites the reference
 * file:line it restates (paths relative to /root/reference).
 *
 * Build: gcc -O2 -fPIC -shared -o oracle/_build/liboem_oracle.so oracle/oem_oracle.c -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* penalty ids shared with oracle/oracle.py (and, by value, with include/oem_b200.h) */
enum {
    PEN_LASSO = 0, PEN_OLS = 1, PEN_ENET = 2, PEN_SCAD = 3, PEN_SCAD_NET = 4,
    PEN_MCP = 5, PEN_MCP_NET = 6, PEN_GRP_LASSO = 7, PEN_GRP_LASSO_NET = 8,
    PEN_GRP_MCP = 9, PEN_GRP_SCAD = 10, PEN_GRP_MCP_NET = 11,
    PEN_GRP_SCAD_NET = 12, PEN_SPARSE_GRP_LASSO = 13
};

/* src/utils.cpp:537-549  stopRule(cur, prev, tolerance) */
int oracle_stop_rule(int q, const double *cur, const double *prev, double tol)
{
    for (int i = 0; i < q; i++) {
        double ac = fabs(cur[i]), ap = fabs(prev[i]);
        if ((ac > 1e-13 && ap <= 1e-13) || (ac <= 1e-13 && ap > 1e-13)) return 0;
        if (ac > 1e-13 && ap > 1e-13 && fabs((cur[i] - prev[i]) / prev[i]) > tol) return 0;
    }
    return 1;
}

/* src/oem_dense.h:76-92  soft_threshold */
static void soft_threshold(int q, double *res, const double *vec, double penalty,
                           const double *pen_fact, double d)
{
    for (int i = 0; i < q; i++) {
        double total_pen = pen_fact[i] * penalty;
        res[i] = 0.0;
        if (vec[i] > total_pen) res[i] = (vec[i] - total_pen) / d;
        else if (vec[i] < -total_pen) res[i] = (vec[i] + total_pen) / d;
    }
}

/* src/oem_dense.h:94-117  soft_threshold_mcp */
static void soft_threshold_mcp(int q, double *res, const double *vec, double penalty,
                               const double *pen_fact, double d, double gamma)
{
    double gammad = gamma * d;
    double d_minus_gammainv = d - 1.0 / gamma;
    for (int i = 0; i < q; i++) {
        double total_pen = pen_fact[i] * penalty;
        res[i] = 0.0;
        if (fabs(vec[i]) > gammad * total_pen) res[i] = vec[i] / d;
        else if (vec[i] > total_pen) res[i] = (vec[i] - total_pen) / d_minus_gammainv;
        else if (vec[i] < -total_pen) res[i] = (vec[i] + total_pen) / d_minus_gammainv;
    }
}

/* src/oem_dense.h:119-149  soft_threshold_scad */
static void soft_threshold_scad(int q, double *res, const double *vec, double penalty,
                                const double *pen_fact, double d, double gamma)
{
    double gammad = gamma * d;
    double gamma_minus1_d = (gamma - 1.0) * d;
    for (int i = 0; i < q; i++) {
        double total_pen = pen_fact[i] * penalty;
        res[i] = 0.0;
        if (fabs(vec[i]) > gammad * total_pen) res[i] = vec[i] / d;
        else if (fabs(vec[i]) > (d + 1.0) * total_pen) {
            double gam_ptr = (gamma - 1.0) * vec[i];
            double gam_pen = gamma * total_pen;
            if (gam_ptr > gam_pen) res[i] = (gam_ptr - gam_pen) / (gamma_minus1_d - 1.0);
            else if (gam_ptr < -gam_pen) res[i] = (gam_ptr + gam_pen) / (gamma_minus1_d - 1.0);
        }
        else if (vec[i] > total_pen) res[i] = (vec[i] - total_pen) / d;
        else if (vec[i] < -total_pen) res[i] = (vec[i] + total_pen) / d;
    }
}

/* src/oem_dense.h:151-174  soft_threshold_scad_norm */
static double scad_norm(double b, double pen, double d, double gamma)
{
    double retval = 0.0;
    double gammad = gamma * d;
    double gamma_minus1_d = (gamma - 1.0) * d;
    if (fabs(b) > gammad * pen) retval = 1;
    else if (fabs(b) > (d + 1.0) * pen) {
        double gam_ptr = (gamma - 1.0);
        double gam_pen = gamma * pen / b;
        if (gam_ptr > gam_pen) retval = d * (gam_ptr - gam_pen) / (gamma_minus1_d - 1.0);
        else if (gam_ptr < -gam_pen) retval = d * (gam_ptr + gam_pen) / (gamma_minus1_d - 1.0);
    }
    else if (b > pen) retval = (1.0 - pen / b);
    else if (b < -pen) retval = (1.0 + pen / b);
    return retval;
}

/* src/oem_dense.h:176-191  soft_threshold_mcp_norm */
static double mcp_norm(double b, double pen, double d, double gamma)
{
    double retval = 0.0;
    double gammad = gamma * d;
    double d_minus_gammainv = d - 1.0 / gamma;
    if (fabs(b) > gammad * pen) retval = 1;
    else if (b > pen) retval = d * (1.0 - pen / b) / d_minus_gammainv;
    else if (b < -pen) retval = d * (1.0 + pen / b) / d_minus_gammainv;
    return retval;
}

/* src/oem_dense.h:193-315  block_soft_threshold{,_mcp,_scad}; kind 0=lasso 1=mcp 2=scad.
 * Groups are given in CSR form (grp_ptr/grp_idx), members ascending as built by
 * get_group_indexes (src/oem_dense.h:421-456).  pen_fact is indexed by group. */
static void block_soft_threshold(int kind, int q, double *res, const double *vec, double penalty,
                                 const double *pen_fact, double d, int ngroups,
                                 const int *unique_grps, const int *grp_ptr, const int *grp_idx,
                                 double gamma)
{
    for (int i = 0; i < q; i++) res[i] = 0.0;
    for (int g = 0; g < ngroups; g++) {
        double thresh_factor;
        if (unique_grps[g] == 0) thresh_factor = 1.0;
        else {
            double ds_norm = 0.0;
            for (int v = grp_ptr[g]; v < grp_ptr[g + 1]; v++) ds_norm += pow(vec[grp_idx[v]], 2);
            ds_norm = sqrt(ds_norm);
            double grp_wts = pen_fact[g];
            if (kind == 0) {
                double t = 1.0 - penalty * grp_wts / ds_norm;
                thresh_factor = (0.0 < t) ? t : 0.0;     /* std::max(0.0, t): NaN -> 0.0 */
            } else if (kind == 1) thresh_factor = mcp_norm(ds_norm, penalty * grp_wts, d, gamma);
            else thresh_factor = scad_norm(ds_norm, penalty * grp_wts, d, gamma);
        }
        if (thresh_factor != 0.0)
            for (int v = grp_ptr[g]; v < grp_ptr[g + 1]; v++) {
                int c = grp_idx[v];
                res[c] = vec[c] * thresh_factor / d;
            }
    }
}

typedef struct {
    int q;                 /* dimension of beta (p, or p+1 with an explicit intercept) */
    int penalty;           /* PEN_* */
    double alpha, gamma, tau;
    const double *pen_fact;     /* q */
    int ngroups;
    const int *unique_groups;   /* ngroups */
    const int *grp_ptr;         /* ngroups+1 */
    const int *grp_idx;
    const double *group_weights; /* ngroups */
} oracle_pen;

/* next_beta dispatch: src/oem_dense.h:527-629 (byte-identical twins in oem_xtx.h:373-470,
 * oem_xval_dense.h:874-971, oem_logistic_dense.h, oem_big.h) */
static void next_beta(const oracle_pen *P, double lambda, double d, const double *u,
                      double *beta, double *tmp)
{
    int q = P->q;
    double denom = d + (1.0 - P->alpha) * lambda;
    double lam = lambda * P->alpha;
    switch (P->penalty) {
    case PEN_LASSO: soft_threshold(q, beta, u, lambda, P->pen_fact, d); break;
    case PEN_OLS: for (int i = 0; i < q; i++) beta[i] = u[i] / d; break;
    case PEN_ENET: soft_threshold(q, beta, u, lam, P->pen_fact, denom); break;
    case PEN_SCAD: soft_threshold_scad(q, beta, u, lambda, P->pen_fact, d, P->gamma); break;
    case PEN_SCAD_NET:
        if (P->alpha == 0) { lam = 0; denom = d + lambda; }
        soft_threshold_scad(q, beta, u, lam, P->pen_fact, denom, P->gamma); break;
    case PEN_MCP: soft_threshold_mcp(q, beta, u, lambda, P->pen_fact, d, P->gamma); break;
    case PEN_MCP_NET: soft_threshold_mcp(q, beta, u, lam, P->pen_fact, denom, P->gamma); break;
    case PEN_GRP_LASSO:
        block_soft_threshold(0, q, beta, u, lambda, P->group_weights, d, P->ngroups,
                             P->unique_groups, P->grp_ptr, P->grp_idx, P->gamma); break;
    case PEN_GRP_LASSO_NET:
        block_soft_threshold(0, q, beta, u, lam, P->group_weights, denom, P->ngroups,
                             P->unique_groups, P->grp_ptr, P->grp_idx, P->gamma); break;
    case PEN_GRP_MCP:
        block_soft_threshold(1, q, beta, u, lambda, P->group_weights, d, P->ngroups,
                             P->unique_groups, P->grp_ptr, P->grp_idx, P->gamma); break;
    case PEN_GRP_SCAD:
        block_soft_threshold(2, q, beta, u, lambda, P->group_weights, d, P->ngroups,
                             P->unique_groups, P->grp_ptr, P->grp_idx, P->gamma); break;
    case PEN_GRP_MCP_NET:
        block_soft_threshold(1, q, beta, u, lam, P->group_weights, denom, P->ngroups,
                             P->unique_groups, P->grp_ptr, P->grp_idx, P->gamma); break;
    case PEN_GRP_SCAD_NET:
        block_soft_threshold(2, q, beta, u, lam, P->group_weights, denom, P->ngroups,
                             P->unique_groups, P->grp_ptr, P->grp_idx, P->gamma); break;
    case PEN_SPARSE_GRP_LASSO: {
        double lam_grp = (1.0 - P->tau) * lambda;
        double lam_l1 = P->tau * lambda;
        soft_threshold(q, tmp, u, lam_l1, P->pen_fact, 1.0);
        block_soft_threshold(0, q, beta, tmp, lam_grp, P->group_weights, d, P->ngroups,
                             P->unique_groups, P->grp_ptr, P->grp_idx, P->gamma); break;
    }
    default: break;
    }
}

/* One solve(maxit) at a fixed lambda: src/oem_base.h:90-110 with
 * next_u = A*beta_prev + XY (src/oem_dense.h:508-513) and, when accelerate != 0,
 * the Nesterov step of src/oem_dense.h:633-651 (state *ak; reset to 1 by init()).
 * A is q x q column-major (symmetric).  beta is read (warm start) and overwritten.
 * Returns i+1 (maxit+1 when the stop rule never fired). */
int oracle_solve(const oracle_pen *P, const double *A, const double *XY, double d, double lambda,
                 int maxit, double tol, int accelerate, double *ak, double *beta)
{
    int q = P->q, i;
    double *u = (double *)malloc(sizeof(double) * q * 4);
    double *beta_prev = u + q, *tmp = u + 2 * q, *beta_last = u + 3 * q;
    for (i = 0; i < maxit; i++) {
        memcpy(beta_prev, beta, sizeof(double) * q);
        /* u = A * beta_prev + XY : column sweep (A symmetric, column-major) */
        for (int r = 0; r < q; r++) u[r] = 0.0;
        for (int c = 0; c < q; c++) {
            double b = beta_prev[c];
            if (b == 0.0) continue;             /* adds exact zeros otherwise */
            const double *col = A + (size_t)c * q;
            for (int r = 0; r < q; r++) u[r] += col[r] * b;
        }
        for (int r = 0; r < q; r++) u[r] += XY[r];
        memcpy(beta_last, beta, sizeof(double) * q);
        next_beta(P, lambda, d, u, beta, tmp);
        if (accelerate) {
            double ak_prev = *ak;
            *ak = 0.5 * (1 + sqrt(1.0 + 4.0 * pow(*ak, 2)));
            double ratio_k = (ak_prev - 1.0) / *ak;
            double adaptive_val = 0.0;
            for (int r = 0; r < q; r++) {
                double beta_update = beta[r];
                double beta_diff = beta[r] - beta_last[r];
                beta[r] += ratio_k * beta_diff;
                adaptive_val += (beta[r] - beta_update) * beta_diff;
            }
            if (adaptive_val > 0) *ak = 1;
        }
        if (oracle_stop_rule(q, beta, beta_prev, tol)) break;
    }
    free(u);
    return i + 1;
}

/* Lower-triangle SYRK G = X'X for a column-major n x p block, row-sliced with
 * per-thread partials summed in a critical section: src/oem_dense.h:318-361
 * (OpenMP variant).  Used as the "port" CPU baseline when numpy/OpenBLAS is not
 * wanted; result is the full symmetric matrix like Eigen's selfadjointView copy. */
void oracle_xtx(long n, int p, const double *X, double *G, int ncores)
{
    if (ncores < 1) ncores = 1;
    for (long k = 0; k < (long)p * p; k++) G[k] = 0.0;
    long first = n / ncores;
#ifdef _OPENMP
#pragma omp parallel num_threads(ncores)
#endif
    {
        double *Gp = (double *)calloc((size_t)p * p, sizeof(double));
#ifdef _OPENMP
#pragma omp for schedule(static) nowait
#endif
        for (int ff = 0; ff < ncores; ff++) {
            long r0 = ff * first;
            long r1 = (ff + 1 == ncores) ? n : r0 + first;
            for (int a = 0; a < p; a++)
                for (int b = a; b < p; b++) {
                    const double *xa = X + (size_t)a * n, *xb = X + (size_t)b * n;
                    double s = 0.0;
                    for (long r = r0; r < r1; r++) s += xa[r] * xb[r];
                    Gp[b + (size_t)a * p] += s;
                }
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        {
            for (long k = 0; k < (long)p * p; k++) G[k] += Gp[k];
        }
        free(Gp);
    }
    for (int a = 0; a < p; a++)
        for (int b = a + 1; b < p; b++) G[a + (size_t)b * p] = G[b + (size_t)a * p];
}
