// path_kernel.cu -- everything after the Gram, in ONE persistent cooperative kernel:
//   * top eigenvalue of XX by Lanczos  -> d = factor * lambda_max      (stands in for the
//     Spectra::SymEigsSolver call sites: src/oem_dense.h:485-498, oem_xtx.h:357-369,
//     oem_xval_dense.h:770-782,837-843, oem_logistic_dense.h:501-514, oem_big.h:546-559)
//   * A = d I - XX                                                     (src/oem_dense.h:501-505)
//   * for every chain (Gram x penalty), the warm-started lambda path of
//       beta <- prox_penalty(A beta + XY)            oemBase::solve  src/oem_base.h:90-110
//       next_u                                       src/oem_dense.h:508-513
//       next_beta + thresholding family              src/oem_dense.h:76-315, 527-629
//       Nesterov option                              src/oem_dense.h:633-651
//       stopRule                                     src/utils.cpp:537-549
//
// Design.  The CTAs are split into one TEAM per Gram (1 team for oem / oem.xtx / big.oem / the
// logistic inner loop, nfolds+1 teams for xval.oem).  A team member owns a slice of columns of A
// (A is symmetric, so a column slice is a row slice) which it keeps in SHARED MEMORY for the whole
// path whenever it fits (148 x ~200 KB covers q = 1001 with room to spare); otherwise the slice
// is streamed from L2.  Each OEM iteration is
//     u[slice] = A[:,slice]' beta + XY[slice]   for all of the team's chains at once (the GEMV
//                                               becomes a skinny GEMM: A is read once per iteration)
//     exchange u through an L2-resident buffer + ONE team barrier
//     every member redundantly applies the prox and the stop rule to the full vector
// All members execute bit-identical arithmetic on identical inputs, so they take the same
// convergence decisions without any further communication.  Reductions use fixed orders.
#include <algorithm>
#include "runtime.h"

namespace oemb200 {

constexpr int PK_THREADS = 256;
constexpr int PK_WARPS = PK_THREADS / 32;
constexpr int PK_CB = 4;          // chains per register batch in the mat-vec
constexpr int LZ_MAX = 768;       // Lanczos step cap
constexpr int PK_MAXCT = 32;      // chains per team cap

struct ChainDev {
    int gram, penalty, nlam, lam_off;
    double alpha, gamma, tau;
    int out_off, pad;
};

struct PathArgs {
    int q, ngram, team_size, cpc, max_ct, Lmax, maxit, accelerate, compute_eig, a_in_smem, ngroups, pad0;
    double tol, eig_factor, eig_tol;
    const double *XX, *XY;
    double *d, *Abuf;
    const ChainDev *chains;
    const int *team_ptr, *team_idx;   // CSR: chains of each team
    const double *lambdas, *pen_fact;
    const int *unique_groups, *grp_ptr, *grp_idx, *grp_cover;
    const double *group_weights, *post_scale, *beta_init;
    double *beta_final, *beta_out;
    int *niter_out, *lanczos_steps;
    double *ubuf;            // ngram x 2 x max_ct x q
    unsigned *barriers;      // ngram counters, zero-initialised
};

__device__ __forceinline__ double pk_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block-wide sum, result returned to every thread
__device__ __forceinline__ double block_sum(double v, double *red) {
    v = pk_warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < PK_WARPS; ++w) s += red[w];
    __syncthreads();
    return s;
}

__device__ __forceinline__ void team_barrier(unsigned *ctr, unsigned &target, int team_size) {
    __syncthreads();
    if (team_size > 1) {
        if (threadIdx.x == 0) {
            target += (unsigned)team_size;
            __threadfence();
            atomicAdd(ctr, 1u);
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            } while ((int)(v - target) < 0);
            __threadfence();
        }
        __syncthreads();
    }
}

// ---- thresholding family (coordinate-wise), operation order as in src/oem_dense.h:76-149 ----
__device__ __forceinline__ double st_lasso(double v, double pen, double d) {
    if (v > pen) return (v - pen) / d;
    if (v < -pen) return (v + pen) / d;
    return 0.0;
}
__device__ __forceinline__ double st_mcp(double v, double pen, double d, double gamma) {
    const double gammad = gamma * d, dmg = d - 1.0 / gamma;
    if (fabs(v) > gammad * pen) return v / d;
    if (v > pen) return (v - pen) / dmg;
    if (v < -pen) return (v + pen) / dmg;
    return 0.0;
}
__device__ __forceinline__ double st_scad(double v, double pen, double d, double gamma) {
    const double gammad = gamma * d, gm1d = (gamma - 1.0) * d;
    if (fabs(v) > gammad * pen) return v / d;
    if (fabs(v) > (d + 1.0) * pen) {
        const double gp = (gamma - 1.0) * v, gpen = gamma * pen;
        if (gp > gpen) return (gp - gpen) / (gm1d - 1.0);
        if (gp < -gpen) return (gp + gpen) / (gm1d - 1.0);
        return 0.0;
    }
    if (v > pen) return (v - pen) / d;
    if (v < -pen) return (v + pen) / d;
    return 0.0;
}
// group factors: src/oem_dense.h:151-191
__device__ __forceinline__ double scad_norm(double b, double pen, double d, double gamma) {
    const double gammad = gamma * d, gm1d = (gamma - 1.0) * d;
    if (fabs(b) > gammad * pen) return 1.0;
    if (fabs(b) > (d + 1.0) * pen) {
        const double gp = gamma - 1.0, gpen = gamma * pen / b;
        if (gp > gpen) return d * (gp - gpen) / (gm1d - 1.0);
        if (gp < -gpen) return d * (gp + gpen) / (gm1d - 1.0);
        return 0.0;
    }
    if (b > pen) return 1.0 - pen / b;
    if (b < -pen) return 1.0 + pen / b;
    return 0.0;
}
__device__ __forceinline__ double mcp_norm(double b, double pen, double d, double gamma) {
    const double gammad = gamma * d, dmg = d - 1.0 / gamma;
    if (fabs(b) > gammad * pen) return 1.0;
    if (b > pen) return d * (1.0 - pen / b) / dmg;
    if (b < -pen) return d * (1.0 + pen / b) / dmg;
    return 0.0;
}

// prox of one chain, in place on v[0..q): on entry v = u, on exit v = next beta.
// Dispatch: src/oem_dense.h:527-629.
__device__ void prox_inplace(const PathArgs &a, const ChainDev &ch, double lambda, double d, double *v) {
    const int q = a.q;
    const double alpha = ch.alpha, gamma = ch.gamma;
    double denom = d + (1.0 - alpha) * lambda;
    double lam = lambda * alpha;
    const int pen = ch.penalty;
    if (pen == OEMB200_PEN_SCAD_NET && alpha == 0.0) { lam = 0.0; denom = d + lambda; }
    const bool net = (pen == OEMB200_PEN_ENET || pen == OEMB200_PEN_SCAD_NET || pen == OEMB200_PEN_MCP_NET ||
                      pen == OEMB200_PEN_GRP_LASSO_NET || pen == OEMB200_PEN_GRP_MCP_NET ||
                      pen == OEMB200_PEN_GRP_SCAD_NET);
    const double lp = net ? lam : lambda, dp = net ? denom : d;
    if (pen < OEMB200_PEN_GRP_LASSO) {
        for (int j = threadIdx.x; j < q; j += PK_THREADS) {
            const double u = v[j];
            const double tp = __ldg(a.pen_fact + j) * lp;
            double r;
            switch (pen) {
                case OEMB200_PEN_OLS: r = u / d; break;
                case OEMB200_PEN_SCAD: case OEMB200_PEN_SCAD_NET: r = st_scad(u, tp, dp, gamma); break;
                case OEMB200_PEN_MCP: case OEMB200_PEN_MCP_NET: r = st_mcp(u, tp, dp, gamma); break;
                default: r = st_lasso(u, tp, dp); break;   // lasso, elastic.net
            }
            v[j] = r;
        }
        __syncthreads();
        return;
    }
    double glam = lp, gd = dp;
    int kind = 0;
    if (pen == OEMB200_PEN_GRP_MCP || pen == OEMB200_PEN_GRP_MCP_NET) kind = 1;
    if (pen == OEMB200_PEN_GRP_SCAD || pen == OEMB200_PEN_GRP_SCAD_NET) kind = 2;
    if (pen == OEMB200_PEN_SPARSE_GRP_LASSO) {
        const double lam_l1 = ch.tau * lambda;
        glam = (1.0 - ch.tau) * lambda;
        gd = d;
        for (int j = threadIdx.x; j < q; j += PK_THREADS) v[j] = st_lasso(v[j], __ldg(a.pen_fact + j) * lam_l1, 1.0);
        __syncthreads();
    }
    // one thread per group: sequential norm in member order like block_soft_threshold (src/oem_dense.h:193-315)
    for (int g = threadIdx.x; g < a.ngroups; g += PK_THREADS) {
        const int b0 = a.grp_ptr[g], b1 = a.grp_ptr[g + 1];
        double tf;
        if (a.unique_groups[g] == 0) tf = 1.0;
        else {
            double nrm = 0.0;
            for (int k = b0; k < b1; ++k) { const double x = v[a.grp_idx[k]]; nrm += x * x; }
            nrm = sqrt(nrm);
            const double gw = a.group_weights[g];
            if (kind == 0) { const double t = 1.0 - glam * gw / nrm; tf = (0.0 < t) ? t : 0.0; }
            else if (kind == 1) tf = mcp_norm(nrm, glam * gw, gd, gamma);
            else tf = scad_norm(nrm, glam * gw, gd, gamma);
        }
        for (int k = b0; k < b1; ++k) {
            const int c = a.grp_idx[k];
            v[c] = (tf != 0.0) ? v[c] * tf / gd : 0.0;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < q; j += PK_THREADS)
        if (!a.grp_cover[j]) v[j] = 0.0;      // variables in no listed group stay 0 (res.setZero())
    __syncthreads();
}

// Largest eigenvalue of the k x k symmetric tridiagonal (al, be) by 32-way multisection on the
// Sturm count, then the backward eigenvector recurrence for the residual bound
// be[k-1] * |s_k| / ||s||.  Executed by warp 0; out[0] = theta, out[1] = |s_k| / ||s||.
__device__ void tridiag_top(const double *al, const double *be, int k, double *out) {
    const int lane = threadIdx.x & 31;
    double lo = -1e300, hi = -1e300;
    for (int i = lane; i < k; i += 32) {
        const double r = (i > 0 ? fabs(be[i - 1]) : 0.0) + (i + 1 < k ? fabs(be[i]) : 0.0);
        lo = fmax(lo, al[i]);
        hi = fmax(hi, al[i] + r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmax(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    hi += 1e-14 * fabs(hi) + 1e-300;
    for (int round = 0; round < 14; ++round) {
        const double w = hi - lo;
        if (!(w > 2e-16 * fmax(fabs(hi), fabs(lo)))) break;
        const double x = lo + w * (double)(lane + 1) / 33.0;
        // all eigenvalues < x  <=>  every Sturm pivot negative
        bool all_below = true;
        double qv = al[0] - x;
        if (qv >= 0.0) all_below = false;
        for (int i = 1; i < k && all_below; ++i) {
            qv = al[i] - x - be[i - 1] * be[i - 1] / qv;
            if (qv >= 0.0) all_below = false;
        }
        const unsigned m = __ballot_sync(0xffffffffu, all_below);
        if (m == 0u) { lo = __shfl_sync(0xffffffffu, x, 31); }
        else {
            const int f = __ffs(m) - 1;
            const double nh = __shfl_sync(0xffffffffu, x, f);
            const double nl = __shfl_sync(0xffffffffu, x, f > 0 ? f - 1 : 0);
            hi = nh;
            if (f > 0) lo = nl;
        }
    }
    if (lane == 0) {
        const double theta = 0.5 * (lo + hi);
        // backward recurrence from s_k = 1 (the growing, hence stable, direction)
        double s_next = 0.0, s_cur = 1.0, nrm2 = 1.0, s_last = 1.0;   // s_last = s_k in current units
        for (int i = k - 1; i >= 1; --i) {
            // row i: be[i-1] s_{i-1} + (al[i]-theta) s_i + be[i] s_{i+1} = 0
            const double bi = (i + 1 < k) ? be[i] : 0.0;
            const double s_prev = -((al[i] - theta) * s_cur + bi * s_next) / be[i - 1];
            s_next = s_cur;
            s_cur = s_prev;
            nrm2 += s_cur * s_cur;
            if (fabs(s_cur) > 1e120) { s_cur *= 1e-120; s_next *= 1e-120; nrm2 *= 1e-240; s_last *= 1e-120; }
        }
        out[0] = theta;
        out[1] = s_last / sqrt(nrm2);   // |last eigenvector component| of the unit Ritz vector
    }
}

__global__ void __launch_bounds__(PK_THREADS, 1) oem_path_kernel(const PathArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int q = a.q;
    const int team = blockIdx.x / a.team_size, rank = blockIdx.x - team * a.team_size;
    const int c0 = min(q, rank * a.cpc), c1 = min(q, c0 + a.cpc);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ct0 = a.team_ptr[team], nct = a.team_ptr[team + 1] - ct0;
    const int nvec = max(a.max_ct, 2);

    double *Asl = sm;
    double *beta = sm + (a.a_in_smem ? (size_t)a.cpc * q : 0);
    double *us = beta + (size_t)nvec * q;
    double *red = us + (size_t)nvec * q;           // 16
    double *lz_al = red + 16;                      // LZ_MAX
    double *lz_be = lz_al + LZ_MAX;                // LZ_MAX
    double *ak = lz_be + LZ_MAX;                   // PK_MAXCT
    double *misc = ak + PK_MAXCT;                  // 8
    int *lam_idx = reinterpret_cast<int *>(misc + 8);   // PK_MAXCT each
    int *iter = lam_idx + PK_MAXCT;
    int *done = iter + PK_MAXCT;
    int *flag = done + PK_MAXCT;

    unsigned *bar = a.barriers + team;
    unsigned bar_target = 0;
    double *ub = a.ubuf + (size_t)team * 2 * a.max_ct * q;
    const double *XXg = a.XX + (size_t)team * q * q;
    double *Aglob = a.Abuf ? a.Abuf + (size_t)team * q * q : nullptr;
    const double *XYg = a.XY + (size_t)team * q;

    // ---- load my column slice of XX ----
    for (int j = c0 + warp; j < c1; j += PK_WARPS) {
        double *dst = a.a_in_smem ? Asl + (size_t)(j - c0) * q : Aglob + (size_t)j * q;
        const double *src = XXg + (size_t)j * q;
        for (int i = lane; i < q; i += 32) dst[i] = src[i];
    }
    __syncthreads();

    // mat-vec over the owned columns for `nv` vectors stored at vec + c*q (c < nv, active[c] != 0):
    // ubuf[par][c][j] = sign * sum_i S[i][j] vec_c[i] + add[j]
    auto matvec = [&](const double *vec, int nv, const int *inactive, const double *add, int par) {
        for (int j = c0 + warp; j < c1; j += PK_WARPS) {
            const double *col = a.a_in_smem ? Asl + (size_t)(j - c0) * q : Aglob + (size_t)j * q;
            for (int cb = 0; cb < nv; cb += PK_CB) {
                double acc[PK_CB];
#pragma unroll
                for (int c = 0; c < PK_CB; ++c) acc[c] = 0.0;
                for (int i = lane; i < q; i += 32) {
                    const double av = col[i];
#pragma unroll
                    for (int c = 0; c < PK_CB; ++c)
                        if (cb + c < nv) acc[c] = fma(av, vec[(size_t)(cb + c) * q + i], acc[c]);
                }
#pragma unroll
                for (int c = 0; c < PK_CB; ++c) {
                    if (cb + c < nv) {
                        const double s = pk_warp_sum(acc[c]);
                        if (lane == 0 && !(inactive && inactive[cb + c]))
                            ub[((size_t)par * a.max_ct + cb + c) * q + j] = s + (add ? add[j] : 0.0);
                    }
                }
            }
        }
    };

    // =========================== phase 0: top eigenvalue ===========================
    double dval;
    if (a.compute_eig) {
        double *v = beta, *vprev = beta + q, *w = us;
        // deterministic pseudo-random start vector (fixed seed), normalised
        double ss = 0.0;
        for (int i = threadIdx.x; i < q; i += PK_THREADS) {
            unsigned h = (unsigned)i * 2654435761u + 0x9E3779B9u;
            h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
            const double x = (double)(h & 0xFFFFFF) / 16777216.0 + 0.25;
            v[i] = x; vprev[i] = 0.0;
            ss += x * x;
        }
        ss = block_sum(ss, red);
        const double inv = 1.0 / sqrt(ss);
        for (int i = threadIdx.x; i < q; i += PK_THREADS) v[i] *= inv;
        __syncthreads();
        double beta_prev = 0.0, theta = 0.0;
        int k = 0, par = 0;
        const int kmax = min(LZ_MAX, max(q, 1));
        bool conv = false;
        while (!conv) {
            matvec(v, 1, nullptr, nullptr, par);
            team_barrier(bar, bar_target, a.team_size);
            const double *src = ub + (size_t)par * a.max_ct * q;
            for (int i = threadIdx.x; i < q; i += PK_THREADS) w[i] = ld_cg(src + i);
            par ^= 1;
            __syncthreads();
            double dot = 0.0;
            for (int i = threadIdx.x; i < q; i += PK_THREADS) dot = fma(v[i], w[i], dot);
            const double alpha_k = block_sum(dot, red);
            double nn = 0.0;
            for (int i = threadIdx.x; i < q; i += PK_THREADS) {
                const double x = w[i] - alpha_k * v[i] - beta_prev * vprev[i];
                w[i] = x;
                nn = fma(x, x, nn);
            }
            const double beta_k = sqrt(block_sum(nn, red));
            if (threadIdx.x == 0) { lz_al[k] = alpha_k; lz_be[k] = beta_k; }
            ++k;
            __syncthreads();
            // convergence check (every step while small, then every 4th)
            const bool breakdown = !(beta_k > 1e-14 * fabs(alpha_k));
            if (k <= 8 || (k & 3) == 0 || breakdown || k >= kmax) {
                if (warp == 0) tridiag_top(lz_al, lz_be, k, misc);
                __syncthreads();
                theta = misc[0];
                const double res = beta_k * misc[1];
                conv = breakdown || k >= kmax || (res <= a.eig_tol * fabs(theta));
                __syncthreads();
            }
            if (!conv) {
                const double ib = 1.0 / beta_k;
                for (int i = threadIdx.x; i < q; i += PK_THREADS) {
                    vprev[i] = v[i];
                    v[i] = w[i] * ib;
                }
                beta_prev = beta_k;
                __syncthreads();
            }
        }
        dval = theta * a.eig_factor;
        if (rank == 0 && threadIdx.x == 0) {
            a.d[team] = dval;
            if (a.lanczos_steps) a.lanczos_steps[team] = k;
        }
    } else {
        dval = a.d[team];
    }

    if (nct == 0) return;

    // ---- A = d I - XX on my slice ----
    for (int j = c0 + warp; j < c1; j += PK_WARPS) {
        double *col = a.a_in_smem ? Asl + (size_t)(j - c0) * q : Aglob + (size_t)j * q;
        for (int i = lane; i < q; i += 32) {
            const double x = -col[i];
            col[i] = (i == j) ? x + dval : x;
        }
    }
    // ---- chain state ----
    for (int c = threadIdx.x; c < nct; c += PK_THREADS) {
        lam_idx[c] = 0; iter[c] = 0; ak[c] = 1.0;
        done[c] = (a.chains[a.team_idx[ct0 + c]].nlam <= 0) ? 1 : 0;
    }
    for (int e = threadIdx.x; e < nct * q; e += PK_THREADS) {
        const int c = e / q, j = e - c * q;
        const int gc = a.team_idx[ct0 + c];
        beta[e] = a.beta_init ? a.beta_init[(size_t)a.chains[gc].out_off * q + j] : 0.0;
    }
    __syncthreads();

    // =========================== phase 1: lambda paths ===========================
    int par = 0;
    for (;;) {
        int nactive = 0;
        for (int c = 0; c < nct; ++c) nactive += done[c] ? 0 : 1;
        if (nactive == 0) break;
        matvec(beta, nct, done, XYg, par);
        team_barrier(bar, bar_target, a.team_size);
        const double *src = ub + (size_t)par * a.max_ct * q;
        for (int e = threadIdx.x; e < nct * q; e += PK_THREADS) {
            const int c = e / q;
            if (!done[c]) us[e] = ld_cg(src + e);
        }
        if (threadIdx.x < nct) flag[threadIdx.x] = 1;
        par ^= 1;
        __syncthreads();
        for (int c = 0; c < nct; ++c) {
            if (done[c]) continue;
            const ChainDev ch = a.chains[a.team_idx[ct0 + c]];
            const double lambda = a.lambdas[ch.lam_off + lam_idx[c]];
            double *bn = us + (size_t)c * q;
            double *bo = beta + (size_t)c * q;
            prox_inplace(a, ch, lambda, dval, bn);
            if (a.accelerate) {     // src/oem_dense.h:633-651
                const double ak_prev = ak[c];
                const double ak_new = 0.5 * (1.0 + sqrt(1.0 + 4.0 * ak_prev * ak_prev));
                const double ratio = (ak_prev - 1.0) / ak_new;
                double adv = 0.0;
                for (int j = threadIdx.x; j < q; j += PK_THREADS) {
                    const double upd = bn[j];
                    const double diff = upd - bo[j];
                    const double acc = upd + ratio * diff;
                    bn[j] = acc;
                    adv += (acc - upd) * diff;
                }
                adv = block_sum(adv, red);
                if (threadIdx.x == 0) ak[c] = (adv > 0.0) ? 1.0 : ak_new;
            }
            // stop rule (src/utils.cpp:537-549)
            bool ok = true;
            for (int j = threadIdx.x; j < q; j += PK_THREADS) {
                const double cur = bn[j], prev = bo[j];
                const double ac = fabs(cur), ap = fabs(prev);
                if ((ac > 1e-13 && ap <= 1e-13) || (ac <= 1e-13 && ap > 1e-13)) ok = false;
                if (ac > 1e-13 && ap > 1e-13 && fabs((cur - prev) / prev) > a.tol) ok = false;
            }
            if (!ok) flag[c] = 0;
        }
        __syncthreads();
        for (int c = 0; c < nct; ++c) {
            if (done[c]) continue;       // uniform: done[] only changes below, after the sync
            const ChainDev ch = a.chains[a.team_idx[ct0 + c]];
            double *bn = us + (size_t)c * q;
            double *bo = beta + (size_t)c * q;
            const int it = iter[c] + 1;
            const bool finished = flag[c] || it >= a.maxit;
            const int li = lam_idx[c];
            if (finished && a.post_scale) {      // oem_xtx get_beta() quirk: src/oem_xtx.h:576-581
                for (int j = threadIdx.x; j < q; j += PK_THREADS) bn[j] *= a.post_scale[j];
            }
            for (int j = threadIdx.x; j < q; j += PK_THREADS) {
                const double x = bn[j];
                bo[j] = x;
                if (finished && rank == 0) a.beta_out[((size_t)ch.out_off * a.Lmax + li) * q + j] = x;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                if (finished) {
                    if (rank == 0) a.niter_out[(size_t)ch.out_off * a.Lmax + li] = flag[c] ? it : a.maxit + 1;
                    iter[c] = 0;
                    lam_idx[c] = li + 1;
                    if (li + 1 >= ch.nlam) done[c] = 1;
                } else {
                    iter[c] = it;
                }
            }
        }
        __syncthreads();
    }
    if (a.beta_final && rank == 0) {
        for (int e = threadIdx.x; e < nct * q; e += PK_THREADS) {
            const int c = e / q, j = e - c * q;
            a.beta_final[(size_t)a.chains[a.team_idx[ct0 + c]].out_off * q + j] = beta[e];
        }
    }
}

// ---------------------------------------------------------------------------------------------
void path_launch(Ctx &cx, const PathProblem &pp) {
    const int q = pp.q, G = pp.ngram;
    if (q <= 0 || G <= 0) fail(OEMB200_EINVAL, "path: empty problem");
    // chains per team
    std::vector<int> tptr(G + 1, 0), tidx(pp.chains.size());
    for (auto &c : pp.chains) {
        if (c.gram < 0 || c.gram >= G) fail(OEMB200_EINVAL, "path: chain refers to Gram %d of %d", c.gram, G);
        tptr[c.gram + 1]++;
    }
    int max_ct = 1;
    for (int g = 0; g < G; ++g) { max_ct = std::max(max_ct, tptr[g + 1]); tptr[g + 1] += tptr[g]; }
    if (max_ct > PK_MAXCT) fail(OEMB200_EUNSUPPORTED, "path: more than %d chains on one Gram", PK_MAXCT);
    {
        std::vector<int> fill(tptr.begin(), tptr.end() - 1);
        for (size_t i = 0; i < pp.chains.size(); ++i) tidx[fill[pp.chains[i].gram]++] = (int)i;
    }
    std::vector<ChainDev> cd(pp.chains.size());
    for (size_t i = 0; i < pp.chains.size(); ++i) {
        const ChainDesc &c = pp.chains[i];
        cd[i] = ChainDev{c.gram, c.penalty, c.nlam, c.lam_off, c.alpha, c.gamma, c.tau, c.out_off, 0};
    }

    // ---- geometry: team size, column slice, shared memory ----
    auto kern = oem_path_kernel;
    const int nvec = std::max(max_ct, 2);
    const size_t fixed_bytes = ((size_t)2 * nvec * q + 16 + 2 * LZ_MAX + PK_MAXCT + 8) * 8 + 4 * PK_MAXCT * 4;
    const size_t smem_cap = cx.smem_optin;
    if (fixed_bytes > smem_cap) fail(OEMB200_EUNSUPPORTED, "path: q=%d with %d chains per Gram exceeds shared memory", q, max_ct);
    OEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap));
    int max_blocks_per_sm = 0;
    OEM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks_per_sm, kern, PK_THREADS, smem_cap));
    const int max_ctas = std::max(1, max_blocks_per_sm) * cx.num_sms;
    if (G > max_ctas) fail(OEMB200_EUNSUPPORTED, "path: %d Grams exceed the %d co-resident CTAs", G, max_ctas);
    int team = max_ctas / G;
    team = std::min(team, (q + 3) / 4);                       // at least 4 columns per member
    if ((size_t)q * q * 8 + fixed_bytes <= smem_cap) team = 1; // whole A fits in one CTA: no global barrier
    team = std::max(team, 1);
    int cpc = (q + team - 1) / team;
    team = (q + cpc - 1) / cpc;                               // drop members that would own nothing
    const bool a_in_smem = (size_t)cpc * q * 8 + fixed_bytes <= smem_cap;
    const size_t smem_bytes = fixed_bytes + (a_in_smem ? (size_t)cpc * q * 8 : 0);

    DBuf<ChainDev> d_chains(std::max<size_t>(1, cd.size()));
    DBuf<int> d_tptr(tptr.size()), d_tidx(std::max<size_t>(1, tidx.size()));
    DBuf<double> d_ubuf((size_t)G * 2 * max_ct * q);
    DBuf<unsigned> d_bar(G);
    DBuf<double> d_A;
    if (!a_in_smem) d_A.alloc((size_t)G * q * q);
    if (!cd.empty()) d_chains.upload(cd.data(), cd.size(), cx.stream);
    d_tptr.upload(tptr.data(), tptr.size(), cx.stream);
    if (!tidx.empty()) d_tidx.upload(tidx.data(), tidx.size(), cx.stream);
    d_bar.zero(cx.stream);

    PathArgs a;
    memset(&a, 0, sizeof a);
    a.q = q; a.ngram = G; a.team_size = team; a.cpc = cpc; a.max_ct = max_ct; a.Lmax = pp.Lmax;
    a.maxit = pp.maxit; a.accelerate = pp.accelerate ? 1 : 0; a.compute_eig = pp.compute_eig ? 1 : 0;
    a.a_in_smem = a_in_smem ? 1 : 0; a.ngroups = pp.ngroups;
    a.tol = pp.tol; a.eig_factor = pp.eig_factor; a.eig_tol = pp.eig_tol;
    a.XX = pp.XX; a.XY = pp.XY; a.d = pp.d; a.Abuf = d_A.p;
    a.chains = d_chains.p; a.team_ptr = d_tptr.p; a.team_idx = d_tidx.p;
    a.lambdas = pp.lambdas; a.pen_fact = pp.pen_fact;
    a.unique_groups = pp.unique_groups; a.grp_ptr = pp.grp_ptr; a.grp_idx = pp.grp_idx; a.grp_cover = pp.grp_cover;
    a.group_weights = pp.group_weights; a.post_scale = pp.post_scale; a.beta_init = pp.beta_init;
    a.beta_final = pp.beta_final; a.beta_out = pp.beta_out; a.niter_out = pp.niter_out;
    a.lanczos_steps = pp.lanczos_steps; a.ubuf = d_ubuf.p; a.barriers = d_bar.p;

    void *kargs[] = {&a};
    OEM_CUDA(cudaLaunchCooperativeKernel((void *)kern, dim3(G * team), dim3(PK_THREADS), kargs, smem_bytes, cx.stream));
    cx.st.kernel_launches += 1;
}

}  // namespace oemb200
