// colstats.cu -- the HBM-bound sweeps over the column-major n x p matrix X:
//
//   colstats_kernel   per column j:  sum_i xs_ij v0_i,  sum_i xs_ij v1_i,  sum_i xs_ij^2   (xs = x - shift_j)
//                     replaces the column loops of oemBig::init_oem (src/oem_big.h:743-837: colsq, X'y,
//                     colsums), oemLogisticDense::init_oem (src/oem_logistic_dense.h:734-792), the X'y
//                     GEMV of oemDense::init_oem (src/oem_dense.h:699-707), DataStd's mean / norm passes
//                     (src/DataStd.h:203-265) and the logistic gradient X'(y-p) (oem_logistic_dense.h:970-992).
//   xb_kernel         eta = X b + b0 with the logistic epilogue prob / y-prob / prob(1-prob)
//                     (src/oem_logistic_dense.h:864-949).
//   vec kernels       sums over y (DataStd.h:102-138).
//
// All are coalesced along rows (the contiguous index), 128-bit vectorised when the leading dimension
// and base allow it, and reduce in two fixed-order stages (per-CTA partials, then one thread per
// output) so that results are bit-reproducible run to run.
#include "runtime.h"

namespace oemb200 {

constexpr int CS_THREADS = 256;
constexpr int CS_COLS = 4;            // columns per CTA: v0/v1 are reused across them
constexpr int CS_ROWS = 32768;        // rows per CTA

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool VEC2>
__global__ void __launch_bounds__(CS_THREADS)
colstats_kernel(const double *__restrict__ X, long long n, int p, long long ld, const double *__restrict__ v0,
                const double *__restrict__ v1, const double *__restrict__ shift, double *__restrict__ partial) {
    const int j0 = blockIdx.x * CS_COLS;
    const long long r0 = (long long)blockIdx.y * CS_ROWS;
    const long long r1 = min(n, r0 + (long long)CS_ROWS);
    double a0[CS_COLS], a1[CS_COLS], a2[CS_COLS], sh[CS_COLS];
    const double *col[CS_COLS];
#pragma unroll
    for (int c = 0; c < CS_COLS; ++c) {
        a0[c] = a1[c] = a2[c] = 0.0;
        const int j = min(j0 + c, p - 1);
        col[c] = X + (size_t)j * ld;
        sh[c] = shift ? shift[j] : 0.0;
    }
    if (VEC2) {
        for (long long r = r0 + 2 * threadIdx.x; r < r1; r += 2 * CS_THREADS) {
            if (r + 1 < r1) {
                double2 w0 = v0 ? *reinterpret_cast<const double2 *>(v0 + r) : make_double2(1.0, 1.0);
                double2 w1 = v1 ? *reinterpret_cast<const double2 *>(v1 + r) : make_double2(1.0, 1.0);
#pragma unroll
                for (int c = 0; c < CS_COLS; ++c) {
                    double2 x = *reinterpret_cast<const double2 *>(col[c] + r);
                    x.x -= sh[c]; x.y -= sh[c];
                    a0[c] = fma(x.x, w0.x, a0[c]); a0[c] = fma(x.y, w0.y, a0[c]);
                    a1[c] = fma(x.x, w1.x, a1[c]); a1[c] = fma(x.y, w1.y, a1[c]);
                    a2[c] = fma(x.x, x.x, a2[c]); a2[c] = fma(x.y, x.y, a2[c]);
                }
            } else {
                const double w0 = v0 ? v0[r] : 1.0, w1 = v1 ? v1[r] : 1.0;
#pragma unroll
                for (int c = 0; c < CS_COLS; ++c) {
                    const double x = col[c][r] - sh[c];
                    a0[c] = fma(x, w0, a0[c]); a1[c] = fma(x, w1, a1[c]); a2[c] = fma(x, x, a2[c]);
                }
            }
        }
    } else {
        for (long long r = r0 + threadIdx.x; r < r1; r += CS_THREADS) {
            const double w0 = v0 ? v0[r] : 1.0, w1 = v1 ? v1[r] : 1.0;
#pragma unroll
            for (int c = 0; c < CS_COLS; ++c) {
                const double x = col[c][r] - sh[c];
                a0[c] = fma(x, w0, a0[c]); a1[c] = fma(x, w1, a1[c]); a2[c] = fma(x, x, a2[c]);
            }
        }
    }
    __shared__ double red[CS_THREADS / 32][3 * CS_COLS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < CS_COLS; ++c) {
        const double s0 = warp_sum(a0[c]), s1 = warp_sum(a1[c]), s2 = warp_sum(a2[c]);
        if (lane == 0) { red[warp][c] = s0; red[warp][CS_COLS + c] = s1; red[warp][2 * CS_COLS + c] = s2; }
    }
    __syncthreads();
    if (threadIdx.x < 3 * CS_COLS) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < CS_THREADS / 32; ++w) s += red[w][threadIdx.x];
        const int stat = threadIdx.x / CS_COLS, c = threadIdx.x % CS_COLS;
        if (j0 + c < p) partial[((size_t)blockIdx.y * 3 + stat) * p + j0 + c] = s;
    }
}

// out[k] (+)= sum over chunks of partial[chunk][k], one thread per output, fixed order
__global__ void sum_partials_kernel(const double *__restrict__ partial, int nchunks, int width, double *__restrict__ out,
                                    int accumulate) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= width) return;
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += partial[(size_t)c * width + k];
    out[k] = accumulate ? out[k] + s : s;
}

void colstats_launch(Ctx &cx, const double *X, int64_t n, int p, int64_t ld, const double *v0, const double *v1,
                     const double *shift, double *out3p, bool accumulate) {
    const int nchunks = (int)((n + CS_ROWS - 1) / CS_ROWS);
    DBuf<double> partial((size_t)nchunks * 3 * p);
    dim3 grid((p + CS_COLS - 1) / CS_COLS, nchunks);
    const bool vec2 = (ld % 2 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
                      (!v0 || (reinterpret_cast<uintptr_t>(v0) & 15) == 0) &&
                      (!v1 || (reinterpret_cast<uintptr_t>(v1) & 15) == 0);
    if (vec2) colstats_kernel<true><<<grid, CS_THREADS, 0, cx.stream>>>(X, n, p, ld, v0, v1, shift, partial.p);
    else colstats_kernel<false><<<grid, CS_THREADS, 0, cx.stream>>>(X, n, p, ld, v0, v1, shift, partial.p);
    OEM_CUDA(cudaGetLastError());
    sum_partials_kernel<<<(3 * p + 255) / 256, 256, 0, cx.stream>>>(partial.p, nchunks, 3 * p, out3p, accumulate ? 1 : 0);
    OEM_CUDA(cudaGetLastError());
    cx.st.kernel_launches += 2;
    cx.st.xtr_launches += 1;
}

// ---------------------------------------------------------------------------------------------
// vector sums: out[0] = sum (v - shift), out[1] = sum (v - shift)^2
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vecsum_kernel(const double *__restrict__ v, long long n, double shift, double *__restrict__ partial) {
    const long long r0 = (long long)blockIdx.x * CS_ROWS, r1 = min(n, r0 + (long long)CS_ROWS);
    double s = 0.0, ss = 0.0;
    for (long long r = r0 + threadIdx.x; r < r1; r += 256) {
        const double x = v[r] - shift;
        s += x;
        ss = fma(x, x, ss);
    }
    __shared__ double red[8][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    s = warp_sum(s); ss = warp_sum(ss);
    if (lane == 0) { red[warp][0] = s; red[warp][1] = ss; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * 2 + threadIdx.x] = t;
    }
}

void vecsum_launch(Ctx &cx, const double *v, int64_t n, double shift, double *out2, bool accumulate) {
    const int nchunks = (int)((n + CS_ROWS - 1) / CS_ROWS);
    DBuf<double> partial((size_t)nchunks * 2);
    vecsum_kernel<<<nchunks, 256, 0, cx.stream>>>(v, n, shift, partial.p);
    OEM_CUDA(cudaGetLastError());
    sum_partials_kernel<<<1, 32, 0, cx.stream>>>(partial.p, nchunks, 2, out2, accumulate ? 1 : 0);
    OEM_CUDA(cudaGetLastError());
    cx.st.kernel_launches += 2;
}

__global__ void affine_kernel(const double *__restrict__ v, long long n, double shift, double divisor,
                              double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (v[i] - shift) / divisor;
}

void affine_launch(Ctx &cx, const double *v, int64_t n, double shift, double divisor, double *out) {
    // (v - shift) / divisor, written as the reference does it: y.array() -= mean; y /= scale
    affine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, cx.stream>>>(v, n, shift, divisor, out);
    OEM_CUDA(cudaGetLastError());
    cx.st.kernel_launches += 1;
}

__global__ void axpy_kernel(long long n, double a, const double *__restrict__ x, double *__restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = fma(a, x[i], y[i]);
}
void axpy_launch(Ctx &cx, int64_t n, double a, const double *x, double *y) {
    axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, cx.stream>>>(n, a, x, y);
    OEM_CUDA(cudaGetLastError());
    cx.st.kernel_launches += 1;
}

// ---------------------------------------------------------------------------------------------
// eta = X b + b0, logistic epilogue.  Each thread owns two adjacent rows (one 128-bit load per
// column); b is staged in shared memory and read as a broadcast.
// ---------------------------------------------------------------------------------------------
constexpr int XB_THREADS = 256;

template <bool VEC2, bool LOGISTIC>
__global__ void __launch_bounds__(XB_THREADS)
xb_kernel(const double *__restrict__ X, long long n, int p, long long ld, const double *__restrict__ b, double b0,
          const double *__restrict__ y, double *__restrict__ eta, double *__restrict__ prob,
          double *__restrict__ resid, double *__restrict__ w, const double *__restrict__ b0_dev) {
    extern __shared__ double bs[];
    if (b0_dev) b0 += *b0_dev;
    for (int j = threadIdx.x; j < p; j += XB_THREADS) bs[j] = b[j];
    __syncthreads();
    const long long r = ((long long)blockIdx.x * XB_THREADS + threadIdx.x) * 2;
    if (r >= n) return;
    const bool two = (r + 1 < n);
    double e0 = 0.0, e1 = 0.0;
    const double *xr = X + r;
    if (VEC2 && two) {
        double f0 = 0.0, f1 = 0.0, g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0;
        int j = 0;
        for (; j + 4 <= p; j += 4) {
            const double2 x0 = *reinterpret_cast<const double2 *>(xr + (size_t)(j + 0) * ld);
            const double2 x1 = *reinterpret_cast<const double2 *>(xr + (size_t)(j + 1) * ld);
            const double2 x2 = *reinterpret_cast<const double2 *>(xr + (size_t)(j + 2) * ld);
            const double2 x3 = *reinterpret_cast<const double2 *>(xr + (size_t)(j + 3) * ld);
            const double b_0 = bs[j], b_1 = bs[j + 1], b_2 = bs[j + 2], b_3 = bs[j + 3];
            e0 = fma(x0.x, b_0, e0); e1 = fma(x0.y, b_0, e1);
            f0 = fma(x1.x, b_1, f0); f1 = fma(x1.y, b_1, f1);
            g0 = fma(x2.x, b_2, g0); g1 = fma(x2.y, b_2, g1);
            h0 = fma(x3.x, b_3, h0); h1 = fma(x3.y, b_3, h1);
        }
        for (; j < p; ++j) {
            const double2 x0 = *reinterpret_cast<const double2 *>(xr + (size_t)j * ld);
            e0 = fma(x0.x, bs[j], e0); e1 = fma(x0.y, bs[j], e1);
        }
        e0 = (e0 + f0) + (g0 + h0);
        e1 = (e1 + f1) + (g1 + h1);
    } else {
        for (int j = 0; j < p; ++j) {
            const double bj = bs[j];
            e0 = fma(xr[(size_t)j * ld], bj, e0);
            if (two) e1 = fma(xr[(size_t)j * ld + 1], bj, e1);
        }
    }
    e0 += b0; e1 += b0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (k == 1 && !two) break;
        const double e = k ? e1 : e0;
        const long long i = r + k;
        if (eta) eta[i] = e;
        if (LOGISTIC) {
            const double pr = 1.0 / (1.0 + exp(-e));
            if (prob) prob[i] = pr;
            if (resid) resid[i] = y[i] - pr;
            if (w) w[i] = pr * (1.0 - pr);
        }
    }
}

void xb_launch(Ctx &cx, const double *X, int64_t n, int p, int64_t ld, const double *b, double b0, const double *y,
               double *eta, double *prob, double *resid, double *w, bool logistic, const double *b0_dev) {
    const bool vec2 = (ld % 2 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
    const unsigned grid = (unsigned)((n + 2 * XB_THREADS - 1) / (2 * XB_THREADS));
    const size_t sm = (size_t)p * 8;
#define OEM_XB(V, L) xb_kernel<V, L><<<grid, XB_THREADS, sm, cx.stream>>>(X, n, p, ld, b, b0, y, eta, prob, resid, w, b0_dev)
    if (vec2 && logistic) OEM_XB(true, true);
    else if (vec2) OEM_XB(true, false);
    else if (logistic) OEM_XB(false, true);
    else OEM_XB(false, false);
#undef OEM_XB
    OEM_CUDA(cudaGetLastError());
    cx.st.kernel_launches += 1;
    cx.st.xb_launches += 1;
}

}  // namespace oemb200
