// logit_fused.cu -- one IRLS data pass of the logistic path with X read from HBM ONCE:
//     eta = X b + b0;  prob = 1 / (1 + exp(-eta));  W = prob (1 - prob);  r = y - prob      (phase 1)
//     grad_j = sum_i x_ij r_i,  grad_0 = sum_i r_i                                            (phase 2)
// replacing the two sweeps of oemLogisticDense::solve (src/oem_logistic_dense.h:864-949 and :970-992).
//
// A persistent CTA walks 32-row blocks of the column-major matrix.  For each block it streams the 32 x p
// slab through a 4-stage TMA ring twice, back to back: the first sweep forms eta (and prob / W / r), the
// second sweep forms the block's contribution to X'r.  The slab of a block is 32 * p * 8 bytes (256 KB at
// p = 1000); with 2 CTAs per SM about 76 MB are "between sweeps" at any time, so the second sweep is served by
// the 126 MB L2 (first-sweep loads carry an evict_last hint, second-sweep loads evict_first) and HBM sees each
// byte of X once.  Column sums are accumulated per CTA in shared memory (each column is owned by one warp per
// box: no atomics) and reduced across CTAs in fixed order by a second tiny kernel: bit-reproducible.
#include <algorithm>
#include "runtime.h"

namespace oemb200 {

constexpr int LF_ROWS = 32;
constexpr int LF_COLS = 64;
constexpr int LF_STAGES = 4;
constexpr int LF_THREADS = 256;
constexpr int LF_BOX_BYTES = LF_ROWS * LF_COLS * 8;      // 16 KB

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_2d_hint(void *dst_smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}

__global__ void __launch_bounds__(LF_THREADS, 2)
logit_fused_kernel(const __grid_constant__ CUtensorMap tmap, long long n, int p, const double *__restrict__ b, double b0,
                   const double *__restrict__ y, double *__restrict__ prob, double *__restrict__ wout,
                   double *__restrict__ partial /* gridDim.x x (p + 1) */) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *boxes = reinterpret_cast<double *>(smem_raw);                       // LF_STAGES x 32 x 64
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + LF_STAGES * LF_BOX_BYTES);
    uint64_t *empty = full + LF_STAGES;
    double *bs = reinterpret_cast<double *>(smem_raw + LF_STAGES * LF_BOX_BYTES + 128);   // pcols (padded to 64)
    const int ncb = (p + LF_COLS - 1) / LF_COLS;
    const int pcols = ncb * LF_COLS;
    double *gacc = bs + pcols;                                                  // pcols
    double *etap = gacc + pcols;                                                // 16 x 32 partial eta
    double *rsm = etap + 16 * LF_ROWS;                                          // 32
    double *rsum = rsm + LF_ROWS;                                               // 1 (+pad)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long nblk = (n + LF_ROWS - 1) / LF_ROWS;
    const long long myblk = (nblk > blockIdx.x) ? (nblk - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total = myblk * 2 * ncb;                                    // boxes this CTA streams

    for (int j = threadIdx.x; j < pcols; j += LF_THREADS) { bs[j] = j < p ? b[j] : 0.0; gacc[j] = 0.0; }
    if (threadIdx.x == 0) {
        for (int s = 0; s < LF_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], LF_THREADS / 32); }
        mbar_fence_init();
        tma_prefetch_desc(&tmap);
        rsum[0] = 0.0;
    }
    __syncthreads();

    uint64_t pol_keep = 0, pol_drop = 0;
    if (threadIdx.x == 0) { pol_keep = l2_policy_evict_last(); pol_drop = l2_policy_evict_first(); }
    auto issue = [&](long long idx) {      // thread 0 only
        const int s = (int)(idx % LF_STAGES);
        const long long kb = idx / (2 * ncb);
        const int rem = (int)(idx - kb * 2 * ncb);
        const int phase = rem / ncb, cb = rem - phase * ncb;
        const long long blk = blockIdx.x + kb * gridDim.x;
        mbar_arrive_expect_tx(&full[s], LF_BOX_BYTES);
        tma_load_2d_hint(boxes + (size_t)s * (LF_ROWS * LF_COLS), &tmap, &full[s], (int)(blk * LF_ROWS), cb * LF_COLS,
                         phase == 0 ? pol_keep : pol_drop);
    };
    if (threadIdx.x == 0)
        for (long long i = 0; i < LF_STAGES && i < total; ++i) issue(i);

    const int rp = threadIdx.x & 15, cg = threadIdx.x >> 4;      // phase 1: row pair x 4-column group (128-bit loads)
    const int lrp = lane & 15, lh = lane >> 4;                   // phase 2: row pair x half of the warp's 8 columns
    long long idx = 0;
    for (long long kb = 0; kb < myblk; ++kb) {
        const long long blk = blockIdx.x + kb * gridDim.x;
        const long long r0 = blk * LF_ROWS;
        // ------------------------- phase 1: eta for the 32 rows -------------------------
        double acc0 = 0.0, acc1 = 0.0;
        for (int cb = 0; cb < ncb; ++cb, ++idx) {
            const int s = (int)(idx % LF_STAGES);
            const uint32_t ph = (uint32_t)((idx / LF_STAGES) & 1);
            mbar_wait(&full[s], ph);
            const double2 *bx = reinterpret_cast<const double2 *>(boxes + (size_t)s * (LF_ROWS * LF_COLS) +
                                                                  (size_t)(cg * 4) * LF_ROWS) + rp;
            const double2 *bb = reinterpret_cast<const double2 *>(bs + cb * LF_COLS + cg * 4);
            const double2 b01 = bb[0], b23 = bb[1];
            const double2 x0 = bx[0], x1 = bx[LF_ROWS / 2], x2 = bx[LF_ROWS], x3 = bx[3 * LF_ROWS / 2];
            acc0 = fma(x0.x, b01.x, acc0); acc1 = fma(x0.y, b01.x, acc1);
            acc0 = fma(x1.x, b01.y, acc0); acc1 = fma(x1.y, b01.y, acc1);
            acc0 = fma(x2.x, b23.x, acc0); acc1 = fma(x2.y, b23.x, acc1);
            acc0 = fma(x3.x, b23.y, acc0); acc1 = fma(x3.y, b23.y, acc1);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (threadIdx.x == 0 && idx + LF_STAGES < total) { mbar_wait(&empty[s], ph); issue(idx + LF_STAGES); }
        }
        etap[cg * LF_ROWS + 2 * rp] = acc0;
        etap[cg * LF_ROWS + 2 * rp + 1] = acc1;
        __syncthreads();
        if (threadIdx.x < LF_ROWS) {
            double e = b0;
#pragma unroll
            for (int k = 0; k < 16; ++k) e += etap[k * LF_ROWS + threadIdx.x];
            const long long i = r0 + threadIdx.x;
            double r = 0.0;
            if (i < n) {
                const double pr = 1.0 / (1.0 + exp(-e));
                r = y[i] - pr;
                if (prob) prob[i] = pr;
                if (wout) wout[i] = pr * (1.0 - pr);
            }
            rsm[threadIdx.x] = r;
            // block sum of r in lane order (fixed order)
            double sr = r;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sr += __shfl_xor_sync(0xffffffffu, sr, o);
            if (threadIdx.x == 0) rsum[0] += sr;
        }
        __syncthreads();
        // ------------------------- phase 2: X' r for the block -------------------------
        const double2 rl = reinterpret_cast<const double2 *>(rsm)[lrp];
        for (int cb = 0; cb < ncb; ++cb, ++idx) {
            const int s = (int)(idx % LF_STAGES);
            const uint32_t ph = (uint32_t)((idx / LF_STAGES) & 1);
            mbar_wait(&full[s], ph);
            const double2 *bx = reinterpret_cast<const double2 *>(boxes + (size_t)s * (LF_ROWS * LF_COLS) +
                                                                  (size_t)(warp * 8 + lh * 4) * LF_ROWS) + lrp;
            double v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double2 x = bx[c * (LF_ROWS / 2)];
                v[c] = fma(x.y, rl.y, x.x * rl.x);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (threadIdx.x == 0 && idx + LF_STAGES < total) { mbar_wait(&empty[s], ph); issue(idx + LF_STAGES); }
            // transposed reduction over the 16 row pairs: 4 column sums with 2+1+2 shuffles
            const bool h8 = lane & 8, h4 = lane & 4;
            double w2[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const double send = h8 ? v[k] : v[k + 2];
                const double keep = h8 ? v[k + 2] : v[k];
                w2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            double s1 = (h4 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, h4 ? w2[0] : w2[1], 4);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
            if ((lane & 3) == 0) {
                const int col = lh * 4 + (h8 ? 2 : 0) + (h4 ? 1 : 0);
                gacc[cb * LF_COLS + warp * 8 + col] += s1;       // this warp owns these 8 columns of the box
            }
        }
    }
    __syncthreads();
    double *out = partial + (size_t)blockIdx.x * (p + 1);
    for (int j = threadIdx.x; j < p; j += LF_THREADS) out[1 + j] = gacc[j];
    if (threadIdx.x == 0) out[0] = rsum[0];
}

__global__ void lf_sum_partials_kernel(const double *__restrict__ partial, int nparts, int width, double *__restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= width) return;
    double s = 0.0;
    for (int c = 0; c < nparts; ++c) s += partial[(size_t)c * width + k];
    out[k] = s;
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool logit_fused_supported(const double *X, int64_t n, int p, int64_t ld) {
    return (ld % 2 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && n < (1ll << 31) && p <= 6000;
}

// grad_out[0] = sum r, grad_out[1 + j] = sum_i x_ij r_i (device, p + 1 doubles)
void logit_fused_launch(Ctx &cx, const double *X, int64_t n, int p, int64_t ld, const double *b, double b0,
                        const double *y, double *prob, double *w, double *grad_out) {
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        fail(OEMB200_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)p};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {LF_ROWS, LF_COLS};
    cuuint32_t es[2] = {1, 1};
    if (reinterpret_cast<EncodeTiledFn2>(fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(X), dims, strides,
                                             box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        fail(OEMB200_ECUDA, "logit_fused: tensor map failed");
    const int ncb = (p + LF_COLS - 1) / LF_COLS, pcols = ncb * LF_COLS;
    const size_t smem = LF_STAGES * LF_BOX_BYTES + 128 + ((size_t)2 * pcols + 16 * LF_ROWS + LF_ROWS + 2) * 8;
    OEM_CUDA(cudaFuncSetAttribute(logit_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    OEM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, logit_fused_kernel, LF_THREADS, smem));
    const int64_t nblk = (n + LF_ROWS - 1) / LF_ROWS;
    const int grid = (int)std::min<int64_t>(nblk, (int64_t)std::max(1, per_sm) * cx.num_sms);
    DBuf<double> partial((size_t)grid * (p + 1));
    logit_fused_kernel<<<grid, LF_THREADS, smem, cx.stream>>>(tm, n, p, b, b0, y, prob, w, partial.p);
    OEM_CUDA(cudaGetLastError());
    lf_sum_partials_kernel<<<(p + 1 + 255) / 256, 256, 0, cx.stream>>>(partial.p, grid, p + 1, grad_out);
    OEM_CUDA(cudaGetLastError());
    cx.st.kernel_launches += 2;
    cx.st.xb_launches += 1;
}

}  // namespace oemb200
