// Microbenchmark (B200): latency / issue cost of FP64 CUDA-core ops, DMMA, LDS and bar.sync with 8 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(double *out, long long *cyc, int n) {
    __shared__ double sm[1024];
    double x = threadIdx.x * 1e-3 + 1.0, y = 1.0000001, z = 0.5;
    sm[threadIdx.x] = x;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = fma(x, y, z);                       // dependent DFMA chain
    long long t1 = clock64();
    double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
    for (int i = 0; i < n; ++i) { a0 = fma(a0, y, z); a1 = fma(a1, y, z); a2 = fma(a2, y, z); a3 = fma(a3, y, z);
                                  a4 = fma(a4, y, z); a5 = fma(a5, y, z); a6 = fma(a6, y, z); a7 = fma(a7, y, z); }   // 8 independent
    long long t2 = clock64();
    double d = x;
    for (int i = 0; i < n; ++i) d = z / (d + 1.5);                      // dependent division
    long long t3 = clock64();
    double c0 = 0, c1 = 0;
    for (int i = 0; i < n; ++i) dmma(c0, c1, x, y);                     // dependent DMMA chain
    long long t4 = clock64();
    double e0 = 0, e1 = 0, f0 = 0, f1 = 0, g0 = 0, g1 = 0, h0 = 0, h1 = 0;
    for (int i = 0; i < n; ++i) { dmma(e0, e1, x, y); dmma(f0, f1, x, y); dmma(g0, g1, x, y); dmma(h0, h1, x, y); }  // 4 independent
    long long t5 = clock64();
    int idx = threadIdx.x;
    for (int i = 0; i < n; ++i) idx = (int)sm[idx & 1023] & 1023;       // dependent LDS (+ cvt)
    long long t6 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    long long t7 = clock64();
    double s = x;
    for (int i = 0; i < n; ++i) s = sqrt(s + 2.0);
    long long t8 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + d + c0 + c1 + e0 + e1 + f0 + f1 + g0 + g1 + h0 + h1 + idx + s;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; cyc[6] = t7 - t6; cyc[7] = t8 - t7;
    }
}
int main() {
    double *out; long long *cyc, h[8];
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 64);
    const int n = 2000;
    for (int threads : {32, 256, 1024}) {
        k<<<148, threads>>>(out, cyc, n); cudaDeviceSynchronize();
        k<<<148, threads>>>(out, cyc, n); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads/CTA=%4d  cycles per iteration: DFMA dep %.1f | 8 indep DFMA %.1f (%.2f/op) | DDIV dep %.1f | DMMA dep %.1f | 4 indep DMMA %.1f (%.2f/op) | LDS dep %.1f | bar.sync %.1f | DSQRT dep %.1f\n",
               threads, h[0] / (double)n, h[1] / (double)n, h[1] / (8.0 * n), h[2] / (double)n, h[3] / (double)n, h[4] / (double)n, h[4] / (4.0 * n),
               h[5] / (double)n, h[6] / (double)n, h[7] / (double)n);
    }
    return 0;
}
