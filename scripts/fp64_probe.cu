// fp64 throughput of ONE SM: plain DFMA vs the tensor-core DMMA.8x8x4 (mma.sync m8n8k4 f64), 16 warps, 8 independent chains each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/fp64_probe.cu -o scripts/fp64_probe.bin
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_k(double* out, int iters, long long* cyc) {
    double a[8], x = 1.0000001 + threadIdx.x * 1e-9, y = 1e-9;
    for (int i = 0; i < 8; ++i) a[i] = i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void dmma_k(double* out, int iters, long long* cyc) {
    double c[8][2], a0 = 1.0 + threadIdx.x * 1e-9, b0 = 1e-3;
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a0), "d"(b0));
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int warps : {1, 4, 16}) {
        long long h;
        dfma_k<<<1, 32 * warps>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        dfma_k<<<1, 32 * warps>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA  %2d warps: %lld cycles, %.2f thread-FMA/clk/SM, %.2f cycles per warp instruction per warp\n", warps, h,
               (double)iters * 8 * 32 * warps / h, (double)h / (iters * 8));
        dmma_k<<<1, 32 * warps>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        dmma_k<<<1, 32 * warps>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DMMA  %2d warps: %lld cycles, %.2f FMA/clk/SM, %.2f cycles per DMMA.8x8x4 per warp\n", warps, h,
               (double)iters * 8 * 256 * warps / h, (double)h / (iters * 8));
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
