// Generic strided SIMT GEMM tile:  Cout[m,n] = alpha * sum_k A[m,k] * B[n,k] + beta * Add[m,n]
//
// Used for the small fp64 linear algebra of the shared factor (Gram, blocked Cholesky updates,
// triangular solves) and as the fp32 reference implementation of the apply GEMMs that the
// tcgen05 kernel is validated against.  Arbitrary element strides on A and B make every
// transpose free; 64x64x16 tiles, 256 threads, 4x4 register blocking.
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>

namespace uce {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

template <typename TA, typename TB, typename TAcc, typename TC>
struct SimtGemmArgs {
    int M, N, Kd;
    const TA* A; long sa_m, sa_k;
    const TB* B; long sb_n, sb_k;
    TC* C; long ldc;
    const TC* Add; long ldadd;   // may alias C (in-place accumulate); nullptr => beta ignored
    TAcc alpha, beta;
    int lower_only;              // skip tiles strictly above the block diagonal (SYRK)
};

template <typename TA, typename TB, typename TAcc, typename TC>
__device__ __forceinline__ void simt_gemm_tile(const SimtGemmArgs<TA, TB, TAcc, TC>& g, int tile_m, int tile_n,
                                               TAcc (*As)[SG_BM + 4], TAcc (*Bs)[SG_BN + 4]) {
    const int tid = threadIdx.x;
    const int m0 = tile_m * SG_BM, n0 = tile_n * SG_BN;
    const int tx = tid % 16, ty = tid / 16;   // 16 x 16 threads, each 4 x 4 outputs
    TAcc acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = TAcc(0);

    const bool a_kfast = (g.sa_k == 1);
    const bool b_kfast = (g.sb_k == 1);

    for (int k0 = 0; k0 < g.Kd; k0 += SG_BK) {
        // ---- stage A[BM x BK] and B[BN x BK] (k-major in smem) ----
#pragma unroll
        for (int it = 0; it < (SG_BM * SG_BK) / SG_THREADS; ++it) {
            int idx = tid + it * SG_THREADS, mm, kk;
            if (a_kfast) { kk = idx % SG_BK; mm = idx / SG_BK; } else { mm = idx % SG_BM; kk = idx / SG_BM; }
            int gm = m0 + mm, gk = k0 + kk;
            TAcc v = TAcc(0);
            if (gm < g.M && gk < g.Kd) v = (TAcc)g.A[(long)gm * g.sa_m + (long)gk * g.sa_k];
            As[kk][mm] = v;
        }
#pragma unroll
        for (int it = 0; it < (SG_BN * SG_BK) / SG_THREADS; ++it) {
            int idx = tid + it * SG_THREADS, nn, kk;
            if (b_kfast) { kk = idx % SG_BK; nn = idx / SG_BK; } else { nn = idx % SG_BN; kk = idx / SG_BN; }
            int gn = n0 + nn, gk = k0 + kk;
            TAcc v = TAcc(0);
            if (gn < g.N && gk < g.Kd) v = (TAcc)g.B[(long)gn * g.sb_n + (long)gk * g.sb_k];
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SG_BK; ++kk) {
            TAcc a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gm = m0 + ty * 4 + i;
        if (gm >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
            if (gn >= g.N) continue;
            TAcc v = g.alpha * acc[i][j];
            if (g.Add) v += g.beta * (TAcc)g.Add[(long)gm * g.ldadd + gn];
            g.C[(long)gm * g.ldc + gn] = (TC)v;
        }
    }
}

template <typename TA, typename TB, typename TAcc, typename TC>
__global__ void __launch_bounds__(SG_THREADS) simt_gemm_kernel(SimtGemmArgs<TA, TB, TAcc, TC> g) {
    __shared__ TAcc As[SG_BK][SG_BM + 4];
    __shared__ TAcc Bs[SG_BK][SG_BN + 4];
    // programmatic dependent launch (the blocked Cholesky of the general factor is ~100 of these back to back): wait for the kernel in
    // front, then let the one behind start its launch; both are no-ops for a plain launch
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (g.lower_only && blockIdx.x > blockIdx.y) return;
    simt_gemm_tile<TA, TB, TAcc, TC>(g, blockIdx.y, blockIdx.x, As, Bs);
}

template <typename TA, typename TB, typename TAcc, typename TC>
inline cudaError_t simt_gemm(cudaStream_t st, int M, int N, int Kd, const TA* A, long sa_m, long sa_k, const TB* B,
                             long sb_n, long sb_k, TC* C, long ldc, TAcc alpha = TAcc(1), const TC* Add = nullptr,
                             long ldadd = 0, TAcc beta = TAcc(0), int lower_only = 0) {
    if (M <= 0 || N <= 0) return cudaSuccess;
    SimtGemmArgs<TA, TB, TAcc, TC> g{M, N, Kd, A, sa_m, sa_k, B, sb_n, sb_k, C, ldc, Add, ldadd, alpha, beta, lower_only};
    dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(SG_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool pdl = getenv("UCE_NO_PDL") == nullptr;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, simt_gemm_kernel<TA, TB, TAcc, TC>, g);
}

}  // namespace uce
