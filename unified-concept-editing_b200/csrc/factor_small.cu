// Low-latency path of the shared factor for the common case n <= 160 concept rows (dual system, n <= K):
// 4 kernels instead of the ~35 launches of the general blocked path in factor.cu.
//
//   gram_splitk   H = Cp Cp^T, fp64 accumulation of exact fp32 products, 32x32 tiles x split-K over many CTAs
//   chol_small    ONE CTA: H (+ lamb/s on the diagonal) into shared memory, blocked Cholesky (NB = 32: warp-shuffle
//                 factorisation of the diagonal block, warp-parallel inverse, panel + trailing updates by all 512
//                 threads), then  Z = H^-1[:, edit]  by blocked forward/backward substitution
//   q_emit        Q = Z^T Cp (fp64 accumulate) -> Q, Qt and the TF32 hi/lo splits consumed by the tcgen05 apply
//
// Same algebra and same fp64 precision as the general path (trainscripts/uce_sd_erase.py:63,71,79,82 — the
// mat2 accumulation and its inverse — done once per edit).
#include "uce_ws.h"

namespace uce {

constexpr int FS_NB = 32;
constexpr int FS_MAX_N = 160;
constexpr int FS_KSPLIT = 64;

__device__ __forceinline__ float fs_tf32_hi(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// ---------------------------------------------------------------------------------------------------------
// H[i,j] += sum_{k in split} Cp[i,k] Cp[j,k]   for tile pairs ti >= tj (lower block triangle); H pre-zeroed.
__global__ void __launch_bounds__(256) gram_splitk_kernel(const float* __restrict__ Cp, int n, int K, double* __restrict__ H, int ld) {
    __shared__ double A[FS_NB][FS_KSPLIT + 1];
    __shared__ double B[FS_NB][FS_KSPLIT + 1];
    // decode the lower-triangular tile pair
    int p = blockIdx.x, ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= p) ++ti;
    const int tj = p - ti * (ti + 1) / 2;
    const int k0 = blockIdx.y * FS_KSPLIT;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < FS_NB * FS_KSPLIT; idx += 256) {
        const int r = idx / FS_KSPLIT, k = idx % FS_KSPLIT;
        const int gi = ti * FS_NB + r, gj = tj * FS_NB + r, gk = k0 + k;
        A[r][k] = (gi < n && gk < K) ? (double)Cp[(long)gi * K + gk] : 0.0;
        B[r][k] = (gj < n && gk < K) ? (double)Cp[(long)gj * K + gk] : 0.0;
    }
    __syncthreads();
    const int tx = tid % 16, ty = tid / 16;   // 2 x 2 outputs per thread
    double acc[2][2] = {{0, 0}, {0, 0}};
#pragma unroll 8
    for (int k = 0; k < FS_KSPLIT; ++k) {
        const double a0 = A[ty * 2][k], a1 = A[ty * 2 + 1][k];
        const double b0 = B[tx * 2][k], b1 = B[tx * 2 + 1][k];
        acc[0][0] = fma(a0, b0, acc[0][0]); acc[0][1] = fma(a0, b1, acc[0][1]);
        acc[1][0] = fma(a1, b0, acc[1][0]); acc[1][1] = fma(a1, b1, acc[1][1]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int gi = ti * FS_NB + ty * 2 + i, gj = tj * FS_NB + tx * 2 + j;
            if (gi < n && gj < n && gj <= gi) atomicAdd(&H[(long)gi * ld + gj], acc[i][j]);
        }
}

// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Warp-level Cholesky of a 32 x 32 block: lane i holds row i in registers; template recursion keeps every
// register index a compile-time constant (a runtime-indexed array would live in local memory).
template <int J, int Kc>
__device__ __forceinline__ void chol_row_update(double (&a)[FS_NB], int lane) {
    if constexpr (Kc < FS_NB) {
        const double akj = __shfl_sync(0xffffffffu, a[J], Kc);     // L[Kc][J]
        if (lane >= Kc) a[Kc] -= a[J] * akj;
        chol_row_update<J, Kc + 1>(a, lane);
    }
}
template <int J>
__device__ __forceinline__ void chol_column(double (&a)[FS_NB], int lane, bool& bad) {
    if constexpr (J < FS_NB) {
        double d = __shfl_sync(0xffffffffu, a[J], J);
        if (!(d > 0.0)) { bad = true; d = 1.0; }
        const double s = sqrt(d), inv = 1.0 / s;
        if (lane == J) a[J] = s; else if (lane > J) a[J] *= inv;
        chol_row_update<J, J + 1>(a, lane);
        chol_column<J + 1>(a, lane, bad);
    }
}
template <int J>
__device__ __forceinline__ void block_load(double (&a)[FS_NB], const double* row) {
    if constexpr (J < FS_NB) { a[J] = row[J]; block_load<J + 1>(a, row); }
}
template <int J>
__device__ __forceinline__ void block_store(const double (&a)[FS_NB], double* row, int lane) {
    if constexpr (J < FS_NB) { row[J] = (J <= lane) ? a[J] : 0.0; block_store<J + 1>(a, row, lane); }
}

// Kept out of line so that the 64 registers of the row block do not compete with the caller's live values.
__device__ __noinline__ void chol_diag_block(double* row, int lane, int* flag, int kb) {
    double a[FS_NB];
    block_load<0>(a, row);
    bool bad = false;
    chol_column<0>(a, lane, bad);
    if (bad && lane == 0) atomicCAS(flag, 0, 1 + kb);
    block_store<0>(a, row, lane);
}

// Single-CTA Cholesky + solve.  S: dynamic smem [n_pad][n_pad + 1] doubles.
constexpr int FS_T = 512;   // threads of the single factor CTA (128 registers each: the diagonal block lives in registers)

__global__ void __launch_bounds__(FS_T, 1)
chol_small_kernel(double* __restrict__ Hg, int n, int n_pad, const double* __restrict__ dadd, int n_pres, int n_edit,
                  double* __restrict__ Linv_g, double* __restrict__ X, int ldx, int write_back, int* flag) {
    extern __shared__ double S[];
    __shared__ double Dinv[FS_NB][FS_NB + 1];
    const int lds = n_pad + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = n_pad / FS_NB;

    for (int idx = tid; idx < n_pad * n_pad; idx += FS_T) {
        const int r = idx / n_pad, c = idx % n_pad;
        double v = (r < n && c <= r) ? Hg[(long)r * n_pad + c] : 0.0;
        if (r == c) v = (r < n) ? v + dadd[r] : 1.0;
        S[r * lds + c] = v;
    }
    // rhs: unit vectors of the edit rows (internal order: preserve rows first)
    for (int idx = tid; idx < n_pad * n_edit; idx += FS_T) {
        const int r = idx / n_edit, j = idx % n_edit;
        __stcg(&X[(long)r * ldx + j], (r == n_pres + j) ? 1.0 : 0.0);
    }
    __syncthreads();

    for (int kb = 0; kb < nblk; ++kb) {
        const int o = kb * FS_NB;
        // (a) diagonal block, one warp, rows in registers
        if (warp == 0) {
            chol_diag_block(&S[(o + lane) * lds + o], lane, flag, kb);
        }
        __syncthreads();
        // (b) inverse of the lower-triangular diagonal block: warp c -> column c
        for (int c = warp; c < FS_NB; c += FS_T / 32) {
            double x = 0.0;
            if (lane == c) x = 1.0 / S[(o + c) * lds + o + c];
            for (int i = c + 1; i < FS_NB; ++i) {
                const double part = (lane >= c && lane < i) ? S[(o + i) * lds + o + lane] * x : 0.0;
                const double s = warp_sum(part);
                if (lane == i) x = -s / S[(o + i) * lds + o + i];
            }
            Dinv[lane][c] = x;     // x == 0 above the diagonal
            Linv_g[((long)kb * FS_NB + lane) * FS_NB + c] = x;
        }
        __syncthreads();
        const int rest = n_pad - (o + FS_NB);
        if (rest > 0) {
            // (c) panel  L_ik = H_ik Linv^T   (results staged in registers: the panel is updated in place)
            double out[FS_MAX_N * FS_NB / FS_T];
#pragma unroll
            for (int it = 0; it < FS_MAX_N * FS_NB / FS_T; ++it) {
                const int idx = tid + it * FS_T;
                out[it] = 0.0;
                if (idx < rest * FS_NB) {
                    const int r = o + FS_NB + idx / FS_NB, c = idx % FS_NB;
                    double s = 0.0;
#pragma unroll 8
                    for (int j = 0; j < FS_NB; ++j) s = fma(S[r * lds + o + j], Dinv[c][j], s);
                    out[it] = s;
                }
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < FS_MAX_N * FS_NB / FS_T; ++it) {
                const int idx = tid + it * FS_T;
                if (idx < rest * FS_NB) S[(o + FS_NB + idx / FS_NB) * lds + o + idx % FS_NB] = out[it];
            }
            __syncthreads();
            // (d) trailing update of the lower triangle
            for (int idx = tid; idx < rest * rest; idx += FS_T) {
                const int r = o + FS_NB + idx / rest, c = o + FS_NB + idx % rest;
                if (c > r) continue;
                double s = 0.0;
#pragma unroll 8
                for (int j = 0; j < FS_NB; ++j) s = fma(S[r * lds + o + j], S[c * lds + o + j], s);
                S[r * lds + c] -= s;
            }
            __syncthreads();
        }
    }
    if (write_back)
        for (int idx = tid; idx < n_pad * n_pad; idx += FS_T) Hg[idx] = S[(idx / n_pad) * lds + idx % n_pad];

    // ---- forward substitution  L Y = rhs  (blocks above the first edit row stay zero) ----
    constexpr int XR = (FS_NB * FS_MAX_N + FS_T - 1) / FS_T;     // staged outputs per thread for a 32 x n_edit block
    for (int kb = n_pres / FS_NB; kb < nblk; ++kb) {
        const int o = kb * FS_NB;
        for (int c = warp; c < FS_NB; c += FS_T / 32) Dinv[lane][c] = Linv_g[((long)kb * FS_NB + lane) * FS_NB + c];
        __syncthreads();
        double out[XR];
#pragma unroll
        for (int it = 0; it < XR; ++it) {
            const int idx = tid + it * FS_T;
            out[it] = 0.0;
            if (idx < FS_NB * n_edit) {
                const int rr = idx / n_edit, j = idx % n_edit;
                double s = 0.0;
                for (int c = 0; c <= rr; ++c) s = fma(Dinv[rr][c], __ldcg(&X[(long)(o + c) * ldx + j]), s);
                out[it] = s;
            }
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < XR; ++it) {
            const int idx = tid + it * FS_T;
            if (idx < FS_NB * n_edit) __stcg(&X[(long)(o + idx / n_edit) * ldx + idx % n_edit], out[it]);
        }
        __syncthreads();
        const int rest = n_pad - (o + FS_NB);
        for (int idx = tid; idx < rest * n_edit; idx += FS_T) {
            const int r = o + FS_NB + idx / n_edit, j = idx % n_edit;
            double s = 0.0;
#pragma unroll 8
            for (int c = 0; c < FS_NB; ++c) s = fma(S[r * lds + o + c], __ldcg(&X[(long)(o + c) * ldx + j]), s);
            __stcg(&X[(long)r * ldx + j], __ldcg(&X[(long)r * ldx + j]) - s);
        }
        __syncthreads();
    }
    // ---- backward substitution  L^T Z = Y ----
    for (int kb = nblk - 1; kb >= 0; --kb) {
        const int o = kb * FS_NB;
        for (int c = warp; c < FS_NB; c += FS_T / 32) Dinv[lane][c] = Linv_g[((long)kb * FS_NB + lane) * FS_NB + c];
        __syncthreads();
        double out[XR];
#pragma unroll
        for (int it = 0; it < XR; ++it) {
            const int idx = tid + it * FS_T;
            out[it] = 0.0;
            if (idx < FS_NB * n_edit) {
                const int rr = idx / n_edit, j = idx % n_edit;
                double s = 0.0;
                for (int c = rr; c < FS_NB; ++c) s = fma(Dinv[c][rr], __ldcg(&X[(long)(o + c) * ldx + j]), s);
                out[it] = s;
            }
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < XR; ++it) {
            const int idx = tid + it * FS_T;
            if (idx < FS_NB * n_edit) __stcg(&X[(long)(o + idx / n_edit) * ldx + idx % n_edit], out[it]);
        }
        __syncthreads();
        for (int idx = tid; idx < o * n_edit; idx += FS_T) {
            const int r = idx / n_edit, j = idx % n_edit;
            double s = 0.0;
#pragma unroll 8
            for (int c = 0; c < FS_NB; ++c) s = fma(S[(o + c) * lds + r], __ldcg(&X[(long)(o + c) * ldx + j]), s);
            __stcg(&X[(long)r * ldx + j], __ldcg(&X[(long)r * ldx + j]) - s);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Q[j, k] = sum_r Z[r, j] Cp[r, k]  (fp64 accumulate), emitted as Q, Qt and the tf32 hi/lo splits of Qt.
// grid: K / 32 column tiles; 256 threads.
__global__ void __launch_bounds__(256) q_emit_kernel(const double* __restrict__ Z, int ldx, const float* __restrict__ Cp, int n,
                                                     int n_edit, int r_pad, int K, float* __restrict__ Q, float* __restrict__ Qt,
                                                     float* __restrict__ Qt_hi, float* __restrict__ Qt_lo) {
    __shared__ float Cs[FS_MAX_N][33];
    const int k0 = blockIdx.x * 32, tid = threadIdx.x;
    for (int idx = tid; idx < n * 32; idx += 256) {
        const int r = idx / 32, c = idx % 32;
        Cs[r][c] = (k0 + c < K) ? Cp[(long)r * K + k0 + c] : 0.f;
    }
    __syncthreads();
    const int c = tid % 32;
    for (int j = tid / 32; j < r_pad; j += 8) {
        double s = 0.0;
        if (j < n_edit)
            for (int r = 0; r < n; ++r) s = fma(Z[(long)r * ldx + j], (double)Cs[r][c], s);
        const float v = (float)s;
        if (k0 + c < K) {
            Q[(long)j * K + k0 + c] = v;
            const long t = (long)(k0 + c) * r_pad + j;
            const float h = fs_tf32_hi(v);
            Qt[t] = v; Qt_hi[t] = h; Qt_lo[t] = v - h;
        }
    }
}

// E = G_e - C_e with its tf32 split, and the packed concept rows; one block per row.
__global__ void pack_rows_split_kernel(const float* __restrict__ C, const float* __restrict__ G, const int* __restrict__ src, int n_act,
                                       int n_pres, int rank_pad, int K, float* __restrict__ Cp, float* __restrict__ E,
                                       float* __restrict__ E_hi, float* __restrict__ E_lo) {
    const int r = blockIdx.x;
    const int n_edit = n_act - n_pres;
    if (r < n_act) {
        const int s = src[r];
        const bool is_edit = r >= n_pres;
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const float v = C[(long)s * K + k];
            Cp[(long)r * K + k] = v;
            if (is_edit) {
                const float e = G[(long)s * K + k] - v, h = fs_tf32_hi(e);
                const long t = (long)(r - n_pres) * K + k;
                E[t] = e; E_hi[t] = h; E_lo[t] = e - h;
            }
        }
    } else {
        const int j = n_edit + (r - n_act);
        if (j < rank_pad)
            for (int k = threadIdx.x; k < K; k += blockDim.x) {
                const long t = (long)j * K + k;
                E[t] = 0.f; E_hi[t] = 0.f; E_lo[t] = 0.f;
            }
    }
}

bool factor_small_applicable(const uce_ws* ws, int n, bool dual) {
    return dual && n <= FS_MAX_N && !ws->force_general;
}

// Preconditions: ws->h_src_idx / h_diag_add staged and copied, flag cleared, n_edit > 0.
int factor_small(uce_ws* ws, const float* C, const float* G, int n, int n_pres, int n_edit, cudaStream_t st, int* launches) {
    const int K = ws->K;
    const int n_pad = round_up(n, FS_NB);
    const int ldx = ws->max_rows;
    ws->sys_n = n_pad;
    pack_rows_split_kernel<<<n + (ws->rank_pad - n_edit), 256, 0, st>>>(C, G, ws->src_idx, n, n_pres, ws->rank_pad, K, ws->Cp, ws->E,
                                                                         ws->E_hi, ws->E_lo);
    UCE_LAUNCH_CHECK(); ++*launches;
    UCE_CUDA(cudaMemsetAsync(ws->H, 0, (size_t)n_pad * n_pad * sizeof(double), st));
    const int nt = n_pad / FS_NB;
    gram_splitk_kernel<<<dim3(nt * (nt + 1) / 2, ceil_div(K, FS_KSPLIT)), 256, 0, st>>>(ws->Cp, n, K, ws->H, n_pad);
    UCE_LAUNCH_CHECK(); ++*launches;
    if (ws->debug) {
        if (!ws->Hcopy) UCE_CUDA(cudaMalloc(&ws->Hcopy, (size_t)ws->sys_max * ws->sys_max * sizeof(double)));
        // debug copy holds the assembled system: gram + diagonal (the kernel adds the diagonal in smem only)
        UCE_CUDA(cudaMemcpyAsync(ws->Hcopy, ws->H, (size_t)n_pad * n_pad * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    const size_t smem = (size_t)n_pad * (n_pad + 1) * sizeof(double);
    static size_t configured = 0;
    if (configured < smem) {
        UCE_CUDA(cudaFuncSetAttribute(chol_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    chol_small_kernel<<<1, FS_T, smem, st>>>(ws->H, n, n_pad, ws->diag_add, n_pres, n_edit, ws->Linv, ws->X, ldx, ws->debug, ws->flag);
    UCE_LAUNCH_CHECK(); ++*launches;
    q_emit_kernel<<<ceil_div(K, 32), 256, 0, st>>>(ws->X, ldx, ws->Cp, n, n_edit, ws->rank_pad, K, ws->Q, ws->Qt, ws->Qt_hi, ws->Qt_lo);
    UCE_LAUNCH_CHECK(); ++*launches;
    return 0;
}

}  // namespace uce
