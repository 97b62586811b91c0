// Phase 1 of the UCE edit: the projection-independent factor.
//
// Reference semantics (trainscripts/uce_sd_erase.py:58-82): every projection l gets
//     W_new = (lamb W + sum_i s_i (W g_i) c_i^T) (lamb I + sum_i s_i c_i c_i^T)^-1 .
// Since (W g_i) is linear in W, W_new = W + (W E^T) Q with
//     E = G_e - C_e                              [n_edit, K]
//     Q = S_e C_e (lamb I + C^T S C)^-1          [n_edit, K]
// and Q is the same for all projections.  Q is obtained in fp64 from the smaller of two
// equivalent SPD systems:
//   dual   (n <= K):  H = lamb S^-1 + C C^T  [n,n];   Q = (rows `edit` of H^-1) C
//   primal (n >  K):  B = lamb I + C^T S C   [K,K];   Q^T = B^-1 (C_e^T S_e)
// (Woodbury; the dual form never places lamb next to the O(1e5) Gram entries in fp32 —
// SURVEY.md §7 H1.)  Internally rows are ordered preserve-first / edit-last.
//
// General-size path (systems larger than factor_small.cu takes): Gram on the fp64 tensor pipe (gram_dmma_kernel), blocked
// right-looking Cholesky (NB = 32) in global memory — diagonal blocks factored and inverted by one CTA, panels and trailing updates as
// strided fp64 SIMT GEMMs launched programmatically behind each other — and for the dual system the substitution + Q emission as one
// fp64 tensor-pipe kernel (solve_emit_dmma_kernel); the primal system keeps GEMM-based triangular solves.
#include "uce_ws.h"
#include "gemm_simt.cuh"
#include "tc_common.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdint>

namespace uce {

// ---------------------------------------------------------------------------------------------
// pack: Cp[r,:] = C[src[r],:]   (all active rows, internal order)
//       E[j,:]  = G[e_j,:] - C[e_j,:] for the active edit rows (internal rows n_pres + j); pad rows 0
//       Cs64[r,:] = s_r * Cp[r,:]  (primal only)
struct alignas(16) FactorTables { unsigned int w[4096]; };     // n row indices (int), padded to 8 bytes, then n diagonal terms (double)
__global__ void __launch_bounds__(256) factor_tables_kernel(const __grid_constant__ FactorTables t, int n, int* src_idx, double* diag_add, int* flag) {
    pdl_wait(); pdl_launch();
    const double* dsrc = reinterpret_cast<const double*>(t.w + n + (n & 1));
    for (int i = threadIdx.x; i < n; i += blockDim.x) { src_idx[i] = (int)t.w[i]; diag_add[i] = dsrc[i]; }
    if (threadIdx.x == 0) *flag = 0;
}

__global__ void pack_rows_kernel(const float* __restrict__ C, const float* __restrict__ G,
                                 const int* __restrict__ src, const double* __restrict__ dadd, int n_act,
                                 int n_pres, int rank_pad, int K, float* __restrict__ Cp, float* __restrict__ E,
                                 double* __restrict__ Cs64) {
    int r = blockIdx.x;                 // 0 .. n_act + (rank_pad - n_edit) - 1
    int n_edit = n_act - n_pres;
    if (r < n_act) {
        int s = src[r];
        const float* c = C + (long)s * K;
        float* cp = Cp + (long)r * K;
        const bool is_edit = r >= n_pres;
        const float* g = is_edit ? G + (long)s * K : nullptr;   // edit rows come first in the API: s < n_edit_api
        float* e = is_edit ? E + (long)(r - n_pres) * K : nullptr;
        double sc = Cs64 ? dadd[r] : 0.0;
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            float v = c[k];
            cp[k] = v;
            if (is_edit) e[k] = g[k] - v;
            if (Cs64) Cs64[(long)r * K + k] = sc * (double)v;
        }
    } else {
        int j = n_edit + (r - n_act);   // padded E row
        if (j < rank_pad)
            for (int k = threadIdx.x; k < K; k += blockDim.x) E[(long)j * K + k] = 0.f;
    }
}

// H[r,r] += dadd[r] (r < n) ; H[r,r] = 1 for the padding rows; (dual: dadd = lamb/s_r, primal: lamb)
__global__ void finish_diag_kernel(double* H, int ld, int n, int n_pad, const double* dadd, double lamb_primal) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    if (r < n) H[(long)r * ld + r] += dadd ? dadd[r] : lamb_primal;
    else       H[(long)r * ld + r] = 1.0;
}

// dual rhs: X[r,j] = (r == n_pres + j);  primal rhs: X[m,j] = Cs64[n_pres + j, m]
__global__ void fill_rhs_kernel(double* X, int ldx, int n_pad, int n_rhs, int n_pres, const double* Cs64, int K) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int r = blockIdx.y;
    if (j >= n_rhs || r >= n_pad) return;
    double v;
    if (Cs64) v = (r < K) ? Cs64[(long)(n_pres + j) * K + r] : 0.0;
    else      v = (r == n_pres + j) ? 1.0 : 0.0;
    X[(long)r * ldx + j] = v;
}

// Q[j,k] (f32, ld K) and Qt[k,j] (f32, ld rank_pad) from the fp64 solution.
//   dual  : handled by GEMM (Q = X^T Cp), this kernel only transposes Q -> Qt
//   primal: X = Q^T [K_pad, n_rhs] -> both
__global__ void emit_q_kernel(const double* X, int ldx, int from_x, int rank, int rank_pad, int K, float* Q, float* Qt) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y;
    if (k >= K || j >= rank_pad) return;
    float v = 0.f;
    if (j < rank) {
        if (from_x) { v = (float)X[(long)k * ldx + j]; Q[(long)j * K + k] = v; }
        else        v = Q[(long)j * K + k];
    } else {
        Q[(long)j * K + k] = 0.f;
    }
    Qt[(long)k * rank_pad + j] = v;
}

// hi = rna_tf32(x), lo = x - hi for the pre-split B operands of the tcgen05 apply kernels (E and Qt)
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float v = x[i];
        uint32_t u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
        const float h = __uint_as_float(u);
        hi[i] = h; lo[i] = v - h;
    }
}
static int split_operand(uce_ws* ws, const float* x, float* hi, float* lo, cudaStream_t st, int* launches) {
    const long n = (long)ws->rank_pad * ws->K;
    split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, hi, lo, n);
    UCE_LAUNCH_CHECK();
    *launches += 1;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Dual system, general size:  X = H^-1 Cp[:, 8 columns]  per CTA by blocked forward substitution over all rows and backward
// substitution down to the block row of the first edit row (they are the LAST rows); the edit rows of X are Q[:, those columns]
// (H^-1 is symmetric: Q = J H^-1 Cp) -> Q, Qt and the tf32 splits.  Same algorithm as solve_emit_kernel of the low-latency path, but L
// (lower triangle of H after the factorisation, row-major, ld) and the dense inverses of its diagonal blocks (Linv) stay in global
// memory / L2 and stream through a double-buffered 32 x 32 shared-memory block: the next block is fetched into registers while the
// current one is used.  Replaces 128 launches of block triangular solves on n_edit unit vectors plus a Q = X^T Cp GEMM (1.9 of the
// 3.6 ms general factor at BASELINE cfg4) by one launch of K / 8 CTAs.
constexpr int SG_CW = 8, SG_PF = 4;
__global__ void __launch_bounds__(256) solve_emit_general_kernel(const double* __restrict__ L, int ld, const double* __restrict__ Linv,
                                                                 const float* __restrict__ Cp, int n, int n_pad, int n_pres, int n_edit, int r_pad, int K,
                                                                 float* __restrict__ Q, float* __restrict__ Qt, float* __restrict__ Qt_hi, float* __restrict__ Qt_lo) {
    extern __shared__ double gsm_d[];
    constexpr int NBK = UCE_NB, P = NBK + 1, XL = SG_CW + 1;
    double* XS = gsm_d;                                   // [n_pad][XL]
    double* BLK = XS + (size_t)n_pad * XL;                // [2][32][33]
    const int tid = threadIdx.x, c = tid & (SG_CW - 1), r = tid >> 3;      // slab column; row of a 32-row block
    const int lr = tid >> 3, lc = (tid & 7) * 4;          // block fetch: row, first of four columns
    const int k0 = blockIdx.x * SG_CW, nblk = n_pad / NBK, kb_e = n_pres / NBK;
    for (int idx = tid; idx < n_pad * SG_CW; idx += 256) {
        const int rr = idx / SG_CW, cc = idx % SG_CW;
        XS[rr * XL + cc] = (rr < n && k0 + cc < K) ? (double)Cp[(long)rr * K + k0 + cc] : 0.0;
    }
    // block sequence: phase 0 forward: (kb: inverse of the diagonal block, then L(i, kb) for i > kb); phase 1 backward:
    // (kb from the last block down to kb_e: inverse of the diagonal block, then L(kb, i) for kb_e <= i < kb)
    struct It { int phase, kb, i; };
    auto valid = [&](const It& t) { return t.phase < 2; };
    auto advance = [&](It t) {
        if (t.phase == 0) {
            if (t.i < 0) t.i = t.kb + 1; else ++t.i;
            if (t.i >= nblk) { ++t.kb; t.i = -1; if (t.kb >= nblk) { t.phase = 1; t.kb = nblk - 1; } }
        } else {
            if (t.i < 0) t.i = kb_e; else ++t.i;
            if (t.i >= t.kb) { --t.kb; t.i = -1; if (t.kb < kb_e) t.phase = 2; }
        }
        return t;
    };
    auto src_of = [&](const It& t, int& sld) -> const double* {
        if (t.i < 0) { sld = NBK; return Linv + (size_t)t.kb * NBK * NBK; }
        sld = ld;
        return t.phase == 0 ? L + (size_t)t.i * NBK * ld + (size_t)t.kb * NBK : L + (size_t)t.kb * NBK * ld + (size_t)t.i * NBK;
    };
    auto fetch = [&](const It& t, double (&v)[4]) {
        int sld; const double* s = src_of(t, sld);
        const double2 a = *reinterpret_cast<const double2*>(s + (size_t)lr * sld + lc), b = *reinterpret_cast<const double2*>(s + (size_t)lr * sld + lc + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    };
    // The blocks of L are read-only here, so they are requested SG_PF block steps ahead (registers: four doubles per thread and step):
    // one step of work (~300 cycles) does not cover an L2 round trip (the first version, one block ahead, spent ~2 200 cycles per block).
    It cur{0, 0, -1}, ahead{0, 0, -1};
    double pf[SG_PF][4];
#pragma unroll
    for (int d = 0; d < SG_PF; ++d)
        if (valid(ahead)) { fetch(ahead, pf[d]); ahead = advance(ahead); }
    int buf = 0;
    while (valid(cur)) {
#pragma unroll
      for (int d = 0; d < SG_PF; ++d) {
        if (!valid(cur)) break;
        double* B = BLK + buf * NBK * P;
#pragma unroll
        for (int q = 0; q < 4; ++q) B[lr * P + lc + q] = pf[d][q];
        __syncthreads();                                   // block visible; everybody is done with the previous block and its writes to XS
        const It nx = advance(cur);
        if (valid(ahead)) { fetch(ahead, pf[d]); ahead = advance(ahead); }      // refill this slot: SG_PF steps ahead
        const int o = cur.kb * NBK;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        if (cur.i < 0) {
            // triangular multiply with the dense inverse: forward Y_k = Linv X_k (Linv[r][j]); backward X_k = Linv^T Y_k (Linv[j][r])
            if (cur.phase == 0) {
#pragma unroll
                for (int j = 0; j < NBK; j += 4) {
                    a0 = fma(B[r * P + j], XS[(o + j) * XL + c], a0);         a1 = fma(B[r * P + j + 1], XS[(o + j + 1) * XL + c], a1);
                    a2 = fma(B[r * P + j + 2], XS[(o + j + 2) * XL + c], a2); a3 = fma(B[r * P + j + 3], XS[(o + j + 3) * XL + c], a3);
                }
            } else {
#pragma unroll
                for (int j = 0; j < NBK; j += 4) {
                    a0 = fma(B[j * P + r], XS[(o + j) * XL + c], a0);           a1 = fma(B[(j + 1) * P + r], XS[(o + j + 1) * XL + c], a1);
                    a2 = fma(B[(j + 2) * P + r], XS[(o + j + 2) * XL + c], a2); a3 = fma(B[(j + 3) * P + r], XS[(o + j + 3) * XL + c], a3);
                }
            }
            __syncthreads();                               // every thread has read the old X_k
            XS[(o + r) * XL + c] = (a0 + a1) + (a2 + a3);
        } else {
            // update of another block row: forward X_i -= L(i,kb) Y_k (B[r][j]); backward Y_i -= L(kb,i)^T X_k (B[j][r])
            if (cur.phase == 0) {
#pragma unroll
                for (int j = 0; j < NBK; j += 4) {
                    a0 = fma(B[r * P + j], XS[(o + j) * XL + c], a0);         a1 = fma(B[r * P + j + 1], XS[(o + j + 1) * XL + c], a1);
                    a2 = fma(B[r * P + j + 2], XS[(o + j + 2) * XL + c], a2); a3 = fma(B[r * P + j + 3], XS[(o + j + 3) * XL + c], a3);
                }
            } else {
#pragma unroll
                for (int j = 0; j < NBK; j += 4) {
                    a0 = fma(B[j * P + r], XS[(o + j) * XL + c], a0);           a1 = fma(B[(j + 1) * P + r], XS[(o + j + 1) * XL + c], a1);
                    a2 = fma(B[(j + 2) * P + r], XS[(o + j + 2) * XL + c], a2); a3 = fma(B[(j + 3) * P + r], XS[(o + j + 3) * XL + c], a3);
                }
            }
            XS[(cur.i * NBK + r) * XL + c] -= (a0 + a1) + (a2 + a3);
        }
        cur = nx;
        buf ^= 1;
      }
    }
    __syncthreads();
    for (int idx = tid; idx < r_pad * SG_CW; idx += 256) {
        const int cc = idx / r_pad, j = idx % r_pad;
        if (k0 + cc >= K) continue;
        const float v = (j < n_edit) ? (float)XS[(n_pres + j) * XL + cc] : 0.f;
        const long t = (long)(k0 + cc) * r_pad + j;
        uint32_t u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
        const float h = __uint_as_float(u);
        Qt[t] = v; Qt_hi[t] = h; Qt_lo[t] = v - h;
        Q[(long)j * K + k0 + cc] = v;
    }
}

// The same solve on the fp64 TENSOR pipe (mma.sync m8n8k4.f64), 16 columns of Cp per CTA.  ncu on the SIMT version above
// (profiles/r02_solve_general_ncu.txt): two shared-memory loads per fma — 71 % of the shared-memory pipe's cycles, the fp64 pipe 12 %
// active, 1.25 ms at BASELINE cfg4.  Here a warp owns an 8-row tile of the block row it works on: the A fragment (8 x 4 of a block of L)
// comes STRAIGHT from global memory / L2 in fragment layout (eight 32-byte sectors per warp instruction, no staging, no barrier), the
// B fragments (Y_k, 32 x 16) are loaded from the slab once per block step and kept in registers for every block of that column, and
// the 8 x 8 result tiles are subtracted from the slab in place.  Warps 0-3 and 4-7 take alternate blocks of the column; the L tiles
// are requested SD_PF blocks ahead.
constexpr int SD_CW = 16, SD_XL = 20, SD_PF = 4;
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256, 1) solve_emit_dmma_kernel(const double* __restrict__ L, int ld, const double* __restrict__ Linv,
                                                                 const float* __restrict__ Cp, int n, int n_pad, int n_pres, int n_edit, int r_pad, int K,
                                                                 float* __restrict__ Q, float* __restrict__ Qt, float* __restrict__ Qt_hi, float* __restrict__ Qt_lo) {
    extern __shared__ double XS[];                         // [n_pad][SD_XL]
    constexpr int NBK = UCE_NB, XL = SD_XL;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int m = warp & 3, grp = warp >> 2;               // 8-row tile of a block; which of the two blocks in flight (diagonal step: column tile)
    const int k0 = blockIdx.x * SD_CW, nblk = n_pad / NBK, kb_e = n_pres / NBK;
    for (int idx = tid; idx < n_pad * SD_CW; idx += 256) {
        const int rr = idx / SD_CW, cc = idx % SD_CW;
        XS[rr * XL + cc] = (rr < n && k0 + cc < K) ? (double)Cp[(long)rr * K + k0 + cc] : 0.0;
    }
    __syncthreads();
    // fragments: A[g][4 ks + t], B[4 ks + t][g], C[g][2 t + {0, 1}]
    for (int phase = 0; phase < 2; ++phase) {
        const int kb_first = phase == 0 ? 0 : nblk - 1, kb_last = phase == 0 ? nblk - 1 : kb_e, step = phase == 0 ? 1 : -1;
        for (int kb = kb_first; phase == 0 ? kb <= kb_last : kb >= kb_last; kb += step) {
            const int o = kb * NBK;
            {   // diagonal block: X_k <- Linv_kk X_k (forward) / Linv_kk^T X_k (backward); this warp: rows m, columns of tile grp
                const double* Li = Linv + (size_t)kb * NBK * NBK;
                double a[8], b[8];
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    a[ks] = phase == 0 ? Li[(m * 8 + g) * NBK + ks * 4 + t] : Li[(ks * 4 + t) * NBK + m * 8 + g];
                    b[ks] = XS[(o + ks * 4 + t) * XL + grp * 8 + g];
                }
                double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;       // two chains
#pragma unroll
                for (int ks = 0; ks < 8; ks += 2) { dmma884(c0, c1, a[ks], b[ks]); dmma884(e0, e1, a[ks + 1], b[ks + 1]); }
                __syncthreads();                                    // every warp has read the old X_k
                XS[(o + m * 8 + g) * XL + grp * 8 + 2 * t] = c0 + e0;
                XS[(o + m * 8 + g) * XL + grp * 8 + 2 * t + 1] = c1 + e1;
                __syncthreads();
            }
            // the other blocks of this column: forward X_i -= L(i, kb) Y_k for i > kb; backward Y_i -= L(kb, i)^T X_k for kb_e <= i < kb
            const int i_lo = phase == 0 ? kb + 1 : kb_e, i_hi = phase == 0 ? nblk : kb;       // [i_lo, i_hi)
            if (i_lo < i_hi) {
                double yb[2][8];
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) { yb[0][ks] = XS[(o + ks * 4 + t) * XL + g]; yb[1][ks] = XS[(o + ks * 4 + t) * XL + 8 + g]; }
                auto load_a = [&](int i, double (&a)[8]) {
                    const double* src = phase == 0 ? L + (size_t)(i * NBK + m * 8 + g) * ld + (size_t)o + t
                                                   : L + (size_t)(o + t) * ld + (size_t)i * NBK + m * 8 + g;
                    const size_t ks_stride = phase == 0 ? 4 : (size_t)4 * ld;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) a[ks] = src[ks * ks_stride];
                };
                double pa[SD_PF][8];
                int i_f = i_lo + grp;
#pragma unroll
                for (int d = 0; d < SD_PF; ++d)
                    if (i_f < i_hi) { load_a(i_f, pa[d]); i_f += 2; }
                for (int i = i_lo + grp; i < i_hi;) {
#pragma unroll
                    for (int d = 0; d < SD_PF; ++d) {
                        if (i >= i_hi) break;
                        double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) { dmma884(c00, c01, pa[d][ks], yb[0][ks]); dmma884(c10, c11, pa[d][ks], yb[1][ks]); }
                        if (i_f < i_hi) { load_a(i_f, pa[d]); i_f += 2; }
                        double* x = XS + (size_t)(i * NBK + m * 8 + g) * XL + 2 * t;
                        x[0] -= c00; x[1] -= c01; x[8] -= c10; x[9] -= c11;
                        i += 2;
                    }
                }
                __syncthreads();                                    // the column is applied: the next diagonal step reads the slab
            }
        }
    }
    for (int idx = tid; idx < r_pad * SD_CW; idx += 256) {
        const int cc = idx / r_pad, j = idx % r_pad;
        if (k0 + cc >= K) continue;
        const float v = (j < n_edit) ? (float)XS[(n_pres + j) * XL + cc] : 0.f;
        const long tq = (long)(k0 + cc) * r_pad + j;
        uint32_t u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
        const float h = __uint_as_float(u);
        Qt[tq] = v; Qt_hi[tq] = h; Qt_lo[tq] = v - h;
        Q[(long)j * K + k0 + cc] = v;
    }
}

// H = Cp Cp^T for the general path on the fp64 tensor pipe: one CTA per 64 x 64 tile of the lower tile triangle, the whole K range in one
// CTA (no split, no atomics: 136 CTAs at n = 1000, one wave).  Concept rows are fp32 — their products are exact in fp64 —, staged in
// shared memory as floats (pitch 36: the fragment loads of a warp hit 32 different banks) and converted when a fragment is built.  A warp
// owns 8 rows of the tile and sweeps its eight 8-column tiles: one A fragment + eight B fragments per 8 DMMA.  Replaces the generic fp64
// SIMT GEMM (0.45 ms of the 1.9 ms factor at BASELINE cfg4).
constexpr int GD_T = 64, GD_KC = 32, GD_P = 36;
__global__ void __launch_bounds__(256, 1) gram_dmma_kernel(const float* __restrict__ Cp, int n, int K, double* __restrict__ H, int ld) {
    __shared__ float As[2][GD_T * GD_P];
    __shared__ float Bs[2][GD_T * GD_P];
    int p = blockIdx.x, ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= p) ++ti;
    const int tj = p - ti * (ti + 1) / 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    // staging: thread = (row = tid / 4 [+ 0], 8 consecutive floats at column (tid % 4) * 8) of a 64 x 32 chunk
    const int lr = tid >> 2, lc = (tid & 3) * 8;
    const bool diag = ti == tj;
    auto fetch = [&](int k0, float4 (&a)[2], float4 (&b)[2]) {
        const int ra = ti * GD_T + lr, rb = tj * GD_T + lr;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = k0 + lc + 4 * h;
            a[h] = (ra < n && k < K) ? *reinterpret_cast<const float4*>(Cp + (long)ra * K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
            b[h] = (!diag && rb < n && k < K) ? *reinterpret_cast<const float4*>(Cp + (long)rb * K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto stash = [&](int buf, const float4 (&a)[2], const float4 (&b)[2]) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            *reinterpret_cast<float4*>(&As[buf][lr * GD_P + lc + 4 * h]) = a[h];
            *reinterpret_cast<float4*>(&Bs[buf][lr * GD_P + lc + 4 * h]) = diag ? a[h] : b[h];
        }
    };
    double acc[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { acc[nt][0] = 0.0; acc[nt][1] = 0.0; }
    float4 pa[2], pb[2];
    fetch(0, pa, pb);
    stash(0, pa, pb);
    __syncthreads();
    const int n_chunks = (K + GD_KC - 1) / GD_KC;
    for (int ch = 0; ch < n_chunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < n_chunks) fetch((ch + 1) * GD_KC, pa, pb);             // in flight during the 64 DMMA below
        const float* Aw = &As[buf][(warp * 8 + g) * GD_P + t];
        const float* Bw = &Bs[buf][g * GD_P + t];
#pragma unroll
        for (int ks = 0; ks < GD_KC / 4; ++ks) {
            const double a = (double)Aw[ks * 4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) dmma884(acc[nt][0], acc[nt][1], a, (double)Bw[nt * 8 * GD_P + ks * 4]);
        }
        if (ch + 1 < n_chunks) stash(buf ^ 1, pa, pb);                     // the other buffer: everybody left it at the previous barrier
        __syncthreads();
    }
    const int row = ti * GD_T + warp * 8 + g;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        double* o = H + (long)row * ld + tj * GD_T + nt * 8 + 2 * t;
        o[0] = acc[nt][0]; o[1] = acc[nt][1];
    }
}

#define UCE_RT(expr)                                                                             \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int)_e;                                                                      \
        }                                                                                        \
        ++launches;                                                                              \
    } while (0)

// Blocked Cholesky of the sys_n x sys_n matrix in ws->H (lower), then solve H X = rhs in place.
// fwd_from: first block row whose rhs is non-zero (dual: rhs = unit vectors of the edit rows).
// factor_only: stop after the factorisation (the dual path solves with solve_emit_general_kernel instead).
static int cholesky_solve(uce_ws* ws, int n_pad, int n_rhs, int ldx, int fwd_from, cudaStream_t st, int& launches, bool factor_only = false) {
    const int nb = UCE_NB, nblk = n_pad / nb, ld = n_pad;
    double* H = ws->H; double* X = ws->X; double* Linv = ws->Linv;
    for (int k = 0; k < nblk; ++k) {
        UCE_RT((cudaError_t)potrf_inv_general(H, ld, k, Linv, ws->flag, st));      // diagonal block: factor in place + dense inverse (factor_small.cu)
        int rest = n_pad - (k + 1) * nb;
        if (rest > 0) {
            double* panel = H + (long)(k + 1) * nb * ld + (long)k * nb;            // [rest, nb], ld
            // L_ik = H_ik * Linv_kk^T   (in place)
            UCE_RT((simt_gemm<double, double, double, double>(st, rest, nb, nb, panel, ld, 1, Linv + (long)k * nb * nb, nb, 1,
                                                              panel, ld)));
            // H_ij -= L_ik L_jk^T   (lower tiles only)
            double* trail = H + (long)(k + 1) * nb * ld + (long)(k + 1) * nb;
            UCE_RT((simt_gemm<double, double, double, double>(st, rest, rest, nb, panel, ld, 1, panel, ld, 1, trail, ld, -1.0,
                                                              trail, ld, 1.0, /*lower_only=*/1)));
        }
    }
    if (factor_only) return 0;
    // forward: L Y = rhs
    for (int k = fwd_from; k < nblk; ++k) {
        double* xk = X + (long)k * nb * ldx;
        UCE_RT((simt_gemm<double, double, double, double>(st, nb, n_rhs, nb, Linv + (long)k * nb * nb, nb, 1, xk, 1, ldx, xk, ldx)));
        int rest = n_pad - (k + 1) * nb;
        if (rest > 0) {
            double* panel = H + (long)(k + 1) * nb * ld + (long)k * nb;
            double* xr = X + (long)(k + 1) * nb * ldx;
            UCE_RT((simt_gemm<double, double, double, double>(st, rest, n_rhs, nb, panel, ld, 1, xk, 1, ldx, xr, ldx, -1.0, xr, ldx,
                                                              1.0)));
        }
    }
    // backward: L^T X = Y
    for (int k = nblk - 1; k >= 0; --k) {
        double* xk = X + (long)k * nb * ldx;
        UCE_RT((simt_gemm<double, double, double, double>(st, nb, n_rhs, nb, Linv + (long)k * nb * nb, 1, nb, xk, 1, ldx, xk, ldx)));
        if (k > 0) {
            double* lrow = H + (long)k * nb * ld;          // L[k-block rows, 0 : k*nb]
            UCE_RT((simt_gemm<double, double, double, double>(st, k * nb, n_rhs, nb, lrow, 1, ld, xk, 1, ldx, X, ldx, -1.0, X, ldx,
                                                              1.0)));
        }
    }
    return 0;
}

int factor_dev(uce_ws* ws, const float* C, const float* G, const float* scales, int n_rows, int n_edit_api, float lamb,
               cudaStream_t st) {
    const int K = ws->K;
    int launches = 0;
    ws->mode = 0;
    // The pinned staging below is read by async copies when they execute: wait until the copies of the
    // previous call have been consumed before overwriting it (skipped while capturing a graph — a captured
    // graph re-reads the staging on replay, so it stays valid only while the workspace keeps this problem).
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    UCE_CUDA(cudaStreamIsCapturing(st, &cap));
    if (cap == cudaStreamCaptureStatusNone && ws->stage_pending) { UCE_CUDA(cudaEventSynchronize(ws->ev_stage)); ws->stage_pending = 0; }
    // ---- host: active rows, internal order = preserve first, edit last ----
    int n_pres = 0, n_edit = 0;
    for (int i = n_edit_api; i < n_rows; ++i)
        if (scales[i] != 0.f) ws->h_src_idx[n_pres++] = i;
    for (int i = 0; i < n_edit_api; ++i)
        if (scales[i] != 0.f) ws->h_src_idx[n_pres + n_edit++] = i;
    const int n = n_pres + n_edit;
    ws->n_act = n; ws->n_edit = n_edit; ws->n_pres = n_pres; ws->lamb = lamb;
    ws->rank = n_edit;
    ws->rank_pad = std::max(UCE_RANK_PAD, round_up(n_edit, UCE_RANK_PAD));
    ws->dense = (n_edit > K / 2) ? 1 : 0;
    ws->launches_factor = 0;
    if (n_edit == 0) { ws->mode = (n <= K) ? 1 : 2; ws->sys_n = 0; return 0; }   // nothing to edit: W_new = W_old

    // the dual system needs S^-1 > 0; with a negative scale fall back to the primal K x K system
    // (still Cholesky: fails with UCE_E_NOT_SPD at uce_ws_check if lamb I + C^T S C is indefinite)
    bool dual = (n <= K);
    for (int r = 0; r < n; ++r) dual = dual && (scales[ws->h_src_idx[r]] > 0.f);
    for (int r = 0; r < n; ++r) {
        double s = (double)scales[ws->h_src_idx[r]];
        ws->h_diag_add[r] = dual ? (double)lamb / s : s;
    }
    if (!dual && !ws->Cs64) UCE_CUDA(cudaMalloc(&ws->Cs64, (size_t)ws->max_rows * K * sizeof(double)));
    if (ws->dense && !ws->Dt) UCE_CUDA(cudaMalloc(&ws->Dt, (size_t)K * K * sizeof(float)));
    // row order, diagonal terms and the cleared status flag go to the device as kernel parameters: a copy-engine transfer on this stream
    // would queue behind the weight uploads of the host-buffer call (uce_api.cu) and hold the whole factor back until they end
    if ((size_t)n * 12 + 4 <= sizeof(FactorTables)) {
        FactorTables t;
        memcpy(t.w, ws->h_src_idx, (size_t)n * 4);
        memcpy(t.w + n + (n & 1), ws->h_diag_add, (size_t)n * 8);
        UCE_CUDA(launch_k(factor_tables_kernel, dim3(1), dim3(256), 0, st, 1, t, n, ws->src_idx, ws->diag_add, ws->flag));
        ++launches;
    } else {
        int rc = table_upload(ws->src_idx, ws->h_src_idx, (size_t)n * sizeof(int), st, &launches);
        if (!rc) rc = table_upload(ws->diag_add, ws->h_diag_add, (size_t)n * sizeof(double), st, &launches);
        if (rc) return rc;
        UCE_CUDA(cudaMemsetAsync(ws->flag, 0, sizeof(int), st));
    }
    if (cap == cudaStreamCaptureStatusNone) { UCE_CUDA(cudaEventRecord(ws->ev_stage, st)); ws->stage_pending = 1; }

    if (factor_small_applicable(ws, n, n_edit, dual)) {
        int rc = factor_small(ws, C, G, n, n_pres, n_edit, st, &launches);
        if (rc) return rc;
        if (ws->debug) {   // the debug copy shows the assembled system (the kernel adds the diagonal in shared memory only)
            finish_diag_kernel<<<ceil_div(ws->sys_n, 256), 256, 0, st>>>(ws->Hcopy, ws->sys_n, n, ws->sys_n, ws->diag_add, 0.0);
            UCE_RT(cudaGetLastError());
        }
        if (ws->dense) {
            UCE_RT((simt_gemm<float, float, double, float>(st, K, K, n_edit, ws->Q, 1, K, ws->E, 1, K, ws->Dt, K)));
        }
        ws->mode = 1;
        ws->launches_factor = launches;
        return 0;
    }
    pack_rows_kernel<<<n + (ws->rank_pad - n_edit), 256, 0, st>>>(C, G, ws->src_idx, ws->diag_add, n, n_pres, ws->rank_pad, K,
                                                                  ws->Cp, ws->E, dual ? nullptr : ws->Cs64);
    UCE_RT(cudaGetLastError());
    if (!ws->dense) {      // tf32 hi/lo split of E right away: the apply's first kernel needs nothing else (uce_edit_dev_f32 forks here)
        int rc2 = split_operand(ws, ws->E, ws->E_hi, ws->E_lo, st, &launches);
        if (rc2) return rc2;
        if (ws->want_ev_E) { UCE_CUDA(cudaEventRecord(ws->ev_E, st)); ws->ev_E_recorded = 1; }
    }

    const int n_sys = dual ? n : K;
    const int n_pad = round_up(n_sys, UCE_NB);
    const int ldx = ws->max_rows;   // row stride of X
    ws->sys_n = n_pad;
    UCE_CUDA(cudaMemsetAsync(ws->H, 0, (size_t)n_pad * n_pad * sizeof(double), st));
    if (dual) {
        // H = Cp Cp^T (fp64 accumulate of exact fp32 products), lower tiles
        if (n_pad % GD_T == 0 && K % 4 == 0 && getenv("UCE_GENERAL_GRAM_SIMT") == nullptr) {
            const int tiles = n_pad / GD_T;
            gram_dmma_kernel<<<tiles * (tiles + 1) / 2, 256, 0, st>>>(ws->Cp, n, K, ws->H, n_pad);
            UCE_RT(cudaGetLastError());
        } else {
            UCE_RT((simt_gemm<float, float, double, double>(st, n, n, K, ws->Cp, K, 1, ws->Cp, K, 1, ws->H, n_pad, 1.0, nullptr, 0, 0.0, 1)));
        }
        finish_diag_kernel<<<ceil_div(n_pad, 256), 256, 0, st>>>(ws->H, n_pad, n, n_pad, ws->diag_add, 0.0);
        UCE_RT(cudaGetLastError());
    } else {
        // B = Cp^T S Cp : A[m,r] = Cs64[r,m], B[nn,r] = Cp[r,nn]
        UCE_RT((simt_gemm<double, float, double, double>(st, K, K, n, ws->Cs64, 1, K, ws->Cp, 1, K, ws->H, n_pad, 1.0, nullptr, 0, 0.0, 1)));
        finish_diag_kernel<<<ceil_div(n_pad, 256), 256, 0, st>>>(ws->H, n_pad, K, n_pad, nullptr, (double)lamb);
        UCE_RT(cudaGetLastError());
    }
    if (ws->debug) {
        if (!ws->Hcopy) UCE_CUDA(cudaMalloc(&ws->Hcopy, (size_t)ws->sys_max * ws->sys_max * sizeof(double)));
        UCE_CUDA(cudaMemcpyAsync(ws->Hcopy, ws->H, (size_t)n_pad * n_pad * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    const size_t smem_sg = ((size_t)n_pad * (SG_CW + 1) + 2 * UCE_NB * (UCE_NB + 1)) * sizeof(double);
    const size_t smem_sd = (size_t)n_pad * SD_XL * sizeof(double);
    const bool fused_solve = dual && smem_sg <= 200 * 1024 && getenv("UCE_GENERAL_SOLVE_GEMMS") == nullptr;
    const bool dmma_solve = fused_solve && smem_sd <= 200 * 1024 && getenv("UCE_GENERAL_SOLVE_SIMT") == nullptr;
    bool qt_split_done = false;
    if (dmma_solve) {
        int rc = cholesky_solve(ws, n_pad, n_edit, ldx, 0, st, launches, /*factor_only=*/true);
        if (rc) return rc;
        static thread_local size_t conf_sd[64] = {0};
        if (conf_sd[ws->device & 63] < smem_sd) {
            UCE_CUDA(cudaFuncSetAttribute(solve_emit_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sd));
            conf_sd[ws->device & 63] = smem_sd;
        }
        solve_emit_dmma_kernel<<<ceil_div(K, SD_CW), 256, smem_sd, st>>>(ws->H, n_pad, ws->Linv, ws->Cp, n, n_pad, n_pres, n_edit, ws->rank_pad, K,
                                                                         ws->Q, ws->Qt, ws->Qt_hi, ws->Qt_lo);
        UCE_RT(cudaGetLastError());
        qt_split_done = true;
    } else if (fused_solve) {
        int rc = cholesky_solve(ws, n_pad, n_edit, ldx, 0, st, launches, /*factor_only=*/true);
        if (rc) return rc;
        static thread_local size_t conf_sg[64] = {0};
        if (conf_sg[ws->device & 63] < smem_sg) {
            UCE_CUDA(cudaFuncSetAttribute(solve_emit_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sg));
            conf_sg[ws->device & 63] = smem_sg;
        }
        solve_emit_general_kernel<<<ceil_div(K, SG_CW), 256, smem_sg, st>>>(ws->H, n_pad, ws->Linv, ws->Cp, n, n_pad, n_pres, n_edit, ws->rank_pad, K,
                                                                            ws->Q, ws->Qt, ws->Qt_hi, ws->Qt_lo);
        UCE_RT(cudaGetLastError());
        qt_split_done = true;
    } else {
        fill_rhs_kernel<<<dim3(ceil_div(n_edit, 128), n_pad), 128, 0, st>>>(ws->X, ldx, n_pad, n_edit, n_pres, dual ? nullptr : ws->Cs64, K);
        UCE_RT(cudaGetLastError());
        int rc = cholesky_solve(ws, n_pad, n_edit, ldx, dual ? n_pres / UCE_NB : 0, st, launches);
        if (rc) return rc;
        if (dual) {
            // Q[j,k] = sum_r X[r,j] Cp[r,k]
            UCE_RT((simt_gemm<double, float, double, float>(st, n_edit, K, n, ws->X, 1, ldx, ws->Cp, 1, K, ws->Q, K)));
        }
        emit_q_kernel<<<dim3(ceil_div(K, 128), ws->rank_pad), 128, 0, st>>>(ws->X, ldx, dual ? 0 : 1, n_edit, ws->rank_pad, K, ws->Q, ws->Qt);
        UCE_RT(cudaGetLastError());
    }
    if (ws->dense) {
        // Dt[a,b] = D[b,a] = sum_j E[j,b] Q[j,a]
        UCE_RT((simt_gemm<float, float, double, float>(st, K, K, n_edit, ws->Q, 1, K, ws->E, 1, K, ws->Dt, K)));
    }
    ws->mode = dual ? 1 : 2;
    if (!ws->dense && ws->rank > 0 && !qt_split_done) {      // tf32 hi/lo split of Qt: always, so the apply implementation may be chosen after the factor
        int rc2 = split_operand(ws, ws->Qt, ws->Qt_hi, ws->Qt_lo, st, &launches);
        if (rc2) return rc2;
    }
    // the general path used H: clear the part the single-CTA factor accumulates into (it skips its own memset when H_dirty == 0)
    if (!ws->debug) {
        const size_t clean = std::min((size_t)ws->sys_max * ws->sys_max, (size_t)160 * 160) * sizeof(double);
        UCE_CUDA(cudaMemsetAsync(ws->H, 0, clean, st));
        ws->H_dirty = 0;
    } else {
        ws->H_dirty = 1;
    }
    ws->launches_factor = launches;
    return 0;
}

}  // namespace uce
