// Low-latency path of the shared factor for the common case n <= 160 concept rows (dual system, n <= K):
// five kernels that hand over by programmatic dependent launch, instead of the ~100 launches of the general blocked path in factor.cu.
//
//   factor_tables (factor.cu) row order, diagonal terms, status flag — as kernel parameters
//   gram_pack     H = Cp Cp^T, fp64 accumulation of exact fp32 products, 32x32 tiles x split-K over many CTAs; the same launch packs
//                 the concept rows (Cp) and E = G_e - C_e with its tf32 split in extra blocks (the Gram blocks read C directly)
//   chol_small    ONE CTA: H (+ lamb/s on the diagonal) into shared memory (lower block triangle), blocked Cholesky
//                 (NB = 32: factorisation of the diagonal block by one warp, panel by per-row forward substitution, trailing
//                 update on the fp64 tensor pipe with the next diagonal block first — lookahead); L goes back to global memory
//                 and H is left cleared for the next factor
//   inv_blocks    one CTA per diagonal block: L_kk^-1 (column sweeps), in place in the global factor
//   solve_emit    MANY CTAs, one per 8 columns of K:  X = H^-1 Cp[:, cols]  by blocked forward substitution and the
//                 part of the backward substitution that reaches the edit rows (they are the LAST rows), on the fp64 tensor
//                 pipe; the edit rows of X are Q[:, cols] (H^-1 is symmetric:  Q = J H^-1 Cp) -> Q, Qt and the
//                 TF32 hi/lo splits consumed by the tcgen05 apply.
//   (+ potrf_inv_general: the diagonal-block step of the general path, built from the same pieces)
//
// Same algebra and same fp64 precision as the general path (trainscripts/uce_sd_erase.py:63,71,79,82 — the
// mat2 accumulation and its inverse — done once per edit).
#include "uce_ws.h"
#include "tc_common.cuh"
#include <cstdlib>
#include <cstring>

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
// The concept rows are read straight from the caller's C through the row order `src` (internal row r = C[src[r]]), so this kernel
// does not wait for the pack kernel: the pack work (pack_rows_split below: Cp, E and its tf32 split) rides in the SAME launch as
// extra blocks (behind the Gram blocks of the 1-D grid).
__device__ void pack_rows_split_block(int r, const float* __restrict__ C, const float* __restrict__ G, const int* __restrict__ src, int n_act,
                                      int n_pres, int rank_pad, int K, float* __restrict__ Cp, float* __restrict__ E,
                                      float* __restrict__ E_hi, float* __restrict__ E_lo);
struct FsSrc { int v[FS_MAX_N]; };                // the row order as a KERNEL PARAMETER: no table to wait for, no index round trip before the rows
__global__ void __launch_bounds__(256) gram_pack_kernel(const float* __restrict__ C, const __grid_constant__ FsSrc srcp, int n, int K, double* __restrict__ H, int ld,
                                                        int n_pairs, const float* __restrict__ G, int n_pres, int rank_pad, int n_pack_rows,
                                                        float* __restrict__ Cp, float* __restrict__ E, float* __restrict__ E_hi, float* __restrict__ E_lo) {
    __shared__ double A[FS_NB][FS_KSPLIT + 1];
    __shared__ double B[FS_NB][FS_KSPLIT + 1];
    // Programmatic launch: this kernel starts while the table kernel in front of it still runs.  It READS only the caller's C / G
    // (the row order arrives as a parameter), so the loads go out at once; the wait comes before the first WRITE (the previous edit's
    // kernels may still read Cp / E / H until the table kernel, which waited for them, has finished).
    pdl_launch();
    const int* src = srcp.v;
    // 1-D grid: the Gram blocks first — (tile pair, K split) = blockIdx.x % n_pairs, / n_pairs — then one block per packed row.  (The
    // first version was a 2-D grid whose pack blocks existed once per K split and returned at once for all but one: 2 148 CTAs, 2.4 waves.)
    const int n_gram = n_pairs * ((K + FS_KSPLIT - 1) / FS_KSPLIT);
    if ((int)blockIdx.x >= n_gram) {
        const int r = (int)blockIdx.x - n_gram;
        pdl_wait();
        if (r < n_pack_rows) pack_rows_split_block(r, C, G, src, n, n_pres, rank_pad, K, Cp, E, E_hi, E_lo);
        return;
    }
    // decode the lower-triangular tile pair
    int p = (int)blockIdx.x % n_pairs, ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= p) ++ti;
    const int tj = p - ti * (ti + 1) / 2;
    const int k0 = ((int)blockIdx.x / n_pairs) * FS_KSPLIT;
    const int tid = threadIdx.x;
    {
        // thread = (row r0 + 4 it, column k): ALL row indices first, then ALL values, then the stores — two memory round trips per CTA.
        // (As one loop of "index, value, store" the compiler kept the eight iterations in order: 16 dependent round trips, 28 k cycles
        // per CTA of which 1.3 k are arithmetic.)
        static_assert(FS_KSPLIT == 64 && FS_NB == 32, "8 rows per thread");
        const int k = tid & 63, r0 = tid >> 6, gk = k0 + k;
        const bool diag = ti == tj;
        int sa[8], sb[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int gi = ti * FS_NB + r0 + 4 * it, gj = tj * FS_NB + r0 + 4 * it;
            sa[it] = (gi < n) ? src[gi] : -1;
            sb[it] = (!diag && gj < n) ? src[gj] : -1;
        }
        float va[8], vb[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            va[it] = (sa[it] >= 0 && gk < K) ? C[(long)sa[it] * K + gk] : 0.f;
            vb[it] = (sb[it] >= 0 && gk < K) ? C[(long)sb[it] * K + gk] : 0.f;
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            A[r0 + 4 * it][k] = (double)va[it];
            B[r0 + 4 * it][k] = (double)(diag ? va[it] : vb[it]);
        }
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
    pdl_wait();
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int gi = ti * FS_NB + ty * 2 + i, gj = tj * FS_NB + tx * 2 + j;
            if (gi < n && gj < n && gj <= gi) atomicAdd(&H[(long)gi * ld + gj], acc[i][j]);
        }
}

// ---------------------------------------------------------------------------------------------------------
// Single-CTA Cholesky + solve, everything in shared memory:
//   SB   lower block triangle of H, 32 x 32 blocks of pitch 33, block (bi,bj) at index bi(bi+1)/2 + bj
//   XS   right-hand sides / solution  [n_pad][xl]   (n_edit <= 64 columns)
// After step kb the diagonal block holds L_kk in its lower triangle (zeros above) and invd[] holds 1 / L_ii; the inverses of the
// diagonal blocks, which only solve_emit needs, are computed by inv_blocks_kernel on the global copy of the factor.
constexpr int FS_T = 512;
constexpr int FS_BLK = FS_NB * (FS_NB + 1);      // doubles per block (pitch 33)
constexpr int FS_MAX_RHS = 64;

__device__ __forceinline__ int fs_blk(int bi, int bj) { return (bi * (bi + 1) / 2 + bj) * FS_BLK; }

// 1 / sqrt(d) in fp64 without the slow software sqrt/div
__device__ __forceinline__ double fs_rsqrt(double d) {
    // fp32 seed (relative error 2^-22) + ONE third-order (Halley) step: e = 1 - d y0^2, y = y0 + y0 e (1/2 + 3/8 e), error
    // ~ e^3 = 2^-63.  Four dependent fp64 operations instead of the six of two Newton steps: the reciprocal square root
    // sits on the critical path of each of the 160 sequential pivots.
    const double y0 = (double)rsqrtf((float)d);
    const double e = fma(-(d * y0), y0, 1.0);
    return fma(y0 * e, fma(0.375, e, 0.5), y0);
}

// fp64 tensor-pipe multiply-add of one warp: D[8 x 8] += A[8 x 4] B[4 x 8];  lane = 4 g + t holds A[g][t], B[t][g], D[g][2 t + {0, 1}]
__device__ __forceinline__ void fs_dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// (Measured and rejected: taking the pivot chain through a reciprocal — d_next = A[j+1][j+1] - A[j+1][j]^2 / d, with the reciprocal
// square root beside it — shortens the dependent chain by two fp64 operations on paper and ran 4 k cycles per block SLOWER: the
// extra fp64 instructions cost more issue time on this part than the shorter chain saves.)
// One pivot step with the trailing update limited to KM columns (straight-line: the column loads are issued ahead of the
// FMAs; a per-column early exit was measured 1.8x SLOWER — the branches serialise load and FMA latencies).
// `d` is the pivot A[j][j] of this step; the return value is the pivot of step j + 1.  The pivot chain does NOT go through shared
// memory: lane j + 1 updates its own diagonal entry itself and broadcasts it by shuffle while the column of L travels through shared
// memory for everybody else (11.6 k -> 11.0 k cycles per block).  Also measured: the multipliers by shuffle instead of shared memory,
// which removes both warp barriers and lets the next pivot's rsqrt interleave with the update — 64 shuffles per pivot cost more
// than that buys (17.1 k cycles per block).
template <int KM>
__device__ __forceinline__ double fs_potrf_step(double (&a)[FS_NB], double* __restrict__ D, double* __restrict__ invd_blk, int lane, int j, double d, bool& bad) {
    if (!(d > 0.0)) { bad = true; d = 1.0; }
    const double y = fs_rsqrt(d);
    const double l = (lane == j) ? d * y : a[0] * y;      // L[lane][j] for lanes >= j
    const double d_next = __shfl_sync(0xffffffffu, fma(-l, l, a[1]), (j + 1) & 31);
    if (lane >= j) D[lane * (FS_NB + 1) + j] = l;
    if (lane == j) invd_blk[j] = y;
    __syncwarp();
    // L[j + k][j] was just written to shared memory by lane j + k: ONE broadcast 64-bit load per column instead of the two
    // 32-bit shuffles a double costs (rows beyond 31 read neighbouring shared memory: finite garbage for columns that do not exist)
    const double* Lj = D + j * (FS_NB + 1) + j;
#pragma unroll
    for (int k = 1; k <= KM; ++k) {                       // A[lane][j + k] -= L[lane][j] L[j + k][j]
        const double lk = Lj[k * (FS_NB + 1)];
        a[k - 1] = fma(-l, lk, a[k]);
    }
    __syncwarp();
    return d_next;
}
__device__ __noinline__ void fs_potrf_warp(double* __restrict__ D, double* __restrict__ invd_blk, int lane, int* flag, int kb) {
    double a[FS_NB];
#pragma unroll
    for (int c = 0; c < FS_NB; ++c) a[c] = D[lane * (FS_NB + 1) + c];
    bool bad = false;
    double d = __shfl_sync(0xffffffffu, a[0], 0);
    // columns j + k <= 31 exist: 31, 23, 15 and 7 trailing columns for the four quarters of the block
#pragma unroll 1
    for (int j = 0; j < 8; ++j) d = fs_potrf_step<31>(a, D, invd_blk, lane, j, d, bad);
#pragma unroll 1
    for (int j = 8; j < 16; ++j) d = fs_potrf_step<23>(a, D, invd_blk, lane, j, d, bad);
#pragma unroll 1
    for (int j = 16; j < 24; ++j) d = fs_potrf_step<15>(a, D, invd_blk, lane, j, d, bad);
#pragma unroll 1
    for (int j = 24; j < 32; ++j) d = fs_potrf_step<7>(a, D, invd_blk, lane, j, d, bad);
    if (bad && lane == 0) atomicCAS(flag, 0, 1 + kb);
}

// Panel row by forward substitution:  x L_kk^T = h  for ONE row h of the panel (thread = row; the row lives in registers and
// shifts left one column per step, like the diagonal block in fs_potrf_step, so every index is static).  The coefficients
// L_kk[j + k][j] are warp-wide broadcasts from shared memory.  Per step: one multiply (x_j = h_j / L_jj) and KM independent fmas.
// This replaces "explicit inverse of L_kk (4.5 k cycles, all warps, a 32-step serial chain) + dense multiply (2.3-7.2 k cycles)" of
// every block step by one 2.5 k-cycle phase; the inverses, which only the solve kernel needs, moved to their own small kernel.
// LS is the diagonal block with every row scaled by the reciprocal of its diagonal entry, LS[c][j] = L_kk[c][j] / L_kk[c][c]: with the
// right-hand side scaled the same way (a_c = h_c / L_cc) a step is  x_j = a_0  and  a_c -= x_j LS[c][j] — ONE dependent fp64
// operation per step on the chain of 32 instead of two (multiply by 1 / L_jj, then the fma): 5.5 k -> ~3 k cycles per panel.
template <int KM>
__device__ __forceinline__ void fs_trsm_step(double (&a)[FS_NB], const double* __restrict__ LS, double* __restrict__ row, int j) {
    const double x = a[0];
    row[j] = x;
    const double* Lj = LS + j * (FS_NB + 1) + j;
#pragma unroll
    for (int k = 1; k <= KM; ++k) a[k - 1] = fma(-x, Lj[k * (FS_NB + 1)], a[k]);
}
__device__ __noinline__ void fs_trsm_row(double* __restrict__ row, const double* __restrict__ LS, const double* __restrict__ invd_blk) {
    double a[FS_NB];
#pragma unroll
    for (int c = 0; c < FS_NB; ++c) a[c] = row[c] * invd_blk[c];
#pragma unroll 1
    for (int j = 0; j < 8; ++j) fs_trsm_step<31>(a, LS, row, j);
#pragma unroll 1
    for (int j = 8; j < 16; ++j) fs_trsm_step<23>(a, LS, row, j);
#pragma unroll 1
    for (int j = 16; j < 24; ++j) fs_trsm_step<15>(a, LS, row, j);
#pragma unroll 1
    for (int j = 24; j < 32; ++j) fs_trsm_step<7>(a, LS, row, j);
}

__global__ void __launch_bounds__(FS_T, 1)
chol_small_kernel(double* __restrict__ Hg, int n, int n_pad, const double* __restrict__ dadd,
                  double* __restrict__ Lg, double* __restrict__ invd_g, int write_back, int* flag, long long* __restrict__ trace, int trail_warps) {
    extern __shared__ double smem_d[];
    pdl_wait(); pdl_launch();                      // inv_blocks is launched now and parks at its own wait
    int trn = 0;
    auto tr = [&]() { if (trace && threadIdx.x == 0 && trn < 64) trace[trn++] = clock64(); };
    tr();
    const int nblk = n_pad / FS_NB;
    double* SB = smem_d;
    double* invd = SB + (nblk * (nblk + 1) / 2) * FS_BLK;      // [n_pad]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int P = FS_NB + 1;                  // block pitch

    // ---- load the lower block triangle (+ diagonal term, identity padding): warp = block row, lane = column; the global
    //      loads of eight rows are requested before the first one is used (the old one-element-per-iteration loop spent
    //      23 k cycles here, one exposed HBM/L2 latency per iteration) ----
    {
        constexpr int MAXB = (FS_MAX_N / FS_NB) * (FS_MAX_N / FS_NB + 1) / 2;     // 15 blocks: ALL loads are requested before any is used
        const int nb = nblk * (nblk + 1) / 2;
        double v[MAXB][2];
#pragma unroll
        for (int b = 0; b < MAXB; ++b) {
            int bi = 0;
            while ((bi + 1) * (bi + 2) / 2 <= b) ++bi;
            const int bj = b - bi * (bi + 1) / 2, c = bj * FS_NB + lane;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int r = bi * FS_NB + warp * 2 + half;
                double x = (b < nb && r < n && c <= r) ? Hg[(long)r * n_pad + c] : 0.0;
                if (r == c) x = (r < n) ? x + dadd[r] : 1.0;
                v[b][half] = x;
            }
        }
#pragma unroll
        for (int b = 0; b < MAXB; ++b)
            if (b < nb) {
#pragma unroll
                for (int half = 0; half < 2; ++half) SB[b * FS_BLK + (warp * 2 + half) * P + lane] = v[b][half];
            }
    }
    __syncthreads();
    tr();   // 1: loaded

    // (a) Cholesky of a diagonal block by one warp (see fs_potrf_warp).  Block 0 here; block kb + 1 is factored by warp 0 WHILE
    //     the other 15 warps finish the trailing update of step kb (lookahead, step (d) below).
    // H is the accumulator of the Gram kernel's atomics: it is left cleared for the next factor (no memset on that critical path) —
    // by the warps that idle during the last diagonal block's factorisation
    auto zero_H = [&](int first, int stride) {
        const int nbt = nblk * (nblk + 1) / 2;
        for (int idx = first; idx < nbt * FS_NB * FS_NB; idx += stride) {
            const int b = idx >> 10, rc = idx & 1023;
            int bi = 0;
            while ((bi + 1) * (bi + 2) / 2 <= b) ++bi;
            const int bj = b - bi * (bi + 1) / 2;
            Hg[(long)(bi * FS_NB + (rc >> 5)) * n_pad + bj * FS_NB + (rc & 31)] = 0.0;
        }
    };
    if (nblk == 1 && !write_back && warp != 0) zero_H(tid - 32, FS_T - 32);
    if (warp == 0) fs_potrf_warp(SB + fs_blk(0, 0), invd, lane, flag, 0);
    __syncthreads();
    tr();   // diag block 0 factored
    for (int kb = 0; kb < nblk; ++kb) {
        const double* D = SB + fs_blk(kb, kb);
        const int o = kb * FS_NB;
        const int mb = nblk - kb - 1;                            // block rows below
        if (mb > 0) {
            // (c) panel  L_ik = H_ik L_kk^-T  for the mb blocks below, in place: one thread per panel row (fs_trsm_row)
            double* LS = invd + n_pad;                           // row-scaled copy of L_kk (one spare block behind the reciprocals)
            for (int idx = tid; idx < (FS_NB + 7) * FS_NB; idx += FS_T) {      // + 7 rows that the last steps read and do not use: finite
                const int r = idx >> 5, c = idx & 31;
                LS[r * P + c] = (r < FS_NB && c < r) ? D[r * P + c] * invd[o + r] : 0.0;
            }
            __syncthreads();
            if (tid < mb * FS_NB) fs_trsm_row(SB + fs_blk(kb + 1 + (tid >> 5), kb) + (tid & 31) * P, LS, invd + o);
            __syncthreads();
            tr();   // panel done
            // (d) trailing update of the lower block triangle:  H_ij -= L_ik L_jk^T  on the fp64 tensor pipe (mma.sync m8n8k4), with
            //     LOOKAHEAD: first only the next diagonal block — the ten 8 x 8 tiles of its lower triangle, one warp each (8 + 8
            //     fragment loads and 8 DMMA in two chains; the SIMT form, 4 x 2 register tiles on 128 threads, took 1.7 k cycles) —
            //     then warp 0 factors it, register resident, while warps 1..15 update the other blocks.
            const int npairs = mb * (mb + 1) / 2;
            if (warp < 10) {
                int tm = 0;
                while ((tm + 1) * (tm + 2) / 2 <= warp) ++tm;
                const int tn = warp - tm * (tm + 1) / 2, g = lane >> 2, t = lane & 3;
                const double* A = SB + fs_blk(kb + 1, kb) + (tm * 8 + g) * P + t;
                const double* B = SB + fs_blk(kb + 1, kb) + (tn * 8 + g) * P + t;
                double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
#pragma unroll
                for (int ks = 0; ks < 8; ks += 2) { fs_dmma884(c0, c1, A[ks * 4], B[ks * 4]); fs_dmma884(e0, e1, A[ks * 4 + 4], B[ks * 4 + 4]); }
                double* Cb = SB + fs_blk(kb + 1, kb + 1) + (tm * 8 + g) * P + tn * 8 + 2 * t;
                if (tn * 8 + 2 * t <= tm * 8 + g) Cb[0] -= c0 + e0;
                if (tn * 8 + 2 * t + 1 <= tm * 8 + g) Cb[1] -= c1 + e1;
            }
            __syncthreads();
            tr();   // next diagonal block updated
            if (warp == 0) {
                fs_potrf_warp(SB + fs_blk(kb + 1, kb + 1), invd + o + FS_NB, lane, flag, kb + 1);
                if (trace && lane == 0) trace[48 + kb] = clock64();
            } else {
                // trailing update on the fp64 tensor pipe: task = (block pair, 8-row tile); the A fragment is loaded once and swept
                // over the pair's 8-column tiles (a diagonal pair needs the tiles on and below its diagonal only).  9 pairs at the
                // first step = 36 tasks on 15 warps: it stays inside the shadow of warp 0's potrf (the SIMT version — 4 x 4 register
                // tiles, 8 shared-memory loads per 16 fma — took 21.8 / 16.9 / 12.7 k cycles against the potrf's 11 k)
                {
                    const int g = lane >> 2, t = lane & 3;
                    // only `trail_warps` warps, none of them on warp 0's scheduler, issue the tensor-pipe work: with all 15 the
                    // fp64 pipe stays saturated and every fp64 instruction of the potrf chain queues behind it (trace: potrf
                    // 22 k cycles under a 15-warp update against 11.9 k alone — the potrf IS the critical path of the phase)
                    const int slot = (warp & 3) ? (warp >> 2) * 3 + (warp & 3) - 1 : -1;       // 0..11 for warps 1,2,3,5,6,7,...
                    for (int task = slot; slot >= 0 && slot < trail_warps && task < (npairs - 1) * 4; task += trail_warps) {
                        const int pr = 1 + (task >> 2), tm = task & 3;
                        int ti = 0;
                        while ((ti + 1) * (ti + 2) / 2 <= pr) ++ti;
                        const int tj = pr - ti * (ti + 1) / 2;
                        const double* A = SB + fs_blk(kb + 1 + ti, kb) + (tm * 8 + g) * P + t;
                        double a[8];
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) a[ks] = A[ks * 4];
                        const int n_tiles = (ti == tj) ? tm + 1 : 4;
                        for (int tn = 0; tn < n_tiles; ++tn) {
                            const double* B = SB + fs_blk(kb + 1 + tj, kb) + (tn * 8 + g) * P + t;
                            double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
#pragma unroll
                            for (int ks = 0; ks < 8; ks += 2) { fs_dmma884(c0, c1, a[ks], B[ks * 4]); fs_dmma884(e0, e1, a[ks + 1], B[ks * 4 + 4]); }
                            double* Cb = SB + fs_blk(kb + 1 + ti, kb + 1 + tj) + (tm * 8 + g) * P + tn * 8 + 2 * t;
                            const bool lower = ti != tj;
                            if (lower || tn * 8 + 2 * t <= tm * 8 + g) Cb[0] -= c0 + e0;
                            if (lower || tn * 8 + 2 * t + 1 <= tm * 8 + g) Cb[1] -= c1 + e1;
                        }
                    }
                }
                if (trace && tid == 32) trace[32 + 3 * kb] = clock64();
                // block column kb (L_kk and the panel below it) is final: it goes to global memory here, in the shadow of warp 0's
                // potrf, instead of in one 4.4 k-cycle pass after the last step
                for (int idx = tid - 32; idx < (mb + 1) * FS_NB * FS_NB; idx += FS_T - 32) {
                    const int bi = kb + (idx >> 10), rc = idx & 1023;
                    Lg[(size_t)(bi * (bi + 1) / 2 + kb) * (FS_NB * FS_NB) + rc] = SB[fs_blk(bi, kb) + (rc >> 5) * P + (rc & 31)];
                }
                if (trace && tid == 32) trace[33 + 3 * kb] = clock64();
                if (mb == 1 && !write_back) zero_H(tid - 32, FS_T - 32);     // last step: these warps have no trailing update left
                if (trace && tid == 32) trace[34 + 3 * kb] = clock64();
            }
            __syncthreads();
        }
        tr();   // trailing done
    }
    if (write_back) {   // debug: L back to global (row-major [n_pad][n_pad], zeros above the diagonal)
        for (int idx = tid; idx < n_pad * n_pad; idx += FS_T) {
            const int r = idx / n_pad, c = idx % n_pad;
            Hg[idx] = (c <= r) ? SB[fs_blk(r >> 5, c >> 5) + (r & 31) * P + (c & 31)] : 0.0;
        }
    }

    // ---- the factor goes back to global memory for solve_emit: the block triangle as it sits in shared memory (diagonal blocks:
    //      L_kk in the lower triangle, the strictly-lower part of L_kk^-1 transposed in the strict upper triangle), 32 x 32 blocks
    //      of pitch 32, and 1 / L_ii ----
    {   // (the block columns before the last one were written during the steps)
        const int kl = nblk - 1;
        for (int idx = tid; idx < FS_NB * FS_NB; idx += FS_T)
            Lg[(size_t)(kl * (kl + 1) / 2 + kl) * (FS_NB * FS_NB) + idx] = SB[fs_blk(kl, kl) + (idx >> 5) * P + (idx & 31)];
        for (int r = tid; r < n_pad; r += FS_T) invd_g[r] = invd[r];
    }
    tr();
}

// ---------------------------------------------------------------------------------------------------------
// Inverses of the diagonal blocks of L, one CTA per block, in place in the global copy of the factor: the strictly-lower part of
// L_kk^-1 goes, transposed, into the strict upper triangle of the block (the solve kernel reads its A fragments straight from that storage).  Column sweeps
// (lane = row): x_j = (delta_jc - acc_j) / L_jj; a warp carries its two columns (c, c + 16) through ONE sweep: the second chain is
// identically zero until j reaches c + 16.  Off the single-CTA kernel's critical path: all blocks at once, 4.5 k cycles.
__global__ void __launch_bounds__(FS_T, 1) inv_blocks_kernel(double* __restrict__ Lg, const double* __restrict__ invd_g) {
    __shared__ double Ds[FS_NB * (FS_NB + 1)];
    __shared__ double iv_s[FS_NB];
    constexpr int P = FS_NB + 1, NW = FS_T / 32;
    const int kb = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* Dg = Lg + (size_t)(kb * (kb + 1) / 2 + kb) * FS_NB * FS_NB;
    pdl_wait(); pdl_launch();                      // wait FIRST: solve_emit, launched by this trigger, reads the factor before its own wait
    for (int idx = tid; idx < FS_NB * FS_NB; idx += FS_T) Ds[(idx >> 5) * P + (idx & 31)] = Dg[idx];
    if (tid < FS_NB) iv_s[tid] = invd_g[kb * FS_NB + tid];
    __syncthreads();
    const int c1 = warp, c2 = warp + NW;
    double acc1 = 0.0, acc2 = 0.0, mine1 = 0.0, mine2 = 0.0;
    for (int j = c1; j < FS_NB; ++j) {
        const double cand1 = ((lane == c1) ? 1.0 : 0.0) - acc1, cand2 = ((lane == c2) ? 1.0 : 0.0) - acc2;
        const double iv = iv_s[j];
        const double x1 = __shfl_sync(0xffffffffu, cand1, j) * iv, x2 = __shfl_sync(0xffffffffu, cand2, j) * iv;
        if (lane == j) { mine1 = x1; mine2 = x2; }
        if (lane > j) {
            const double l = Ds[lane * P + j];
            acc1 = fma(l, x1, acc1); acc2 = fma(l, x2, acc2);
        }
    }
    if (lane > c1) Dg[c1 * FS_NB + lane] = mine1;            // strict upper triangle <- transposed strict lower of L^-1
    if (lane > c2) Dg[c2 * FS_NB + lane] = mine2;
}

// ---------------------------------------------------------------------------------------------------------
// General blocked path (factor.cu, systems larger than FS_MAX_N): factor diagonal block kb of H (row-major, ld) in place and emit
// the dense inverse of its Cholesky factor — the warp-resident potrf and the column-sweep inverse of this file instead of the first
// version's 1 024-thread kernel with three block barriers per pivot (26 us per block; 32 blocks per factor at BASELINE cfg4).
__global__ void __launch_bounds__(FS_T, 1) potrf_inv_general_kernel(double* __restrict__ H, int ld, int kb, double* __restrict__ Linv, int* flag) {
    constexpr int P = FS_NB + 1, NW = FS_T / 32;
    __shared__ double D[(FS_NB + 8) * P];           // + 8 rows of slack: the last pivots read L[j + k][j] for rows up to 38
    __shared__ double invd[FS_NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* blk = H + (size_t)kb * FS_NB * ld + (size_t)kb * FS_NB;
    pdl_wait(); pdl_launch();
    for (int idx = tid; idx < (FS_NB + 8) * FS_NB; idx += FS_T) {
        const int r = idx >> 5, c = idx & 31;
        D[r * P + c] = (r < FS_NB && c <= r) ? blk[(size_t)r * ld + c] : 0.0;
    }
    __syncthreads();
    if (warp == 0) fs_potrf_warp(D, invd, lane, flag, kb);
    __syncthreads();
    for (int idx = tid; idx < FS_NB * FS_NB; idx += FS_T) {
        const int r = idx >> 5, c = idx & 31;
        blk[(size_t)r * ld + c] = (c <= r) ? D[r * P + c] : 0.0;
    }
    // inverse by column sweeps (lane = row), two columns per warp: see inv_blocks_kernel
    const int c1 = warp, c2 = warp + NW;
    double acc1 = 0.0, acc2 = 0.0, mine1 = 0.0, mine2 = 0.0;
    for (int j = c1; j < FS_NB; ++j) {
        const double cand1 = ((lane == c1) ? 1.0 : 0.0) - acc1, cand2 = ((lane == c2) ? 1.0 : 0.0) - acc2;
        const double iv = invd[j];
        const double x1 = __shfl_sync(0xffffffffu, cand1, j) * iv, x2 = __shfl_sync(0xffffffffu, cand2, j) * iv;
        if (lane == j) { mine1 = x1; mine2 = x2; }
        if (lane > j) {
            const double l = D[lane * P + j];
            acc1 = fma(l, x1, acc1); acc2 = fma(l, x2, acc2);
        }
    }
    double* out = Linv + (size_t)kb * FS_NB * FS_NB;       // dense [row][column], zeros above the diagonal
    out[lane * FS_NB + c1] = mine1;
    out[lane * FS_NB + c2] = mine2;
}
int potrf_inv_general(double* H, int ld, int kb, double* Linv, int* flag, cudaStream_t st) {
    static_assert(FS_NB == UCE_NB && FS_T == 32 * 16, "two inverse columns per warp");
    return (int)launch_k(potrf_inv_general_kernel, dim3(1), dim3(FS_T), 0, st, 1, H, ld, kb, Linv, flag);
}

// ---------------------------------------------------------------------------------------------------------
// X = H^-1 Cp[:, cols] for one slab of SE_CW columns, then Q[j, cols] = X[n_pres + j, cols]  (see the file header).
//   forward   L Y = Cp[:, cols]        over all block rows
//   backward  L^T X = Y                from the last block row up to the block row of the first edit row
constexpr int SE_T = 256;
// The solve on the fp64 TENSOR pipe (mma.sync m8n8k4.f64).  Its SIMT predecessor issued two shared-memory loads per fma and was bound by
// shared-memory wavefronts and the latency of its dependent chains (the general path's twin measured 71 % of the shared-memory pipe,
// 12 % of the fp64 pipe: profiles/r02_solve_general_ncu.txt; cfg2 step 0.1315 ms with it, 0.1268 with this one).  A warp owns 8-row tiles: the A fragment of a
// tile (8 x 4 of a block of L, or of the inverse of a diagonal block read through its transposed-upper-triangle storage) is eight
// loads per thread and 32 x 32 x CW multiply-adds; the slab rows of the current block (B fragments) are read once per block step.
// Blocks sit in shared memory with pitch 36: the 64-bit fragment loads of a warp then touch every bank pair exactly twice.
template <int CW>
__global__ void __launch_bounds__(SE_T, 1)
solve_emit_dmma_kernel(const double* __restrict__ Lg, const double* __restrict__ invd_g, const float* __restrict__ Cp, int n, int n_pad,
                       int n_pres, int n_edit, int r_pad, int K, float* __restrict__ Q, float* __restrict__ Qt, float* __restrict__ Qt_hi,
                       float* __restrict__ Qt_lo) {
    extern __shared__ double smem_d[];
    constexpr int PS = 36, BLKS = FS_NB * PS, XL = CW + 4, NT = CW / 8, NW = SE_T / 32;
    const int nblk = n_pad / FS_NB, nb = nblk * (nblk + 1) / 2;
    double* SB = smem_d;
    double* invd = SB + nb * BLKS;
    double* XS = invd + n_pad;                    // [n_pad][XL]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int k0 = blockIdx.x * CW;
    // Launched (programmatically) while inv_blocks still runs: the slab, the reciprocals and the off-diagonal blocks of L come from
    // kernels that had finished before inv_blocks passed its own wait and are loaded right away; the diagonal blocks (inverted in
    // place by inv_blocks) after the wait.  The apply's second kernel is launched by the trigger below and prefetches beside us.
    pdl_launch();
    for (int idx = tid; idx < n_pad * CW; idx += SE_T) {
        const int r = idx / CW, c = idx % CW;
        XS[r * XL + c] = (r < n && k0 + c < K) ? (double)Cp[(long)r * K + k0 + c] : 0.0;
    }
    auto is_diag = [](int b) { int kb = 0; while ((kb + 1) * (kb + 2) / 2 <= b) ++kb; return b == kb * (kb + 1) / 2 + kb; };
    const double2* L2 = reinterpret_cast<const double2*>(Lg);
    const int n2 = nb * FS_NB * FS_NB / 2;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {         // 0: off-diagonal blocks (before the wait), 1: diagonal blocks
        if (pass == 1) pdl_wait();
        for (int i0 = tid; i0 < n2; i0 += 8 * SE_T) {
            double2 v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int idx = i0 + q * SE_T;
                v[q] = (idx < n2 && is_diag((2 * idx) >> 10) == (pass == 1)) ? L2[idx] : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int idx = i0 + q * SE_T;
                if (idx < n2) {
                    const int e = 2 * idx, b = e >> 10, rc = e & 1023;
                    if (is_diag(b) == (pass == 1)) {
                        double* d = SB + b * BLKS + (rc >> 5) * PS + (rc & 31);
                        d[0] = v[q].x; d[1] = v[q].y;
                    }
                }
            }
        }
        if (pass == 0) for (int r = tid; r < n_pad; r += SE_T) invd[r] = invd_g[r];
    }
    __syncthreads();
    auto blk = [&](int bi, int bj) { return SB + (bi * (bi + 1) / 2 + bj) * BLKS; };
    const int kb_e = n_pres / FS_NB;
    // fragments: A[g][4 ks + t], B[4 ks + t][g], C[g][2 t + {0, 1}]
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {      // 0 forward  L Y = Cp;  1 backward  L^T X = Y, down to the block row of the first edit row
#pragma unroll 1
        for (int kb = phase == 0 ? 0 : nblk - 1; phase == 0 ? kb < nblk : kb >= kb_e; kb += phase == 0 ? 1 : -1) {
            const int o = kb * FS_NB;
            const double* D = blk(kb, kb);
            double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
            const int dm = warp & 3, dnt = warp >> 2;            // diagonal step: 8-row tile, 8-column tile
            if (dnt < NT) {
                const int r = dm * 8 + g;
                double a[8], b[8];
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const int j = ks * 4 + t;
                    // forward: (L^-1)[r][j], stored transposed in the strict upper triangle; backward: (L^-T)[r][j] = (L^-1)[j][r]
                    const double off = phase == 0 ? D[j * PS + r] : D[r * PS + j];
                    const bool strict = phase == 0 ? (j < r) : (r < j);
                    a[ks] = strict ? off : (j == r ? invd[o + r] : 0.0);
                    b[ks] = XS[(o + j) * XL + dnt * 8 + g];
                }
#pragma unroll
                for (int ks = 0; ks < 8; ks += 2) { fs_dmma884(c0, c1, a[ks], b[ks]); fs_dmma884(e0, e1, a[ks + 1], b[ks + 1]); }
            }
            __syncthreads();                                     // every warp has read the old X_k
            if (dnt < NT) {
                XS[(o + dm * 8 + g) * XL + dnt * 8 + 2 * t] = c0 + e0;
                XS[(o + dm * 8 + g) * XL + dnt * 8 + 2 * t + 1] = c1 + e1;
            }
            __syncthreads();
            const int i_lo = phase == 0 ? kb + 1 : kb_e, i_hi = phase == 0 ? nblk : kb;       // the other blocks of this column
            if (i_lo < i_hi) {
                double yb[NT][8];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) yb[nt][ks] = XS[(o + ks * 4 + t) * XL + nt * 8 + g];
                for (int tile = warp; tile < (i_hi - i_lo) * 4; tile += NW) {
                    const int i = i_lo + (tile >> 2), m = tile & 3;
                    // forward  X_i -= L(i, kb) Y_k:  A[r][j] = block(i, kb)[r][j];  backward  Y_i -= L(kb, i)^T X_k:  A[r][j] = block(kb, i)[j][r]
                    const double* A = phase == 0 ? blk(i, kb) + (m * 8 + g) * PS + t : blk(kb, i) + t * PS + m * 8 + g;
                    const int ks_stride = phase == 0 ? 4 : 4 * PS;
                    double a[8];
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) a[ks] = A[ks * ks_stride];
                    double acc[NT][2];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = 0.0; acc[nt][1] = 0.0; }
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) fs_dmma884(acc[nt][0], acc[nt][1], a[ks], yb[nt][ks]);
                    double* x = XS + (i * FS_NB + m * 8 + g) * XL + 2 * t;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) { x[nt * 8] -= acc[nt][0]; x[nt * 8 + 1] -= acc[nt][1]; }
                }
                __syncthreads();
            }
        }
    }
    // ---- emit: Q [r_pad, K] row-major, Qt [K, r_pad] and its tf32 split; rows j >= n_edit (rank padding) are exact zeros ----
    for (int idx = tid; idx < r_pad * CW; idx += SE_T) {
        const int cc = idx / r_pad, j = idx % r_pad;
        if (k0 + cc >= K) continue;
        const float v = (j < n_edit) ? (float)XS[(n_pres + j) * XL + cc] : 0.f;
        const long tq = (long)(k0 + cc) * r_pad + j;
        const float h = fs_tf32_hi(v);
        Qt[tq] = v; Qt_hi[tq] = h; Qt_lo[tq] = v - h;
        Q[(long)j * K + k0 + cc] = v;
    }
}

// E = G_e - C_e with its tf32 split, and the packed concept rows; one block per row.
__device__ void pack_rows_split_block(int r, const float* __restrict__ C, const float* __restrict__ G, const int* __restrict__ src, int n_act,
                                      int n_pres, int rank_pad, int K, float* __restrict__ Cp, float* __restrict__ E,
                                      float* __restrict__ E_hi, float* __restrict__ E_lo) {
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

bool factor_small_applicable(const uce_ws* ws, int n, int n_edit, bool dual) {
    return dual && n <= FS_MAX_N && n_edit <= FS_MAX_RHS && !ws->force_general;
}

// Preconditions: ws->h_src_idx / h_diag_add staged and copied, flag cleared, n_edit > 0.
int factor_small(uce_ws* ws, const float* C, const float* G, int n, int n_pres, int n_edit, cudaStream_t st, int* launches) {
    const int K = ws->K;
    const int n_pad = round_up(n, FS_NB);
    ws->sys_n = n_pad;
    // H accumulates the Gram kernel's atomics and must start from zero: chol_small_kernel clears what it read, so in steady state no
    // memset sits between the table kernel and the Gram kernel (it would also break their programmatic launch edge)
    // (the whole region any size of this path can touch, FS_MAX_N^2, so that a later call with another n_pad finds zeros too)
    if (ws->H_dirty || ws->debug) {
        const size_t sm = (size_t)ws->sys_max * ws->sys_max, fm = (size_t)FS_MAX_N * FS_MAX_N;
        UCE_CUDA(cudaMemsetAsync(ws->H, 0, (sm < fm ? sm : fm) * sizeof(double), st));
    }
    const int nt = n_pad / FS_NB, n_pairs = nt * (nt + 1) / 2, n_pack = n + (ws->rank_pad - n_edit);
    FsSrc srcp;
    memcpy(srcp.v, ws->h_src_idx, (size_t)n * sizeof(int));
    for (int i = n; i < FS_MAX_N; ++i) srcp.v[i] = 0;
    UCE_CUDA(launch_k(gram_pack_kernel, dim3(n_pairs * ceil_div(K, FS_KSPLIT) + n_pack), dim3(256), 0, st, 1, C, srcp, n, K, ws->H, n_pad,
                      n_pairs, G, n_pres, ws->rank_pad, n_pack, ws->Cp, ws->E, ws->E_hi, ws->E_lo));
    ++*launches;
    // E and its split are complete (pack kernel) and the many-CTA Gram kernel is behind us: from here on the factor is one CTA wide,
    // the point where uce_edit_dev_f32 lets the apply's first kernel (which needs E only) start on its own stream
    if (ws->want_ev_E) { UCE_CUDA(cudaEventRecord(ws->ev_E, st)); ws->ev_E_recorded = 1; }
    if (ws->debug) {
        if (!ws->Hcopy) UCE_CUDA(cudaMalloc(&ws->Hcopy, (size_t)ws->sys_max * ws->sys_max * sizeof(double)));
        UCE_CUDA(cudaMemcpyAsync(ws->Hcopy, ws->H, (size_t)n_pad * n_pad * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    // + 8 rows of slack: the last pivots of a diagonal block read L[j + k][j] for rows up to 38 (columns that do not exist, results unused)
    const size_t smem_c = ((size_t)(nt * (nt + 1) / 2) * FS_BLK + n_pad + FS_BLK + 8 * (FS_NB + 1)) * sizeof(double);
    static thread_local size_t conf_dev[64] = {0};        // per-device function attribute
    size_t& conf_c = conf_dev[ws->device & 63];
    if (conf_c < smem_c) {
        UCE_CUDA(cudaFuncSetAttribute(chol_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        conf_c = smem_c;
    }
    long long* trace = nullptr;
    const char* trace_path = getenv("UCE_CHOL_TRACE");
    if (trace_path) { UCE_CUDA(cudaMalloc(&trace, 64 * sizeof(long long))); UCE_CUDA(cudaMemsetAsync(trace, 0, 64 * sizeof(long long), st)); }
    static_assert(FS_MAX_N == 160, "ws->Lsmall is sized for 15 blocks + 160 reciprocals");
    static const int trail_warps = [] { const char* e = getenv("UCE_CHOL_TRAIL_WARPS"); const int v = e ? atoi(e) : 6; return v < 1 ? 1 : (v > 12 ? 12 : v); }();
    double* Lg = ws->Lsmall; double* invd_g = ws->Lsmall + 15 * 1024;
    UCE_CUDA(launch_k(chol_small_kernel, dim3(1), dim3(FS_T), smem_c, st, 1, ws->H, n, n_pad, (const double*)ws->diag_add, Lg, invd_g, (int)ws->debug, ws->flag, trace, trail_warps));
    ++*launches;
    ws->H_dirty = ws->debug ? 1 : 0;
    // the apply's first kernel goes to its own stream now (the host encodes it while the single-CTA factor runs); this stream waits for
    // it BEFORE the last two kernels, so that the apply's second kernel directly follows solve_emit and launches programmatically
    if (ws->hook_after_E) {
        ws->mode = 1;                                  // this path is the dual system; the apply's planner asks
        int rc = ws->hook_after_E(ws->hook_ctx);
        if (rc) return rc;
        if (ws->hook_done) UCE_CUDA(cudaStreamWaitEvent(st, ws->hook_done, 0));
    }
    if (trace) {   // debugging aid: phase boundaries of the single factor CTA (synchronises)
        long long h[64];
        UCE_CUDA(cudaStreamSynchronize(st));
        UCE_CUDA(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
        UCE_CUDA(cudaFree(trace));
        if (FILE* f = fopen(trace_path, "w")) {
            for (int i = 0; i < 32 && h[i]; ++i) fprintf(f, "%d %lld %lld\n", i, h[i] - h[0], i ? h[i] - h[i - 1] : 0LL);
            // inside the lookahead phases: 32 + 3 kb: warp 1 done with the trailing update, + 1: with the write-out, + 2: with clearing H; 48 + kb: warp 0's potrf done
            for (int i = 32; i < 64; ++i) if (h[i]) fprintf(f, "%d %lld\n", i, h[i] - h[0]);
            fclose(f);
        }
    }
    UCE_CUDA(launch_k(inv_blocks_kernel, dim3(nt), dim3(FS_T), 0, st, 1, Lg, (const double*)invd_g));
    ++*launches;
    // solve: 8 columns of Cp per CTA (cfg2 step 0.1268 ms; 16 columns, UCE_SOLVE_EMIT=16: 0.1299)
    static const int se_cw = [] { const char* e = getenv("UCE_SOLVE_EMIT"); return (e && !strcmp(e, "16")) ? 16 : 8; }();
    {
        const int cw = se_cw;
        const size_t smem_d = ((size_t)(nt * (nt + 1) / 2) * FS_NB * 36 + n_pad + (size_t)n_pad * (cw + 4)) * sizeof(double);
        static thread_local size_t conf_dev_d[2][64] = {{0}, {0}};
        size_t& conf_d = conf_dev_d[cw == 16][ws->device & 63];
        if (conf_d < smem_d) {
            if (cw == 16) UCE_CUDA(cudaFuncSetAttribute(solve_emit_dmma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));
            else UCE_CUDA(cudaFuncSetAttribute(solve_emit_dmma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));
            conf_d = smem_d;
        }
        if (cw == 16)
            UCE_CUDA(launch_k(solve_emit_dmma_kernel<16>, dim3(ceil_div(K, 16)), dim3(SE_T), smem_d, st, 1, (const double*)Lg, (const double*)invd_g, (const float*)ws->Cp,
                              n, n_pad, n_pres, n_edit, ws->rank_pad, K, ws->Q, ws->Qt, ws->Qt_hi, ws->Qt_lo));
        else
            UCE_CUDA(launch_k(solve_emit_dmma_kernel<8>, dim3(ceil_div(K, 8)), dim3(SE_T), smem_d, st, 1, (const double*)Lg, (const double*)invd_g, (const float*)ws->Cp,
                              n, n_pad, n_pres, n_edit, ws->rank_pad, K, ws->Q, ws->Qt, ws->Qt_hi, ws->Qt_lo));
    }
    ++*launches;
    return 0;
}

}  // namespace uce
