// Fused attention of the SD U-Net transformer blocks on tcgen05 (flash-style, no score matrix in HBM):
//
//     O[b, q, h, :] = softmax(Q[b, q, h, :] K[b, :, h, :]^T / sqrt(dh)) V[b, :, h, :]
//
// replaces the three-kernel sequence  S = QK^T (fp32 to HBM) -> softmax -> PV  that the first version of the engine
// used (1 GB of fp32 scores per 64x64 self-attention).
//
// One CTA per (128-query tile, head, image), looping over 128-key tiles; TWO CTAs share an SM (256 TMEM columns and
// <= 98 KB of shared memory each), so the exp-bound softmax of one overlaps the MMAs / TMEM traffic of the other:
//   warp 0      TMA: Q once, then K tile [128 keys x dhp] and V^T tile [dhp x 128 keys] per iteration
//   warp 1      MMA: S = Q K^T into TMEM columns [0,128); O += P V with P read from TENSOR MEMORY — the bf16 P tile is
//               written by the softmax warps over the first 64 columns of S itself (the tensor pipe executes
//               PV(j) before QK(j+1), so the alias is safe)
//   warps 2-5   softmax: thread = query row = TMEM lane.  Pass 1 reads S for the row maximum, pass 2 re-reads it and emits
//               P = exp2(s*c - m*c) (one FFMA + one MUFU per score).  The running maximum m is only raised when the row
//               maximum exceeds it by more than 2^8 (the final 1/l normalisation cancels the stale scale exactly), so the
//               O accumulator in TMEM is rescaled a handful of times per row instead of once per key tile.
// Operands bf16, accumulation fp32.  Heads are padded to dhp in {64, 128, 192} (zero columns); keys beyond Lk are masked.
#include "tc_common.cuh"
#include "unet_attn.h"
#include <cuda_bf16.h>

namespace uce {
using namespace tc;

constexpr int FA_THREADS = 192;
constexpr int FA_BM = 128, FA_BN = 128;
constexpr uint32_t FA_S_COL = 0, FA_P_COL = 0, FA_O_COL = 128;     // S [0,128) fp32, P aliases [0,64) as bf16 | O [128,128+dhp)

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// A operand (bf16) from tensor memory, B from shared memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

struct FaSmem { int q_bytes, kv_bytes, stages, k_off, v_off, bar_off, total; };
__host__ __device__ inline FaSmem fa_smem_layout(int dhp) {
    FaSmem s;
    s.q_bytes = FA_BM * dhp * 2;            // dhp/64 atoms of [128 x 64] bf16
    s.kv_bytes = FA_BN * dhp * 2;           // K tile [128 keys x dhp]  /  V^T tile [dhp x 128 keys]: same size
    s.stages = dhp <= 64 ? 2 : 1;           // two CTAs per SM must fit: 80 KB (dhp 64, 2 stages) / 96 KB (dhp 128, 1 stage)
    s.k_off = s.q_bytes;
    s.v_off = s.k_off + s.stages * s.kv_bytes;
    s.bar_off = s.v_off + s.stages * s.kv_bytes;
    s.total = s.bar_off + 256;
    return s;
}
__host__ __device__ inline uint32_t fa_tmem_cols(int dhp) { return 128 + dhp <= 256 ? 256u : 512u; }

__global__ void __launch_bounds__(FA_THREADS, 2) unet_attn_kernel(const __grid_constant__ AttnDesc g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    pdl_launch();
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int dhp = g.dhp;
    const FaSmem L = fa_smem_layout(dhp);
    const int NS = L.stages;
    const uint32_t bars = base + L.bar_off;
    const uint32_t bar_q = bars;
    auto bar_k_full = [&](int s) { return bars + 8u * (1 + s); };
    auto bar_k_empty = [&](int s) { return bars + 8u * (3 + s); };
    auto bar_v_full = [&](int s) { return bars + 8u * (5 + s); };
    auto bar_v_empty = [&](int s) { return bars + 8u * (7 + s); };
    const uint32_t bar_s_full = bars + 8u * 9, bar_p_full = bars + 8u * 10, bar_o_done = bars + 8u * 11;
    const uint32_t tmem_slot = bars + 8u * 12;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
    const int n_kt = (g.Lk + FA_BN - 1) / FA_BN;
    const int n_at = dhp / 64;

    if (threadIdx.x == 0) {
        mbar_init(bar_q, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(bar_k_full(s), 1); mbar_init(bar_k_empty(s), 1); mbar_init(bar_v_full(s), 1); mbar_init(bar_v_empty(s), 1); }
        mbar_init(bar_s_full, 1); mbar_init(bar_p_full, 4); mbar_init(bar_o_done, 1);
        mbar_fence_init();
        tma_prefetch_desc(&g.tmQ); tma_prefetch_desc(&g.tmK); tma_prefetch_desc(&g.tmV);
    }
    if (warp == 1) tmem_alloc(tmem_slot, fa_tmem_cols(dhp));
    fence_before();
    __syncthreads();
    fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_wait();                                        // q, k, v^T are written by the projections launched just before

    auto k_st = [&](int s) { return base + (uint32_t)(L.k_off + s * L.kv_bytes); };
    auto v_st = [&](int s) { return base + (uint32_t)(L.v_off + s * L.kv_bytes); };

    if (warp == 0) {
        // all lanes run the loop and the barrier waits; one elected lane issues (see elect_one() in tc_common.cuh)
        if (elect_one()) {
            mbar_arrive_expect_tx(bar_q, (uint32_t)L.q_bytes);
            for (int a = 0; a < n_at; ++a) tma_load_4d(base + a * 16384, &g.tmQ, bar_q, a * 64, q_tile * FA_BM, head, img);
        }
        __syncwarp();
        for (int j = 0; j < n_kt; ++j) {
            const int s = j % NS;
            const uint32_t ph = (uint32_t)(((j / NS) & 1) ^ 1);
            mbar_wait(bar_k_empty(s), ph);
            __syncwarp();
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_k_full(s), (uint32_t)L.kv_bytes);
                for (int a = 0; a < n_at; ++a) tma_load_4d(k_st(s) + a * 16384, &g.tmK, bar_k_full(s), a * 64, j * FA_BN, head, img);
            }
            mbar_wait(bar_v_empty(s), ph);
            __syncwarp();
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_v_full(s), (uint32_t)L.kv_bytes);
                for (int a = 0; a < 2; ++a) tma_load_4d(v_st(s) + a * (dhp * 128), &g.tmV, bar_v_full(s), j * FA_BN + a * 64, 0, head, img);
            }
        }
    } else if (warp == 1) {
        // the whole warp runs the loop and the barrier waits (converged); ONE elected lane issues the MMAs and commits
        const uint32_t idesc_s = idesc_bf16(FA_BM, FA_BN), idesc_o = idesc_bf16(FA_BM, dhp);
        mbar_wait(bar_q, 0);
        auto issue_qk = [&](int j) {
            const int s = j % NS;
            mbar_wait(bar_k_full(s), (uint32_t)((j / NS) & 1));
            fence_after();
            __syncwarp();
            if (elect_one()) {
                for (int a = 0; a < n_at; ++a) {
                    const uint64_t qd = umma_desc_sw128(base + a * 16384), kd = umma_desc_sw128(k_st(s) + a * 16384);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + FA_S_COL, qd + 2u * k, kd + 2u * k, idesc_s, (a | k) != 0);
                }
                umma_commit(bar_k_empty(s));
                umma_commit(bar_s_full);          // also: every earlier PV is complete (in-order pipe) -> O may be rescaled
            }
        };
        issue_qk(0);
        for (int j = 0; j < n_kt; ++j) {
            const int s = j % NS;
            mbar_wait(bar_v_full(s), (uint32_t)((j / NS) & 1));
            mbar_wait(bar_p_full, (uint32_t)(j & 1));
            fence_after();
            __syncwarp();
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {          // 128 keys = 8 x 16; P: 8 TMEM columns per step; V^T: two 64-key atoms
                    const uint64_t vd = umma_desc_sw128(v_st(s) + (ks >> 2) * (dhp * 128)) + 2u * (ks & 3);
                    umma_bf16_ts(tmem_base + FA_O_COL, tmem_base + FA_P_COL + 8u * ks, vd, idesc_o, (j | ks) != 0);
                }
                umma_commit(bar_v_empty(s));
                if (j + 1 >= n_kt) umma_commit(bar_o_done);
            }
            if (j + 1 < n_kt) issue_qk(j + 1);            // overwrites S (and the P alias) strictly after PV(j) in the pipe
        }
    } else {
        // ---- softmax / correction / epilogue: thread = query row ----
        const int qd = warp & 3;
        const uint32_t lane_base = (uint32_t)(32 * qd) << 16;
        const float c = g.scale * 1.44269504088896340736f;      // exp2 domain
        const uint32_t NEG_INF = 0xff800000u;
        float m = -INFINITY, l = 0.f;                             // m in raw score units
        for (int j = 0; j < n_kt; ++j) {
            mbar_wait(bar_s_full, (uint32_t)(j & 1));
            fence_after();
            const int kbase = j * FA_BN;
            const bool tail = kbase + FA_BN > g.Lk;
            // ---- pass 1: row maximum ----
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t a[32], b[32];
                tmem_ld32_nowait(tmem_base + lane_base + FA_S_COL + 64u * hf, a);
                tmem_ld32_nowait(tmem_base + lane_base + FA_S_COL + 64u * hf + 32u, b);
                tmem_wait_ld();
                if (tail) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (kbase + 64 * hf + i >= g.Lk) a[i] = NEG_INF;
                        if (kbase + 64 * hf + 32 + i >= g.Lk) b[i] = NEG_INF;
                    }
                }
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    mx0 = fmaxf(mx0, __uint_as_float(a[i])); mx1 = fmaxf(mx1, __uint_as_float(a[i + 1]));
                    mx2 = fmaxf(mx2, __uint_as_float(b[i])); mx3 = fmaxf(mx3, __uint_as_float(b[i + 1]));
                }
            }
            const float rmax = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            const bool need = (rmax - m) * c > 8.f;                // first tile: m = -inf -> true
            float alpha = 1.f;
            if (need) { alpha = ex2_fast((m - rmax) * c); m = rmax; }
            if (j > 0 && __any_sync(0xffffffffu, need)) {          // PV(j-1) is complete (see issue_qk): O may be touched
                for (int c0 = 0; c0 < dhp; c0 += 32) {
                    uint32_t o[32];
                    tmem_ld32(tmem_base + lane_base + FA_O_COL + (uint32_t)c0, o);
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st32(tmem_base + lane_base + FA_O_COL + (uint32_t)c0, o);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");      // o[] is loaded into again by the next iteration
                }
            }
            l *= alpha;
            const float nmc = -m * c;
            // ---- pass 2: P = exp2(s c - m c) -> bf16 -> TMEM, over the S columns already consumed ----
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t a[32], b[32], pk[32];
                tmem_ld32_nowait(tmem_base + lane_base + FA_S_COL + 64u * hf, a);
                tmem_ld32_nowait(tmem_base + lane_base + FA_S_COL + 64u * hf + 32u, b);
                tmem_wait_ld();
                if (tail) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (kbase + 64 * hf + i >= g.Lk) a[i] = NEG_INF;
                        if (kbase + 64 * hf + 32 + i >= g.Lk) b[i] = NEG_INF;
                    }
                }
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float p0 = ex2_fast(fmaf(__uint_as_float(a[i]), c, nmc)), p1 = ex2_fast(fmaf(__uint_as_float(a[i + 1]), c, nmc));
                    const float p2 = ex2_fast(fmaf(__uint_as_float(b[i]), c, nmc)), p3 = ex2_fast(fmaf(__uint_as_float(b[i + 1]), c, nmc));
                    s0 += p0; s1 += p1; s2 += p2; s3 += p3;
                    const __nv_bfloat162 pa = __floats2bfloat162_rn(p0, p1), pb = __floats2bfloat162_rn(p2, p3);
                    pk[i / 2] = *reinterpret_cast<const uint32_t*>(&pa);
                    pk[16 + i / 2] = *reinterpret_cast<const uint32_t*>(&pb);
                }
                tmem_st32(tmem_base + lane_base + FA_P_COL + 32u * hf, pk);
                if (hf == 0) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");     // pk[] is rewritten by the second half (the store after it is waited for below)
            }
            l += (s0 + s1) + (s2 + s3);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p_full);
        }
        // ---- epilogue ----
        mbar_wait(bar_o_done, 0);
        fence_after();
        const int row = q_tile * FA_BM + 32 * qd + lane;
        const float inv = 1.f / l;
        __nv_bfloat16* op = g.out + ((long)img * g.L + row) * g.ldo + (long)head * dhp;
        for (int c0 = 0; c0 < dhp; c0 += 32) {
            uint32_t o[32];
            tmem_ld32(tmem_base + lane_base + FA_O_COL + (uint32_t)c0, o);
            if (row < g.L) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    uint4 u;
                    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                    for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(__uint_as_float(o[8 * jj + 2 * e]) * inv, __uint_as_float(o[8 * jj + 2 * e + 1]) * inv);
                    *reinterpret_cast<uint4*>(op + c0 + 8 * jj) = u;
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) { fence_after(); tmem_dealloc(tmem_base, fa_tmem_cols(dhp)); }
}

// ------------------------------------------------------------------------------------------ host
static int fa_encode(CUtensorMap* m, const void* ptr, const long (&dims)[4], const long (&str)[4], const int (&box)[4]) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return -1;
    cuuint64_t gd[4]; cuuint64_t gs[3]; cuuint32_t bx[4]; cuuint32_t es[4] = {1, 1, 1, 1};
    for (int i = 0; i < 4; ++i) { gd[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; }
    for (int i = 1; i < 4; ++i) gs[i - 1] = (cuuint64_t)str[i] * 2;
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -1;
}

bool attn_fused_supported(int dhp) { return dhp == 64 || dhp == 128 || dhp == 192; }

int attn_desc_make(AttnDesc* g, const void* q, const void* k, const void* vt, void* out, int NB, int heads, int dhp, long L, int Lk, int Lkp,
                   float scale, long ldq, long ldk) {
    memset(g, 0, sizeof(*g));
    const long HD = (long)heads * dhp;
    if (ldq <= 0) ldq = HD;
    if (ldk <= 0) ldk = HD;
    g->NB = NB; g->heads = heads; g->dhp = dhp; g->L = (int)L; g->Lk = Lk; g->scale = scale;
    g->out = (__nv_bfloat16*)out; g->ldo = HD;
    {
        const long dims[4] = {dhp, L, heads, NB}, str[4] = {1, ldq, dhp, L * ldq};
        const int box[4] = {64, FA_BM, 1, 1};
        if (fa_encode(&g->tmQ, q, dims, str, box)) return -1;
    }
    {
        const long dims[4] = {dhp, Lk, heads, NB}, str[4] = {1, ldk, dhp, (long)Lk * ldk};
        const int box[4] = {64, FA_BN, 1, 1};
        if (fa_encode(&g->tmK, k, dims, str, box)) return -1;
    }
    {   // V^T [NB][heads*dhp][Lkp]: a B-operand atom is [dhp rows x 64 keys]
        const long dims[4] = {Lk, dhp, heads, NB}, str[4] = {1, Lkp, (long)dhp * Lkp, HD * Lkp};
        const int box[4] = {64, dhp, 1, 1};
        if (fa_encode(&g->tmV, vt, dims, str, box)) return -1;
    }
    return 0;
}

int attn_launch(const AttnDesc& g, cudaStream_t st) {
    const FaSmem L = fa_smem_layout(g.dhp);
    const int smem = L.total + 1024;
    int dev = 0;
    cudaGetDevice(&dev);                                   // the opt-in shared-memory size is a per-device function attribute
    static thread_local int configured_dev[64] = {0};
    int& configured = configured_dev[dev & 63];
    if (configured < smem) {
        cudaError_t e = cudaFuncSetAttribute(unet_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    dim3 grid((g.L + FA_BM - 1) / FA_BM, g.heads, g.NB);
    cudaError_t e = launch_k(unet_attn_kernel, grid, dim3(FA_THREADS), (size_t)smem, st, 1, g);
    return (int)(e == cudaSuccess ? cudaGetLastError() : e);
}

}  // namespace uce
