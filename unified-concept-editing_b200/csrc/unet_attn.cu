// Fused attention of the SD U-Net transformer blocks on tcgen05 (flash-style, no score matrix in HBM):
//
//     O[b, q, h, :] = softmax(Q[b, q, h, :] K[b, :, h, :]^T / sqrt(dh)) V[b, :, h, :]
//
// replaces the three-kernel sequence  S = QK^T (fp32 to HBM) -> softmax -> PV  that the first version of the engine
// used (profiles/r01_unet_launches.txt: 1 GB of fp32 scores per 64x64 self-attention).
//
// One CTA per (128-query tile, head, image); loop over 128-key tiles:
//   warp 0      TMA: Q once, then K tile [128 keys x dhp] and V^T tile [dhp x 128 keys] per iteration (2 stages)
//   warp 1      MMA: S[buf] = Q K^T into one of two 128-column TMEM accumulators (so QK of tile j+1 overlaps the
//               softmax of tile j), then O += P V with P taken from TENSOR MEMORY (bf16, written by the softmax warps)
//   warps 2-5   softmax: thread = query row = TMEM lane; online max / sum in the exp2 domain, P -> TMEM, rescale of the
//               O accumulator in TMEM only when some row maximum in the warp grew
// Operands bf16, accumulation fp32.  Heads are padded to dhp in {64, 128} (zero columns); keys beyond Lk are masked.
#include "tc_common.cuh"
#include "unet_attn.h"
#include <cuda_bf16.h>

namespace uce {
using namespace tc;

constexpr int FA_THREADS = 192;
constexpr int FA_BM = 128, FA_BN = 128;
constexpr uint32_t FA_S_COL = 0, FA_O_COL = 256, FA_P_COL = 384;     // S0 [0,128) S1 [128,256) | O [256,256+dhp) | P [384,448)

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
// A operand (bf16) from tensor memory, B from shared memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

struct FaSmem { int q_bytes, kv_bytes, k_off, v_off, bar_off, total; };
__host__ __device__ inline FaSmem fa_smem_layout(int dhp) {
    FaSmem s;
    s.q_bytes = FA_BM * dhp * 2;            // dhp/64 atoms of [128 x 64] bf16
    s.kv_bytes = FA_BN * dhp * 2;           // K tile [128 keys x dhp]  /  V^T tile [dhp x 128 keys]: same size
    s.k_off = s.q_bytes;
    s.v_off = s.k_off + 2 * s.kv_bytes;
    s.bar_off = s.v_off + 2 * s.kv_bytes;
    s.total = s.bar_off + 256;
    return s;
}

__global__ void __launch_bounds__(FA_THREADS, 1) unet_attn_kernel(const __grid_constant__ AttnDesc g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int dhp = g.dhp;
    const FaSmem L = fa_smem_layout(dhp);
    const uint32_t bars = base + L.bar_off;
    const uint32_t bar_q = bars;
    auto bar_k_full = [&](int s) { return bars + 8u * (1 + s); };
    auto bar_k_empty = [&](int s) { return bars + 8u * (3 + s); };
    auto bar_v_full = [&](int s) { return bars + 8u * (5 + s); };
    auto bar_v_empty = [&](int s) { return bars + 8u * (7 + s); };
    auto bar_s_full = [&](int b) { return bars + 8u * (9 + b); };
    auto bar_s_empty = [&](int b) { return bars + 8u * (11 + b); };
    const uint32_t bar_p_full = bars + 8u * 13, bar_o_ready = bars + 8u * 14;
    const uint32_t tmem_slot = bars + 8u * 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
    const int n_kt = (g.Lk + FA_BN - 1) / FA_BN;
    const int n_at = dhp / 64;

    if (threadIdx.x == 0) {
        mbar_init(bar_q, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_k_full(s), 1); mbar_init(bar_k_empty(s), 1); mbar_init(bar_v_full(s), 1); mbar_init(bar_v_empty(s), 1);
            mbar_init(bar_s_full(s), 1); mbar_init(bar_s_empty(s), 4);
        }
        mbar_init(bar_p_full, 4); mbar_init(bar_o_ready, 1);
        mbar_fence_init();
        tma_prefetch_desc(&g.tmQ); tma_prefetch_desc(&g.tmK); tma_prefetch_desc(&g.tmV);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    fence_before();
    __syncthreads();
    fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    auto k_st = [&](int s) { return base + (uint32_t)(L.k_off + s * L.kv_bytes); };
    auto v_st = [&](int s) { return base + (uint32_t)(L.v_off + s * L.kv_bytes); };

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_q, (uint32_t)L.q_bytes);
            for (int a = 0; a < n_at; ++a) tma_load_4d(base + a * 16384, &g.tmQ, bar_q, a * 64, q_tile * FA_BM, head, img);
            for (int j = 0; j < n_kt; ++j) {
                const int s = j & 1;
                const uint32_t ph = (uint32_t)(((j >> 1) & 1) ^ 1);
                mbar_wait(bar_k_empty(s), ph);
                mbar_arrive_expect_tx(bar_k_full(s), (uint32_t)L.kv_bytes);
                for (int a = 0; a < n_at; ++a) tma_load_4d(k_st(s) + a * 16384, &g.tmK, bar_k_full(s), a * 64, j * FA_BN, head, img);
                mbar_wait(bar_v_empty(s), ph);
                mbar_arrive_expect_tx(bar_v_full(s), (uint32_t)L.kv_bytes);
                for (int a = 0; a < 2; ++a) tma_load_4d(v_st(s) + a * (dhp * 128), &g.tmV, bar_v_full(s), j * FA_BN + a * 64, 0, head, img);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = idesc_bf16(FA_BM, FA_BN), idesc_o = idesc_bf16(FA_BM, dhp);
            mbar_wait(bar_q, 0);
            auto issue_qk = [&](int j) {
                const int s = j & 1, b = j & 1;
                mbar_wait(bar_k_full(s), (uint32_t)((j >> 1) & 1));
                mbar_wait(bar_s_empty(b), (uint32_t)(((j >> 1) & 1) ^ 1));
                fence_after();
                for (int a = 0; a < n_at; ++a) {
                    const uint64_t qd = umma_desc_sw128(base + a * 16384), kd = umma_desc_sw128(k_st(s) + a * 16384);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + FA_S_COL + 128u * b, qd + 2u * k, kd + 2u * k, idesc_s, (a | k) != 0);
                }
                umma_commit(bar_k_empty(s));
                umma_commit(bar_s_full(b));
            };
            issue_qk(0);
            for (int j = 0; j < n_kt; ++j) {
                if (j + 1 < n_kt) issue_qk(j + 1);
                const int s = j & 1;
                mbar_wait(bar_v_full(s), (uint32_t)((j >> 1) & 1));
                mbar_wait(bar_p_full, (uint32_t)(j & 1));
                fence_after();
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {          // 128 keys = 8 x 16; P: 8 TMEM columns per step; V^T: two 64-key atoms
                    const uint64_t vd = umma_desc_sw128(v_st(s) + (ks >> 2) * (dhp * 128)) + 2u * (ks & 3);
                    umma_bf16_ts(tmem_base + FA_O_COL, tmem_base + FA_P_COL + 8u * ks, vd, idesc_o, (j | ks) != 0);
                }
                umma_commit(bar_v_empty(s));
                umma_commit(bar_o_ready);
            }
        }
    } else {
        // ---- softmax / correction / epilogue: thread = query row ----
        const int qd = warp & 3;
        const uint32_t lane_base = (uint32_t)(32 * qd) << 16;
        const float c = g.scale * 1.44269504088896340736f;      // scores in the exp2 domain
        float m = -INFINITY, l = 0.f;
        for (int j = 0; j < n_kt; ++j) {
            const int b = j & 1;
            mbar_wait(bar_s_full(b), (uint32_t)((j >> 1) & 1));
            fence_after();
            uint32_t v[4][32];
#pragma unroll
            for (int h4 = 0; h4 < 4; ++h4) tmem_ld32(tmem_base + lane_base + FA_S_COL + 128u * b + 32u * h4, v[h4]);
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_s_empty(b));           // S[buf] is in registers: the next QK^T may overwrite it
            const int kbase = j * FA_BN;
            float mx = m;
#pragma unroll
            for (int h4 = 0; h4 < 4; ++h4)
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float sv = __uint_as_float(v[h4][i]) * c;
                    if (kbase + 32 * h4 + i >= g.Lk) sv = -INFINITY;
                    v[h4][i] = __float_as_uint(sv);
                    mx = fmaxf(mx, sv);
                }
            const float alpha = exp2f(m - mx);                     // m = -inf on the first tile: alpha = 0, l = 0
            float sum = 0.f;
            uint32_t pk[64];
#pragma unroll
            for (int h4 = 0; h4 < 4; ++h4)
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float p0 = exp2f(__uint_as_float(v[h4][i]) - mx), p1 = exp2f(__uint_as_float(v[h4][i + 1]) - mx);
                    sum += p0 + p1;
                    const __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
                    pk[16 * h4 + i / 2] = *reinterpret_cast<const uint32_t*>(&pb);
                }
            l = l * alpha + sum;
            const bool grew = mx > m;
            m = mx;
            if (j > 0) {
                mbar_wait(bar_o_ready, (uint32_t)((j - 1) & 1));   // P V of the previous tile is complete: O and P may be touched
                fence_after();
                if (__any_sync(0xffffffffu, grew)) {
                    for (int c0 = 0; c0 < dhp; c0 += 32) {
                        uint32_t o[32];
                        tmem_ld32(tmem_base + lane_base + FA_O_COL + (uint32_t)c0, o);
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(tmem_base + lane_base + FA_O_COL + (uint32_t)c0, o);
                    }
                }
            }
            {
                uint32_t t0[32], t1[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) { t0[i] = pk[i]; t1[i] = pk[32 + i]; }
                tmem_st32(tmem_base + lane_base + FA_P_COL, t0);
                tmem_st32(tmem_base + lane_base + FA_P_COL + 32u, t1);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p_full);
        }
        // ---- epilogue ----
        mbar_wait(bar_o_ready, (uint32_t)((n_kt - 1) & 1));
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
    if (warp == 1) { fence_after(); tmem_dealloc(tmem_base, 512); }
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

bool attn_fused_supported(int dhp) { return dhp == 64 || dhp == 128; }

int attn_desc_make(AttnDesc* g, const void* q, const void* k, const void* vt, void* out, int NB, int heads, int dhp, long L, int Lk, int Lkp,
                   float scale) {
    memset(g, 0, sizeof(*g));
    const long HD = (long)heads * dhp;
    g->NB = NB; g->heads = heads; g->dhp = dhp; g->L = (int)L; g->Lk = Lk; g->scale = scale;
    g->out = (__nv_bfloat16*)out; g->ldo = HD;
    {
        const long dims[4] = {dhp, L, heads, NB}, str[4] = {1, HD, dhp, L * HD};
        const int box[4] = {64, FA_BM, 1, 1};
        if (fa_encode(&g->tmQ, q, dims, str, box)) return -1;
    }
    {
        const long dims[4] = {dhp, Lk, heads, NB}, str[4] = {1, HD, dhp, (long)Lk * HD};
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
    static int configured = 0;
    if (configured < smem) {
        cudaError_t e = cudaFuncSetAttribute(unet_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    dim3 grid((g.L + FA_BM - 1) / FA_BM, g.heads, g.NB);
    unet_attn_kernel<<<grid, FA_THREADS, smem, st>>>(g);
    return (int)cudaGetLastError();
}

}  // namespace uce
