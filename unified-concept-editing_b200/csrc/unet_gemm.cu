// bf16 tcgen05 GEMM of the SD U-Net denoise step: every Linear, 1x1 conv, 3x3 conv
// (implicit GEMM) and the two attention contractions run through this one kernel.
//
//     D[b][m][n] = alpha * sum_k A[b][m][k] * B[b][n][k]  (+ bias[n]) (+ rowbias[img(m)][n]) (+ residual[b][m][n])
//
// * operands are K-major bf16, fetched by TMA into 128B-swizzled smem stages (3-stage mbarrier pipeline);
//   accumulation in TMEM (fp32) by tcgen05.mma.kind::f16, M = 128, N = 128 per CTA, K step 64.
// * conv3x3 is an implicit GEMM without im2col: activations are NHWC, the A tile of an output rectangle
//   (tn x th x tw = 128 pixels) for tap (ky,kx) is ONE 4-D TMA box shifted by the tap offset; out-of-image
//   coordinates are zero-filled by the TMA unit (= the conv padding), stride 2 uses the tensor map's element
//   strides.  Weights are stored tap-major [Cout][ky][kx][Cin].
// * warp roles: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2-5 epilogue (TMEM -> registers ->
//   bias / time-embedding / residual -> bf16 or fp32 global stores).
#include "tc_common.cuh"
#include "unet_gemm.h"
#include <cuda_bf16.h>

namespace uce {
using namespace tc;

constexpr int UG_BM = 128, UG_BN = 128, UG_BK = 64, UG_MAX_STAGES = 6;
constexpr int UG_THREADS = 192;
constexpr int UG_ATOM_BYTES = (UG_BM + UG_BN) * UG_BK * 2;      // 32 KB: one 64-wide k atom of A and of B
inline int ug_smem_bytes(int stages, int katoms) { return stages * katoms * UG_ATOM_BYTES + 1024 + 256; }

#ifdef UG_TRACE      // scripts/gemm_probe.cu: clock64 stamps of CTA (0,0,0)
__device__ long long ug_trace[16];
#define UG_STAMP(i) do { if ((blockIdx.x | blockIdx.y | blockIdx.z) == 0) ug_trace[i] = clock64(); } while (0)
#else
#define UG_STAMP(i) do { } while (0)
#endif


// ---- epilogue shared by both GEMM kernels: TMEM lane = tile row; `bn` accumulator columns starting at global column n_base.
// The bias / time-embedding / residual operands of a 32-column chunk are fetched (vector loads) BEFORE the accumulator is
// needed — chunk 0 while the main loop is still running, chunk c+1 right after chunk c's values are consumed — so their
// latency is off the critical path (the first version paid ~3000 cycles per chunk on dependent scalar loads).
struct EpiOperands { float4 b[8]; uint4 r[4]; };
__device__ __forceinline__ void epi_fetch(const GemmDesc& g, EpiOperands& o, bool active, int n0, int img, long rbase) {
    const bool full = active && (n0 + 32 <= g.N);
#pragma unroll
    for (int j = 0; j < 8; ++j) o.b[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) o.r[j] = make_uint4(0u, 0u, 0u, 0u);
    if (!active || g.ksplit > 1) return;
    if (g.bias) {
        if (full) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o.b[j] = __ldg(reinterpret_cast<const float4*>(g.bias + n0) + j);
        } else {
            float* bf = reinterpret_cast<float*>(o.b);
#pragma unroll
            for (int i = 0; i < 32; ++i) if (n0 + i < g.N) bf[i] = g.bias[n0 + i];
        }
    }
    if (g.rowbias) {
        const float* rb = g.rowbias + (long)img * g.N + n0;
        float* bf = reinterpret_cast<float*>(o.b);
        if (full && ((g.N & 3) == 0)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float4 t = __ldg(reinterpret_cast<const float4*>(rb) + j); bf[4 * j] += t.x; bf[4 * j + 1] += t.y; bf[4 * j + 2] += t.z; bf[4 * j + 3] += t.w; }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (n0 + i < g.N) bf[i] += rb[i];
        }
    }
    if (g.residual) {
        const __nv_bfloat16* rp = g.residual + rbase + n0;
        if (full && ((g.ldr & 7) == 0)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) o.r[j] = *reinterpret_cast<const uint4*>(rp + 8 * j);
        } else {
            __nv_bfloat16* rh = reinterpret_cast<__nv_bfloat16*>(o.r);
#pragma unroll
            for (int i = 0; i < 32; ++i) if (n0 + i < g.N) rh[i] = rp[i];
        }
    }
}

__device__ __forceinline__ void gemm_epilogue(const GemmDesc& g, uint32_t tmem_base, uint32_t bar_acc, int q, long row, bool row_ok, int img,
                                              int n_base, int bn, int zk, int b1, int b2) {
    const long obase = (long)b1 * g.out_b1_stride + (long)b2 * g.out_b2_stride + row * g.ldo;
    const long rbase = (long)b1 * g.res_b1_stride + (long)b2 * g.res_b2_stride + row * g.ldr;
    EpiOperands eo;
    epi_fetch(g, eo, row_ok && n_base < g.N, n_base, img, rbase);
    mbar_wait(bar_acc, 0);
    fence_after();
    if (threadIdx.x == 64) UG_STAMP(5);
#pragma unroll 1
    for (int c = 0; c < bn; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)c, v);
        const int n0 = n_base + c;
        const bool active = row_ok && n0 < g.N;
        float f[32];
        {
            const float* bf = reinterpret_cast<const float*>(eo.b);
            const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(eo.r);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const float2 t2 = __bfloat1622float2(rh[i / 2]);
                f[i] = fmaf(g.alpha, __uint_as_float(v[i]), bf[i] + t2.x);
                f[i + 1] = fmaf(g.alpha, __uint_as_float(v[i + 1]), bf[i + 1] + t2.y);
            }
        }
        if (c + 32 < bn) epi_fetch(g, eo, row_ok && n0 + 32 < g.N, n0 + 32, img, rbase);      // operands of the next chunk
        if (!active) continue;
        const bool full = (n0 + 32 <= g.N);
        if (g.ksplit > 1) {            // partial sum of this k range -> its own slab; the finalize pass sums the slabs in a
                                       // fixed order (deterministic, unlike atomics) and applies the epilogue terms
            float* wp = g.splitk_ws + ((long)zk * g.M + row) * (long)g.N + n0;
            if (full) {                // N % 4 == 0 is a precondition of splitting
#pragma unroll
                for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(wp + 4 * j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) if (n0 + i < g.N) wp[i] = f[i];
            }
            continue;
        }
        if (g.out_fp32) {
            float* op = reinterpret_cast<float*>(g.out) + obase + n0;
            if (full && ((g.ldo & 3) == 0)) {
#pragma unroll
                for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(op + 4 * j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) if (n0 + i < g.N) op[i] = f[i];
            }
        } else {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(g.out) + obase + n0;
            if (full && ((g.ldo & 7) == 0)) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 u;
                    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                    for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(f[8 * j + 2 * e], f[8 * j + 2 * e + 1]);
                    *reinterpret_cast<uint4*>(op + 8 * j) = u;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) if (n0 + i < g.N) op[i] = __float2bfloat16(f[i]);
            }
        }
    }
}

__global__ void __launch_bounds__(UG_THREADS, 2) unet_gemm_kernel(const __grid_constant__ GemmDesc g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    pdl_launch();
    if (threadIdx.x == 0) UG_STAMP(0);
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int NS = g.stages, KA = g.katoms;
    const int stage_bytes = KA * UG_ATOM_BYTES;
    const uint32_t bars = base + NS * stage_bytes;
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_empty = [&](int s) { return bars + 8u * (UG_MAX_STAGES + s); };
    const uint32_t bar_acc = bars + 8u * (2 * UG_MAX_STAGES);
    const uint32_t tmem_slot = bars + 8u * (2 * UG_MAX_STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int n_tile = blockIdx.x, m_tile = blockIdx.y;
    const int zb = (g.ksplit > 1) ? 0 : blockIdx.z, zk = (g.ksplit > 1) ? blockIdx.z : 0;
    const int b1 = zb % g.b1cnt, b2 = zb / g.b1cnt;
    const int kpt = g.cin / UG_BK + ((g.cin % UG_BK) ? 1 : 0);     // k-iterations per tap
    const int n_k_total = g.taps * kpt;
    const int k_begin = (g.ksplit > 1) ? (int)((long)n_k_total * zk / g.ksplit) : 0;
    const int k_end = (g.ksplit > 1) ? (int)((long)n_k_total * (zk + 1) / g.ksplit) : n_k_total;
    const int n_atoms = k_end - k_begin;                   // 64-wide k atoms of this CTA
    const int n_k = (n_atoms + KA - 1) / KA;               // pipeline iterations

    // conv: decode the output rectangle of this m-tile
    int img0 = 0, h0 = 0, w0 = 0;
    if (g.conv) {
        const int tiles_w = g.Wo / g.tw, tiles_h = g.Ho / g.th;
        int t = m_tile;
        w0 = (t % tiles_w) * g.tw; t /= tiles_w;
        h0 = (t % tiles_h) * g.th; t /= tiles_h;
        img0 = t * g.tn;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        mbar_init(bar_acc, 1);
        mbar_fence_init();
        tma_prefetch_desc(&g.tmA); tma_prefetch_desc(&g.tmB);
    }
    if (warp == 1) tmem_alloc(tmem_slot, UG_BN);
    fence_before();
    __syncthreads();
    fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (threadIdx.x == 0) UG_STAMP(1);

    if (warp == 0) {
        // all lanes run the loop and the barrier waits; one elected lane issues (see elect_one() in tc_common.cuh)
        // which = 1: B (weights: not produced by an earlier kernel of the schedule unless batched), 2: A, 3: both
        auto load_stage = [&](int it, int which) {
            const int s = it % NS;
            const int na = min(KA, n_atoms - it * KA);
            if (which & 1) mbar_arrive_expect_tx(bar_full(s), (uint32_t)(na * (g.a_bytes + UG_BN * UG_BK * 2)));
            for (int a = 0; a < na; ++a) {
                const uint32_t a_dst = base + s * stage_bytes + a * (UG_BM * UG_BK * 2);
                const uint32_t b_dst = base + s * stage_bytes + KA * (UG_BM * UG_BK * 2) + a * (UG_BN * UG_BK * 2);
                const int kit = k_begin + it * KA + a;
                const int tap = kit / kpt, c0 = (kit % kpt) * UG_BK;
                if (which & 2) {
                    if (g.conv) {
                        const int ky = tap / 3, kx = tap % 3;
                        tma_load_4d(a_dst, &g.tmA, bar_full(s), c0, w0 * g.stride + kx - g.pad, h0 * g.stride + ky - g.pad, img0);
                    } else {
                        tma_load_4d(a_dst, &g.tmA, bar_full(s), c0, m_tile * UG_BM, g.a_batched ? b1 : 0, g.a_batched ? b2 : 0);
                    }
                }
                if (which & 1) tma_load_4d(b_dst, &g.tmB, bar_full(s), tap * g.cin + c0, n_tile * UG_BN, g.b_batched ? b1 : 0, g.b_batched ? b2 : 0);
            }
        };
        // weight tiles of the first stages are requested before the dependency wait: they overlap the previous kernel's tail
        const int pre = min(NS, n_k);
        const bool early_b = !g.b_batched;
        if (early_b && elect_one()) for (int it = 0; it < pre; ++it) load_stage(it, 1);
        __syncwarp();
        pdl_wait();
        if (elect_one()) for (int it = 0; it < pre; ++it) load_stage(it, early_b ? 2 : 3);
        __syncwarp();
        for (int it = pre; it < n_k; ++it) {
            if (it == pre && lane == 0) UG_STAMP(2);
            mbar_wait(bar_empty(it % NS), (uint32_t)(((it / NS) & 1) ^ 1));
            __syncwarp();
            if (elect_one()) load_stage(it, 3);
        }
    } else if (warp == 1) {
        // N of this tile's MMAs = the valid columns rounded up to 16 (a 128 x 256 tile variant was measured and was not
        // faster end to end: fewer CTAs on the many under-filled grids of the U-Net offset the cheaper issue)
        int n_valid = g.N - n_tile * UG_BN;
        n_valid = n_valid >= UG_BN ? UG_BN : ((n_valid + 15) & ~15);
        const uint32_t idesc = idesc_bf16(UG_BM, n_valid);
        for (int it = 0; it < n_k; ++it) {
            const int s = it % NS;
            mbar_wait(bar_full(s), (uint32_t)((it / NS) & 1));
            fence_after();
            __syncwarp();
            if (elect_one()) {
                if (it == 0) UG_STAMP(3);
                const int na = min(KA, n_atoms - it * KA);
                for (int a = 0; a < na; ++a) {
                    const uint64_t a_desc = umma_desc_sw128(base + s * stage_bytes + a * (UG_BM * UG_BK * 2));
                    const uint64_t b_desc = umma_desc_sw128(base + s * stage_bytes + KA * (UG_BM * UG_BK * 2) + a * (UG_BN * UG_BK * 2));
#pragma unroll
                    for (int k = 0; k < UG_BK / 16; ++k)
                        umma_bf16(tmem_base, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (it | a | k) != 0);
                }
                umma_commit(bar_empty(s));
                if (it == n_k - 1) { umma_commit(bar_acc); UG_STAMP(4); }
            }
        }
    } else {
        // ---- epilogue: TMEM lane = tile row ----
        const int q = warp & 3;                       // TMEM lane quarter this warp may touch
        const int r = 32 * q + lane;                  // tile row
        long row;                                     // output row index within the batch (pixel index for conv)
        int img = 0;
        bool row_ok;
        if (g.conv) {
            const int per_img = g.th * g.tw;
            const int ti = r / per_img, rem = r % per_img;
            img = img0 + ti;
            const int hh = h0 + rem / g.tw, ww = w0 + rem % g.tw;
            row = ((long)img * g.Ho + hh) * g.Wo + ww;
            row_ok = r < g.tile_rows;
        } else {
            row = (long)m_tile * UG_BM + r;
            row_ok = row < g.M;
            if (g.rows_per_img > 0) img = (int)(row / g.rows_per_img);
        }
        pdl_wait();                                   // residual / time-embedding rows come from earlier kernels
        gemm_epilogue(g, tmem_base, bar_acc, q, row, row_ok, img, n_tile * UG_BN, UG_BN, zk, b1, b2);
    }
    if (threadIdx.x == 64) UG_STAMP(6);
    fence_before();
    __syncthreads();
    if (warp == 1) { fence_after(); tmem_dealloc(tmem_base, UG_BN); }
    if (threadIdx.x == 0) UG_STAMP(7);
}



// ---- TMA epilogue (CTA-pair kernel): the row-per-lane register epilogue above issues 16-byte accesses to 32 different
// lines per warp instruction (bias / residual / output), which made the epilogue as long as a 20-iteration main loop.  Here the
// residual tile is fetched by TMA into 128B-swizzled shared memory while the main loop runs, the accumulator is combined with
// it IN PLACE (lane = row: the swizzle makes the 16-byte accesses of a quarter-warp conflict-free), and each [128 rows x 64
// columns] box leaves through one TMA store (which also clips rows >= M and columns >= N).  Split-K partial tiles take the
// same route as fp32 [128 x 32] boxes into their slab.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }

// c1..c3: tile coordinates of the row dimension(s): linear (m0, 0, 0), conv (w0, h0, img0)
__device__ __forceinline__ void gemm_epilogue_tma(const GemmDesc& g, uint32_t tmem_base, uint32_t bar_acc, uint32_t bar_res, uint32_t stage,
                                                  uint32_t res_base, uint32_t sbias, int q, int lane, int img, bool row_ok, int n_base, int bn,
                                                  int c1, int c2, int c3, int zk) {
    const int r = 32 * q + lane, et = threadIdx.x - 64;
    const uint32_t lane_base = (uint32_t)(32 * q) << 16;
    const int n_cols = min(bn, g.N - n_base);
    for (int i = et; i < bn; i += 128) {
        const float b = (g.bias && g.ksplit <= 1 && n_base + i < g.N) ? g.bias[n_base + i] : 0.f;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbias + 4u * i), "f"(b) : "memory");
    }
    epi_bar_sync();
    if (g.tma_epi == 1 && g.residual) mbar_wait(bar_res, 0);
    mbar_wait(bar_acc, 0);
    fence_after();
    if (threadIdx.x == 64) UG_STAMP(5);
    const uint32_t row_off = (uint32_t)r * 128u, sw = (uint32_t)(r & 7);
    if (g.tma_epi == 1) {
        const uint32_t obuf = g.residual ? res_base : stage;           // combine in place when there is a residual tile
        const int n_boxes = (n_cols + 63) / 64;
        for (int b = 0; b < n_boxes; ++b) {
            const uint32_t box = obuf + (uint32_t)b * 16384u;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_base + (uint32_t)(b * 64 + hf * 32), v);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = hf * 4 + jj, col = b * 64 + j * 8;
                    const uint32_t off = row_off + (((uint32_t)j ^ sw) << 4);
                    float add[8];
                    {
                        const uint4 b0 = lds128(sbias + 4u * col), b1 = lds128(sbias + 4u * col + 16u);
                        add[0] = __uint_as_float(b0.x); add[1] = __uint_as_float(b0.y); add[2] = __uint_as_float(b0.z); add[3] = __uint_as_float(b0.w);
                        add[4] = __uint_as_float(b1.x); add[5] = __uint_as_float(b1.y); add[6] = __uint_as_float(b1.z); add[7] = __uint_as_float(b1.w);
                    }
                    if (g.rowbias && row_ok && n_base + col < g.N) {
                        const float4* rb = reinterpret_cast<const float4*>(g.rowbias + (long)img * g.N + n_base + col);
                        const float4 t0 = __ldg(rb), t1 = __ldg(rb + 1);
                        add[0] += t0.x; add[1] += t0.y; add[2] += t0.z; add[3] += t0.w; add[4] += t1.x; add[5] += t1.y; add[6] += t1.z; add[7] += t1.w;
                    }
                    if (g.residual) {
                        const uint4 rr = lds128(box + off);
                        const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
                        for (int e = 0; e < 4; ++e) { const float2 t2 = __bfloat1622float2(rh[e]); add[2 * e] += t2.x; add[2 * e + 1] += t2.y; }
                    }
                    uint4 o;
                    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        oh[e] = __floats2bfloat162_rn(fmaf(g.alpha, __uint_as_float(v[jj * 8 + 2 * e]), add[2 * e]),
                                                      fmaf(g.alpha, __uint_as_float(v[jj * 8 + 2 * e + 1]), add[2 * e + 1]));
                    sts128(box + off, o);
                }
            }
            fence_proxy_async();
            epi_bar_sync();
            if (threadIdx.x == 64) { tma_store_4d(&g.tmO, box, n_base + b * 64, c1, c2, c3); tma_store_commit(); }
        }
    } else {                       // split-K partial sums: fp32 boxes of 32 columns into slab zk
        const int n_boxes = (n_cols + 31) / 32;
        for (int b = 0; b < n_boxes; ++b) {
            const uint32_t box = stage + (uint32_t)b * 16384u;
            uint32_t v[32];
            tmem_ld32(tmem_base + lane_base + (uint32_t)(b * 32), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint4 o;
                o.x = __float_as_uint(g.alpha * __uint_as_float(v[4 * j])); o.y = __float_as_uint(g.alpha * __uint_as_float(v[4 * j + 1]));
                o.z = __float_as_uint(g.alpha * __uint_as_float(v[4 * j + 2])); o.w = __float_as_uint(g.alpha * __uint_as_float(v[4 * j + 3]));
                sts128(box + row_off + (((uint32_t)j ^ sw) << 4), o);
            }
            fence_proxy_async();
            epi_bar_sync();
            if (threadIdx.x == 64) {
                if (g.conv) tma_store_5d(&g.tmP, box, n_base + b * 32, c1, c2, c3, zk);
                else tma_store_4d(&g.tmP, box, n_base + b * 32, c1, zk, 0);
                tma_store_commit();
            }
        }
    }
    if (threadIdx.x == 64) tma_store_wait_read();        // shared memory must stay alive until the bulk stores have read it
}

// ------------------------------------------------------------------------------------------ CTA-pair kernel
// Same GEMM on a 256 x bn tile per CLUSTER of two CTAs (tcgen05.mma.cta_group::2, M = 256): each CTA of the pair owns 128
// rows (its own A tile and TMEM accumulator) and stages only HALF of the B tile; the tensor cores of both SMs read both
// halves.  An SM ingests ~40 B/clk through TMA (scripts/gemm_probe.cu), a 128 x 128 x 64 step of the single-CTA kernel costs
// it 32 KB for 64 MMA-cycles x 4: the pair halves the B bytes per flop and allows bn up to 256, i.e. up to 2x the
// flops per ingested byte.  Protocol per stage: both CTAs' TMA loads signal the LEADER's full barrier (cta_group::2 TMA,
// peer bit cleared in the barrier address); the leader issues the MMAs and commits, multicast, to the empty barrier of both.
constexpr uint32_t UG_PEER_MASK = 0xFEFFFFFFu;            // shared::cluster address bit 24 = CTA rank within the pair
constexpr int UG2_STAGE_BYTES = 32768;                     // A 16 KB + B half (<= 128 rows) 16 KB

__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {     // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

__global__ void __launch_bounds__(UG_THREADS, 2) unet_gemm_pair_kernel(const __grid_constant__ GemmDesc g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    pdl_launch();
    if (threadIdx.x == 0) UG_STAMP(0);
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int NS = g.stages;
    const uint32_t bars = base + NS * UG2_STAGE_BYTES;
    auto bar_full = [&](int s) { return bars + 8u * s; };
    auto bar_empty = [&](int s) { return bars + 8u * (UG_MAX_STAGES + s); };
    const uint32_t bar_acc = bars + 8u * (2 * UG_MAX_STAGES);
    const uint32_t tmem_slot = bars + 8u * (2 * UG_MAX_STAGES + 1);
    const uint32_t bar_res = bars + 8u * (2 * UG_MAX_STAGES + 2);
    const uint32_t sbias = bars + 256u;                        // bn floats
    const uint32_t res_base = bars + 2048u;                    // residual / output staging boxes (1024-aligned: NS * 32 KB + 2 KB)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();

    const int bn = g.bn;
    const int m_tile = blockIdx.x, n_tile = blockIdx.y, zk = blockIdx.z;     // the pair is two consecutive m-tiles: clusters of (2,1,1)
    const int n_base = n_tile * bn;
    int n_mma = g.N - n_base;                                  // columns of this tile's MMAs: valid columns rounded up to 32
    n_mma = n_mma >= bn ? bn : ((n_mma + 31) & ~31);
    const int n_half = n_mma / 2;
    const int kpt = g.cin / UG_BK + ((g.cin % UG_BK) ? 1 : 0);
    const int n_k_total = g.taps * kpt;
    const int k_begin = (g.ksplit > 1) ? (int)((long)n_k_total * zk / g.ksplit) : 0;
    const int k_end = (g.ksplit > 1) ? (int)((long)n_k_total * (zk + 1) / g.ksplit) : n_k_total;
    const int n_k = k_end - k_begin;

    int img0 = 0, h0 = 0, w0 = 0;
    if (g.conv) {
        const int tiles_w = g.Wo / g.tw, tiles_h = g.Ho / g.th;
        int t = m_tile;
        w0 = (t % tiles_w) * g.tw; t /= tiles_w;
        h0 = (t % tiles_h) * g.th; t /= tiles_h;
        img0 = t * g.tn;                                       // a padding m-tile (odd tile count) lands beyond the last image: zero fill
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
        mbar_init(bar_acc, 1); mbar_init(bar_res, 1);
        mbar_fence_init();
        tma_prefetch_desc(&g.tmA); tma_prefetch_desc(&g.tmB);
        if (g.tma_epi == 1) { tma_prefetch_desc(&g.tmO); if (g.residual) tma_prefetch_desc(&g.tmR); }
        if (g.tma_epi == 2) tma_prefetch_desc(&g.tmP);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)g.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    fence_before();
    cluster_sync_all();                                        // the peer's barriers exist before anything signals them
    __syncthreads();                                           // (the cluster barrier already orders the TMEM-address slot; this CTA barrier is what
                                                               //  compute-sanitizer racecheck recognises: 2 209 false RAW reports on the slot without it)
    fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (threadIdx.x == 0) UG_STAMP(1);

    const int b_bytes = (bn / 2) * UG_BK * 2;
    const int ec1 = g.conv ? w0 : m_tile * UG_BM, ec2 = g.conv ? h0 : 0, ec3 = g.conv ? img0 : 0;     // row coordinates of this tile
    if (warp == 0) {
        // all lanes run the loop and the barrier waits; one elected lane issues (see elect_one() in tc_common.cuh)
        // which = 1: this CTA's half of the B (weight) tile, 2: its A tile, 3: both
        auto load_stage = [&](int it, int which) {
            const int s = it % NS;
            // the leader's barrier counts the bytes of both CTAs (a peer load may land before this expect_tx: the
            // transaction count is signed, and the phase cannot complete before the leader's own arrival)
            if ((which & 1) && rank == 0) mbar_arrive_expect_tx(bar_full(s), (uint32_t)(2 * (g.a_bytes + b_bytes)));
            const uint32_t lead_bar = bar_full(s) & UG_PEER_MASK;
            const uint32_t a_dst = base + s * UG2_STAGE_BYTES, b_dst = a_dst + UG_BM * UG_BK * 2;
            const int kit = k_begin + it;
            const int tap = kit / kpt, c0 = (kit % kpt) * UG_BK;
            if (which & 2) {
                if (g.conv) {
                    const int ky = tap / 3, kx = tap % 3;
                    tma_load_4d_pair(a_dst, &g.tmA, lead_bar, c0, w0 * g.stride + kx - g.pad, h0 * g.stride + ky - g.pad, img0);
                } else {
                    tma_load_4d_pair(a_dst, &g.tmA, lead_bar, c0, m_tile * UG_BM, 0, 0);
                }
            }
            if (which & 1) tma_load_4d_pair(b_dst, &g.tmB, lead_bar, tap * g.cin + c0, n_base + (int)rank * n_half, 0, 0);
        };
        // weight tiles of the first stages are requested before the dependency wait: they overlap the previous kernel's tail
        const int pre = min(NS, n_k);
        if (elect_one()) for (int it = 0; it < pre; ++it) load_stage(it, 1);
        __syncwarp();
        pdl_wait();
        if (elect_one()) {
            if (g.tma_epi == 1 && g.residual) {               // residual tile -> its own shared-memory boxes, ahead of the activations
                const int n_boxes = (min(bn, g.N - n_base) + 63) / 64;
                mbar_arrive_expect_tx(bar_res, (uint32_t)(n_boxes * 16384));
                for (int b = 0; b < n_boxes; ++b) tma_load_4d(res_base + (uint32_t)b * 16384u, &g.tmR, bar_res, n_base + b * 64, ec1, ec2, ec3);
            }
            for (int it = 0; it < pre; ++it) load_stage(it, 2);
        }
        __syncwarp();
        for (int it = pre; it < n_k; ++it) {
            if (it == pre && lane == 0) UG_STAMP(2);
            mbar_wait(bar_empty(it % NS), (uint32_t)(((it / NS) & 1) ^ 1));
            __syncwarp();
            if (elect_one()) load_stage(it, 3);
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // the whole warp of the leader CTA runs the loop and the barrier waits; ONE elected lane issues the MMAs and commits
            const uint32_t idesc = idesc_bf16(2 * UG_BM, n_mma);
            for (int it = 0; it < n_k; ++it) {
                const int s = it % NS;
                mbar_wait(bar_full(s), (uint32_t)((it / NS) & 1));
                fence_after();
                __syncwarp();
                if (elect_one()) {
                    if (it == 0) UG_STAMP(3);
                    const uint64_t a_desc = umma_desc_sw128(base + s * UG2_STAGE_BYTES);
                    const uint64_t b_desc = umma_desc_sw128(base + s * UG2_STAGE_BYTES + UG_BM * UG_BK * 2);
#pragma unroll
                    for (int k = 0; k < UG_BK / 16; ++k)
                        umma_bf16_pair(tmem_base, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (it | k) != 0);
                    umma_commit_pair(bar_empty(s));
                    if (it == n_k - 1) { umma_commit_pair(bar_acc); UG_STAMP(4); }
                }
            }
        }
    } else {
        const int q = warp & 3;
        const int r = 32 * q + lane;
        long row; int img = 0; bool row_ok;
        if (g.conv) {
            const int per_img = g.th * g.tw;
            const int ti = r / per_img, rem = r % per_img;
            img = img0 + ti;
            const int hh = h0 + rem / g.tw, ww = w0 + rem % g.tw;
            row = ((long)img * g.Ho + hh) * g.Wo + ww;
            row_ok = m_tile < g.m_tiles;
        } else {
            row = (long)m_tile * UG_BM + r;
            row_ok = row < g.M;
            if (g.rows_per_img > 0) img = (int)(row / g.rows_per_img);
        }
        pdl_wait();                                   // bias is a weight, but rowbias / residual come from earlier kernels
        if (g.tma_epi) gemm_epilogue_tma(g, tmem_base, bar_acc, bar_res, base, res_base, sbias, q, lane, img, row_ok, n_base, bn, ec1, ec2, ec3, zk);
        else gemm_epilogue(g, tmem_base, bar_acc, q, row, row_ok, img, n_base, bn, zk, 0, 0);
    }
    if (threadIdx.x == 64) UG_STAMP(6);
    fence_before();
    cluster_sync_all();                                        // neither CTA may free TMEM / exit while the pair's MMAs or barriers are live
    if (warp == 1) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)g.tmem_cols) : "memory");
    }
    if (threadIdx.x == 0) UG_STAMP(7);
}

// ------------------------------------------------------------------------------------------ host
static int encode_map(CUtensorMap* m, bool f32, const void* ptr, int rank, const long* dims, const long* strides_elems,
                      const int* box, const int* estr) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return -1;
    cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; es[i] = (cuuint32_t)(estr ? estr[i] : 1); }
    for (int i = 1; i < rank; ++i) gs[i - 1] = (cuuint64_t)strides_elems[i] * (f32 ? 4 : 2);
    CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gd, gs,
                     bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}
static int encode_bf16_map(CUtensorMap* m, const void* ptr, int rank, const long* dims, const long* strides_elems,
                           const int* box, const int* estr) {
    return encode_map(m, false, ptr, rank, dims, strides_elems, box, estr);
}

int gemm_desc_linear(GemmDesc* g, const void* A, long lda, long a_b1_stride, long a_b2_stride, const void* B, long ldb,
                     long b_b1_stride, long b_b2_stride, int M, int N, int Kd, int b1cnt, int b2cnt, int a_batched, int b_batched) {
    memset(g, 0, sizeof(*g));
    g->m_tiles = (M + UG_BM - 1) / UG_BM; g->tile_rows = UG_BM; g->a_bytes = UG_BM * UG_BK * 2;
    g->M = M; g->N = N; g->Kd = Kd; g->batch = b1cnt * b2cnt; g->b1cnt = b1cnt; g->conv = 0; g->taps = 1; g->cin = Kd; g->alpha = 1.f;
    g->a_batched = a_batched; g->b_batched = b_batched;
    g->b_ptr = B; g->b_ld = ldb;
    {
        long dims[4] = {Kd, M, a_batched ? b1cnt : 1, a_batched ? b2cnt : 1};
        long str[4] = {1, lda, a_batched && a_b1_stride ? a_b1_stride : (long)M * lda, a_batched && a_b2_stride ? a_b2_stride : (long)M * lda};
        int box[4] = {UG_BK, UG_BM, 1, 1};
        if (encode_bf16_map(&g->tmA, A, 4, dims, str, box, nullptr)) return -1;
    }
    {
        long dims[4] = {Kd, N, b_batched ? b1cnt : 1, b_batched ? b2cnt : 1};
        long str[4] = {1, ldb, b_batched && b_b1_stride ? b_b1_stride : (long)N * ldb, b_batched && b_b2_stride ? b_b2_stride : (long)N * ldb};
        int box[4] = {UG_BK, UG_BN, 1, 1};
        if (encode_bf16_map(&g->tmB, B, 4, dims, str, box, nullptr)) return -1;
    }
    return 0;
}

int gemm_desc_conv(GemmDesc* g, const void* act_nhwc, int NB, int Hin, int Win, int Cin, const void* w_tapmajor, int Cout,
                   int ksize, int stride) {
    memset(g, 0, sizeof(*g));
    const int pad = ksize / 2;
    const int Ho = (Hin + 2 * pad - ksize) / stride + 1, Wo = (Win + 2 * pad - ksize) / stride + 1;
    g->conv = 1; g->taps = ksize * ksize; g->cin = Cin; g->stride = stride; g->pad = pad; g->alpha = 1.f;
    g->Ho = Ho; g->Wo = Wo; g->NBimg = NB; g->batch = 1; g->b1cnt = 1; g->a_batched = 0; g->b_batched = 0;
    g->M = NB * Ho * Wo; g->N = Cout; g->Kd = g->taps * Cin;
    g->b_ptr = w_tapmajor; g->b_ld = (long)g->taps * Cin;
    // output rectangle of 128 pixels
    int tw = Wo < 128 ? Wo : 128, th = 128 / tw; if (th > Ho) th = Ho;
    int tn = 128 / (tw * th); if (tn > NB) tn = NB;
    if (tw * th * tn > 128 || Wo % tw || Ho % th || NB % tn) return -2;
    g->tw = tw; g->th = th; g->tn = tn;
    g->tile_rows = tw * th * tn;                       // < 128 on tiny spatial levels: the remaining accumulator rows are ignored
    g->a_bytes = g->tile_rows * UG_BK * 2;
    g->m_tiles = (NB / tn) * (Ho / th) * (Wo / tw);
    {
        long dims[4] = {Cin, Win, Hin, NB}, str[4] = {1, Cin, (long)Win * Cin, (long)Hin * Win * Cin};
        int box[4] = {UG_BK, (tw - 1) * stride + 1, (th - 1) * stride + 1, tn};
        int es[4] = {1, stride, stride, 1};
        if (encode_bf16_map(&g->tmA, act_nhwc, 4, dims, str, box, es)) return -1;
    }
    {
        long dims[4] = {(long)g->taps * Cin, Cout, 1, 1};
        long str[4] = {1, (long)g->taps * Cin, (long)Cout * g->taps * Cin, (long)Cout * g->taps * Cin};
        int box[4] = {UG_BK, UG_BN, 1, 1};
        if (encode_bf16_map(&g->tmB, w_tapmajor, 4, dims, str, box, nullptr)) return -1;
    }
    return 0;
}

// out = sum_z ws[z] (+bias) (+rowbias[img]) (+residual).  One thread per 4 consecutive columns.
__global__ void __launch_bounds__(256) splitk_finalize_kernel(GemmDesc g, int rows_per_img) {
    pdl_launch(); pdl_wait();
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nv = g.N / 4;
    if (idx >= (long)g.M * nv) return;
    const long row = idx / nv; const int n0 = (int)(idx % nv) * 4;
    const float* wp = g.splitk_ws + row * g.N + n0;
    const long slab = (long)g.M * g.N;
    float4 v = *reinterpret_cast<const float4*>(wp);
#pragma unroll 4
    for (int z = 1; z < g.ksplit; ++z) {
        const float4 t = *reinterpret_cast<const float4*>(wp + z * slab);
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    float f[4] = {v.x, v.y, v.z, v.w};
    if (g.bias) { const float4 b = *reinterpret_cast<const float4*>(g.bias + n0); f[0] += b.x; f[1] += b.y; f[2] += b.z; f[3] += b.w; }
    // 16-byte / 8-byte accesses where the leading dimensions allow it (N % 4 == 0 always; ldr / ldo multiples of 4 in every U-Net use)
    if (g.rowbias) { const float4 rb = *reinterpret_cast<const float4*>(g.rowbias + (row / rows_per_img) * g.N + n0); f[0] += rb.x; f[1] += rb.y; f[2] += rb.z; f[3] += rb.w; }
    if (g.residual) {
        const __nv_bfloat16* rp = g.residual + row * g.ldr + n0;
        if ((g.ldr & 3) == 0) {
            const uint2 r2 = *reinterpret_cast<const uint2*>(rp);
            const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&r2);
            const float2 a = __bfloat1622float2(rh[0]), b = __bfloat1622float2(rh[1]);
            f[0] += a.x; f[1] += a.y; f[2] += b.x; f[3] += b.y;
        } else {
            for (int i = 0; i < 4; ++i) f[i] += __bfloat162float(rp[i]);
        }
    }
    if (g.out_fp32) {
        float* op = reinterpret_cast<float*>(g.out) + row * g.ldo + n0;
        if ((g.ldo & 3) == 0) *reinterpret_cast<float4*>(op) = make_float4(f[0], f[1], f[2], f[3]);
        else for (int i = 0; i < 4; ++i) op[i] = f[i];
    } else {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(g.out) + row * g.ldo + n0;
        if ((g.ldo & 3) == 0) {
            uint2 o2;
            __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o2);
            oh[0] = __floats2bfloat162_rn(f[0], f[1]); oh[1] = __floats2bfloat162_rn(f[2], f[3]);
            *reinterpret_cast<uint2*>(op) = o2;
        } else {
            for (int i = 0; i < 4; ++i) op[i] = __float2bfloat16(f[i]);
        }
    }
}

// Pipeline shape.  Measured on the SD-1.4 schedule (profiles/): neither a deeper pipeline (6 stages), nor two k atoms per
// barrier round trip, nor 128 x 256 tiles changed the end-to-end step time — the mid-size GEMMs of the U-Net (80-320 CTAs of
// 128 x 128 tiles streaming 32 KB per k-iteration) sit at ~8 TB/s of L2->SM operand traffic.  Raising the arithmetic
// intensity per L2 byte (2-CTA tcgen05.mma with multicast TMA) is the next step; until then: 3 single-atom stages with two
// CTAs per SM for full grids, 6 stages for under-filled ones.
static int gemm_ctas(const GemmDesc& g);
int gemm_choose_stages(const GemmDesc& g, int sm_count, int* katoms) {
    const long ctas = (long)gemm_ctas(g) * (g.ksplit > 1 ? g.ksplit : g.batch);
    *katoms = 1;
    return ctas <= (long)sm_count + sm_count / 4 ? UG_MAX_STAGES : 3;
}

// CTA count of the (unsplit) grid
static int gemm_ctas(const GemmDesc& g) {
    if (g.pair) return ((g.N + g.bn - 1) / g.bn) * ((g.m_tiles + 1) / 2 * 2);
    return ((g.N + UG_BN - 1) / UG_BN) * g.m_tiles;
}

// Switch a descriptor to the CTA-pair kernel when the problem allows it: full 128-row tiles, at least two of them, no batch.
// The tile width is N split evenly into ceil(N/256) tiles, rounded up to 32 (so each CTA stages a multiple of 16 B rows).
int gemm_enable_pair(GemmDesc* g) {
    if (g->batch != 1 || g->m_tiles < 2 || g->tile_rows != UG_BM || g->N < 64 || !g->b_ptr) return 0;
    const int n_tiles = (g->N + 255) / 256;
    int bn = ((g->N + n_tiles - 1) / n_tiles + 63) & ~63;      // multiple of 64: the epilogue's 64-column TMA boxes never straddle two tiles
    if (bn > 256) bn = 256;
    g->pair = 1; g->bn = bn; g->tmem_cols = bn <= 128 ? 128 : 256;
    long dims[4] = {g->Kd, g->N, 1, 1};
    long str[4] = {1, g->b_ld, (long)g->N * g->b_ld, (long)g->N * g->b_ld};
    int box[4] = {UG_BK, bn / 2, 1, 1};
    if (encode_bf16_map(&g->tmB, g->b_ptr, 4, dims, str, box, nullptr)) return -1;
    return 1;
}


// Output / residual / split-K-slab tensor maps of the TMA epilogue (pair kernel).  Call once out, residual, ksplit and
// splitk_ws are final.  Returns 1 when enabled, 0 when the register epilogue stays, < 0 on error.
int gemm_enable_tma_epilogue(GemmDesc* g) {
    g->tma_epi = 0;
    if (!g->pair || g->batch != 1) return 0;
    const long N = g->N;
    if (g->ksplit > 1) {
        if (g->conv) {
            long dims[5] = {N, g->Wo, g->Ho, g->NBimg, g->ksplit};
            long str[5] = {1, N, (long)g->Wo * N, (long)g->Ho * g->Wo * N, (long)g->M * N};
            int box[5] = {32, g->tw, g->th, g->tn, 1};
            if (encode_map(&g->tmP, true, g->splitk_ws, 5, dims, str, box, nullptr)) return -1;
        } else {
            long dims[4] = {N, g->M, g->ksplit, 1};
            long str[4] = {1, N, (long)g->M * N, (long)g->M * N * g->ksplit};
            int box[4] = {32, UG_BM, 1, 1};
            if (encode_map(&g->tmP, true, g->splitk_ws, 4, dims, str, box, nullptr)) return -1;
        }
        g->tma_epi = 2;
        return 1;
    }
    if (g->out_fp32 || (g->ldo & 7) || (N & 7) || (g->residual && (g->ldr & 7))) return 0;
    if (((uintptr_t)g->out & 15) || ((uintptr_t)g->residual & 15)) return 0;
    for (int pass = 0; pass < 2; ++pass) {
        const void* ptr = pass == 0 ? g->out : (const void*)g->residual;
        const long ld = pass == 0 ? g->ldo : g->ldr;
        if (!ptr) continue;
        CUtensorMap* m = pass == 0 ? &g->tmO : &g->tmR;
        if (g->conv) {
            long dims[4] = {N, g->Wo, g->Ho, g->NBimg};
            long str[4] = {1, ld, (long)g->Wo * ld, (long)g->Ho * g->Wo * ld};
            int box[4] = {64, g->tw, g->th, g->tn};
            if (encode_bf16_map(m, ptr, 4, dims, str, box, nullptr)) return -1;
        } else {
            long dims[4] = {N, g->M, 1, 1};
            long str[4] = {1, ld, (long)g->M * ld, (long)g->M * ld};
            int box[4] = {64, UG_BM, 1, 1};
            if (encode_bf16_map(m, ptr, 4, dims, str, box, nullptr)) return -1;
        }
    }
    g->tma_epi = 1;
    return 1;
}

// Dynamic shared memory of a pair-kernel launch: pipeline stages, barriers + bias (2 KB), residual boxes, alignment slack.
static size_t pair_smem_bytes(const GemmDesc& g, int stages) {
    const int res_boxes = (g.tma_epi == 1 && g.residual) ? (g.bn + 63) / 64 : 0;
    return (size_t)stages * UG2_STAGE_BYTES + 2048 + (size_t)res_boxes * 16384 + 1024;
}

int gemm_choose_ksplit(const GemmDesc& g, int sm_count) {
    if (g.batch != 1 || (g.N % 4)) return 1;
    const int ctas = gemm_ctas(g);
    if (ctas * 2 > sm_count) return 1;
    const int kpt = g.cin / UG_BK + ((g.cin % UG_BK) ? 1 : 0);
    const int n_k = g.taps * kpt;
    if (n_k < 40) return 1;                  // short reductions: the finalize pass would cost more than the split gains
    int ks = (2 * sm_count + ctas - 1) / ctas;
    if (ks > n_k / 8) ks = n_k / 8;          // keep at least 8 k-iterations per CTA
    return ks < 2 ? 1 : ks;
}

int gemm_launch(const GemmDesc& g, cudaStream_t st) {
    int dev = 0;
    cudaGetDevice(&dev);                                   // the opt-in shared-memory size is a per-device function attribute
    static thread_local bool configured_dev[64] = {false}, configured2_dev[64] = {false};
    bool& configured = configured_dev[dev & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(unet_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ug_smem_bytes(UG_MAX_STAGES, 1));
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    if (g.pair) {
        bool& configured2 = configured2_dev[dev & 63];
        if (!configured2) {
            cudaError_t e = cudaFuncSetAttribute(unet_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return (int)e;
            configured2 = true;
        }
        GemmDesc gg = g;
        gg.stages = (g.stages >= 2 && g.stages <= UG_MAX_STAGES) ? g.stages : 3;
        {   // output staging of the TMA epilogue lives in the (then idle) pipeline stages unless it is done in place on the residual
            const int boxes = g.tma_epi == 2 ? (g.bn + 31) / 32 : ((g.tma_epi == 1 && !g.residual) ? (g.bn + 63) / 64 : 0);
            const int need = (boxes * 16384 + UG2_STAGE_BYTES - 1) / UG2_STAGE_BYTES;
            if (gg.stages < need) gg.stages = need;
            while (gg.stages > 2 && pair_smem_bytes(g, gg.stages) > 227 * 1024) --gg.stages;
        }
        dim3 grid((g.m_tiles + 1) / 2 * 2, (g.N + g.bn - 1) / g.bn, g.ksplit > 1 ? g.ksplit : 1);
        // clusters of (2,1,1): cta_group::2 pairs form along x
        cudaError_t e = launch_k(unet_gemm_pair_kernel, grid, dim3(UG_THREADS), pair_smem_bytes(g, gg.stages), st, 2, gg);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess)
            fprintf(stderr, "uce: pair GEMM launch failed (%s): grid %u x %u x %u, smem %zu\n", cudaGetErrorString(e), grid.x, grid.y, grid.z,
                    pair_smem_bytes(g, gg.stages));
        if (e != cudaSuccess) return (int)e;
        if (g.ksplit > 1) {
            const long n = (long)g.M * (g.N / 4);
            const int rows_per_img = g.conv ? g.Ho * g.Wo : (g.rows_per_img > 0 ? g.rows_per_img : 1);
            e = launch_k(splitk_finalize_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, 1, g, rows_per_img);
        }
        return (int)e;
    }
    dim3 grid((g.N + UG_BN - 1) / UG_BN, g.m_tiles, g.ksplit > 1 ? g.ksplit : g.batch);
    GemmDesc gg = g;
    gg.katoms = (g.katoms == 2) ? 2 : 1;
    gg.stages = (g.stages >= 2 && g.stages * gg.katoms <= UG_MAX_STAGES) ? g.stages : 3;
    cudaError_t e = launch_k(unet_gemm_kernel, grid, dim3(UG_THREADS), ug_smem_bytes(gg.stages, gg.katoms), st, 1, gg);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    if (g.ksplit > 1) {
        const long n = (long)g.M * (g.N / 4);
        const int rows_per_img = g.conv ? g.Ho * g.Wo : (g.rows_per_img > 0 ? g.rows_per_img : 1);
        e = launch_k(splitk_finalize_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, 1, g, rows_per_img);
    }
    return (int)e;
}

}  // namespace uce
