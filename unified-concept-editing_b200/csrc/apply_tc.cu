// Fused low-rank apply on the 5th-gen tensor cores (tcgen05 / TMEM / TMA), fp32 fidelity via 3xTF32.
//
//   W_new[rows,:] = W_old[rows,:] + (W_old[rows,:] E^T) Q            (uce_sd_erase.py:45-82, see apply.cu)
//
// One CTA per 128-row tile of one projection.  HBM traffic = read W_old once + write W_new once; the second
// read of the tile (epilogue addend) is an L2 hit.
//
//   phase A  P[128,R] = W_tile[128,K] . E[R,K]^T   accumulated in TMEM columns [0,R)
//            * 1 TMA warp streams raw fp32 W chunks [128 x 32] into a 5-deep smem ring (80 KB in flight per SM,
//              one tensor map per projection passed as a __grid_constant__ array, L2 evict_last)
//            * 8 transform warps split every value into hi = rna_tf32(w) and lo = w - hi and store both with
//              tcgen05.st into TENSOR MEMORY (two 64-column A stages): the MMA then takes A from TMEM and only the
//              small E tiles from shared memory.  (Earlier versions kept hi/lo in smem: phase A was bound by
//              shared-memory bandwidth, 152 KB of traffic per 16 KB chunk — see profiles/.)
//            * 1 TMA warp fetches the matching [R,32] tiles of the pre-split E_hi / E_lo (L2 resident)
//            * 1 MMA thread issues per 8-wide k-step  hi.hi + lo.hi + hi.lo  (tcgen05.mma kind::tf32, M=128, N=R)
//   phase B  dW^T[K-chunk of 128, 128 rows] = Qt[128,R] . P[128,R]^T in two ping-pong TMEM accumulators
//            * P is read back from TMEM, split hi/lo and written to smem as the B operand
//            * Qt_hi / Qt_lo tiles arrive by TMA; 3 MMAs per k-step again
//            * the transposed product puts consecutive W columns on consecutive TMEM lanes, so the epilogue's
//              per-register global accesses (W_old addend load, W_new store) are 128-byte coalesced per warp
//
// Requirements: K % 128 == 0, rank_pad in {32, 64, 96, 128}, 16-byte aligned weights.  Anything else goes to
// the SIMT kernels in apply.cu.
#include "uce_ws.h"
#include <cuda.h>
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace uce {

constexpr int TC_TILE_M = 128;
constexpr int TC_PRODUCER_WARPS = 8;
constexpr int TC_THREADS = (TC_PRODUCER_WARPS + 3) * 32;   // + W TMA warp + E/Qt TMA warp + MMA warp
constexpr int TC_NRAW = 5;                                  // raw W chunks in flight per SM (5 x 16 KB by TMA)
constexpr int TC_SA = 4;                                    // hi/lo A-operand stages in tensor memory
constexpr int TC_MAX_LAYERS = 160;                          // per-launch tensor maps for W_old (one per projection)
constexpr int WARP_W_TMA = TC_PRODUCER_WARPS, WARP_E_TMA = TC_PRODUCER_WARPS + 1, WARP_MMA = TC_PRODUCER_WARPS + 2;

struct TcMaps { CUtensorMap e_hi, e_lo, qt_hi, qt_lo; };
struct TcWMaps { CUtensorMap w[TC_MAX_LAYERS]; };

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if ((spin & 255u) == 255u) {          // watchdog: a protocol bug must fail loudly, never hang the GPU
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ll) {
                printf("uce apply_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
                __trap();
            }
        }
    }
}
// one lane of a converged warp (see tc_common.cuh: behind `lane == 0` ptxas wraps every UTCHMMA / UTMALDG in an ELECT loop)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

// K-major, 128B-swizzled operand tile (rows of 32 fp32 = 128 B, 8-row groups 1024 B apart), Blackwell descriptor v1
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, 16-byte units, bits [0,14)
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset between 8-row groups, bits [32,46)
    d |= (uint64_t)1 << 46;                            // descriptor version 1 (sm_100)
    d |= (uint64_t)2 << 61;                            // layout type SWIZZLE_128B
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from tensor memory (lane = row, one 32-bit column per k element), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
// L2 policies: the W_old tile is read twice (phase A from HBM, epilogue addend from L2) -> keep it (evict_last);
// W_new is written once and never read here -> evict_first, so that it does not push W_old tiles out of L2.
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
__device__ __forceinline__ float4 ldg_nc_v4(const float* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float ldg_f32(const float* p) {
    float r;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void stg_f32_hint(float* p, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ int tc_find_layer(const LayerRef* layers, int n_layers, int tile) {
    int lo = 0, hi = n_layers - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (layers[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Shared-memory carve-up (bytes), identical on host and device.
//   phase A   [0, 80K) raw ring: TC_NRAW x 16 KB fp32 W chunks written by TMA
//             [80K, 80K + ne*e_stage) E ring: ne x {E_hi R*128, E_lo R*128}  (deep enough to hide the L2 latency)
//             (the hi/lo split of W lives in TENSOR MEMORY, not in smem: see TC_A_COL0)
//   phase B   [0, nq*32K) Qt ring, [nq*32K, + 2*128*R*4) P_hi | P_lo   (aliases phase A; written after it is drained)
struct TcSmem {
    int ne, nq;
    int e_stage_bytes, e_off, p_off;
    int bar_off, total;
};
__host__ __device__ inline TcSmem tc_smem_layout(int R) {
    TcSmem s;
    s.ne = (R <= 64) ? 4 : (R <= 96 ? 3 : 2);
    s.nq = (R <= 64) ? 4 : (R <= 96 ? 3 : 2);
    s.e_stage_bytes = 2 * R * 128;
    s.e_off = TC_NRAW * 16384;
    s.p_off = s.nq * 32768;
    const int end_a = s.e_off + s.ne * s.e_stage_bytes;
    const int end_b = s.p_off + 2 * TC_TILE_M * R * 4;
    s.bar_off = end_a > end_b ? end_a : end_b;
    s.total = s.bar_off + 512;
    return s;
}
// Tensor-memory columns
//   phase A: [0,2R) P accumulator as two halves  hi.hi + lo.hi | hi.lo   (the B operand stacks E_hi over E_lo, so the
//            hi.hi and hi.lo products come out of ONE N = 2R MMA: the single issuing thread, ~60 cycles per
//            tcgen05.mma, is the limiter with 32-cycle MMAs — 2 instead of 3 instructions per k-step)
//            [256,512) four A stages {W_hi 32 cols, W_lo 32 cols}
//   phase B: two ping-pong accumulators of 256 columns, [0,256) and [256,512), again as two halves (B = P_hi over P_lo)
constexpr uint32_t TC_A_COL0 = 256;

__global__ void __launch_bounds__(TC_THREADS, 1)
apply_tc_kernel(const LayerRef* __restrict__ layers, int n_layers, int K, int R,
                const __grid_constant__ TcMaps maps, const __grid_constant__ TcWMaps wmaps, long long* __restrict__ trace) {
    // optional timeline of CTA 0 (UCE_TC_TRACE=<file>): trace[(role * 64 + index) * 4 + event] = clock64()
    auto tr = [&](int role, int idx, int ev) {
        if (trace && blockIdx.x == 0 && idx < 64) trace[(role * 64 + idx) * 4 + ev] = clock64();
    };
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // dynamic smem base is at least 16-byte aligned; swizzled tiles need 1024
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const TcSmem L = tc_smem_layout(R);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- barriers ----
    const uint32_t bars = base + L.bar_off;
    auto bar_raw_full  = [&](int r) { return bars + 8u * r; };                 // [0,5)   W TMA -> transform warps
    auto bar_raw_empty = [&](int r) { return bars + 8u * (5 + r); };           // [5,10)
    auto bar_full_w = [&](int s) { return bars + 8u * (10 + s); };             // [10,14) transform warps -> MMA (A stage in TMEM written)
    auto bar_empty  = [&](int s) { return bars + 8u * (14 + s); };             // [14,18) MMA -> transform warps
    auto bar_full_e  = [&](int s) { return bars + 8u * (18 + s); };            // [18,22) E TMA -> MMA
    auto bar_empty_e = [&](int s) { return bars + 8u * (22 + s); };            // [22,26) MMA -> E TMA
    const uint32_t bar_p_full = bars + 8u * 26, bar_p_smem = bars + 8u * 27;
    auto bar_q_full  = [&](int t) { return bars + 8u * (28 + t); };            // [28,32)
    auto bar_q_empty = [&](int t) { return bars + 8u * (32 + t); };            // [32,36)
    auto bar_acc_full  = [&](int b) { return bars + 8u * (36 + b); };
    auto bar_acc_empty = [&](int b) { return bars + 8u * (38 + b); };
    const uint32_t tmem_slot = bars + 8u * 41;

    const int tile = blockIdx.x;
    const int layer = tc_find_layer(layers, n_layers, tile);
    const LayerRef Lr = layers[layer];
    const int lt = tile - Lr.tile_begin;
    const int rows_valid = min(TC_TILE_M, Lr.d - lt * TC_TILE_M);
    const float* __restrict__ w_old = Lr.w_old + (size_t)lt * TC_TILE_M * K;
    float* __restrict__ w_new = Lr.w_new + (size_t)lt * TC_TILE_M * K;
    const int n_chunks = K / 32;          // phase A k-chunks (32 fp32 = one swizzle atom row)
    const int n_kc = K / 128;             // phase B chunks of 128 W columns
    const int n_rc = R / 32;              // r atoms

    if (threadIdx.x == 0) {
        for (int r = 0; r < TC_NRAW; ++r) { mbar_init(bar_raw_full(r), 1); mbar_init(bar_raw_empty(r), TC_PRODUCER_WARPS); }
        for (int s = 0; s < TC_SA; ++s) { mbar_init(bar_full_w(s), TC_PRODUCER_WARPS); mbar_init(bar_empty(s), 1); }
        for (int s = 0; s < 4; ++s) { mbar_init(bar_full_e(s), 1); mbar_init(bar_empty_e(s), 1); mbar_init(bar_q_full(s), 1); mbar_init(bar_q_empty(s), 1); }
        mbar_init(bar_p_full, 1); mbar_init(bar_p_smem, TC_PRODUCER_WARPS);
        for (int t = 0; t < 2; ++t) { mbar_init(bar_acc_full(t), 1); mbar_init(bar_acc_empty(t), TC_PRODUCER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WARP_MMA) {   // MMA warp owns the TMEM allocation (all 512 columns: one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == WARP_E_TMA && lane == 0) {
        tma_prefetch_desc(&wmaps.w[layer]);
        tma_prefetch_desc(&maps.e_hi); tma_prefetch_desc(&maps.e_lo); tma_prefetch_desc(&maps.qt_hi); tma_prefetch_desc(&maps.qt_lo);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int NE = L.ne, NQ = L.nq;
    auto raw_st = [&](int r) { return base + (uint32_t)(r * 16384); };
    auto stage_e_hi = [&](int s) { return base + (uint32_t)(L.e_off + s * L.e_stage_bytes); };
    auto stage_e_lo = [&](int s) { return base + (uint32_t)(L.e_off + s * L.e_stage_bytes + R * 128); };
    auto p_hi_atom = [&](int rc) { return base + (uint32_t)(L.p_off + rc * 32768); };              // [P_hi(rc) 128 rows | P_lo(rc) 128 rows]
    auto p_lo_atom = [&](int rc) { return base + (uint32_t)(L.p_off + rc * 32768 + 16384); };
    auto qt_hi_st = [&](int t) { return base + (uint32_t)(t * 32768); };
    auto qt_lo_st = [&](int t) { return base + (uint32_t)(t * 32768 + 16384); };

    if (warp < TC_PRODUCER_WARPS) {
        // =============================== W producers, then P converters, then epilogue ===============================
        // thread = one row of the tile (TMEM lane), half of the 32 k-columns of a chunk (warps 0-3: columns 0..15, 4-7: 16..31)
        const int tq = warp & 3, thalf = warp >> 2;
        const int trow = 32 * tq + lane;
        const uint64_t pol_stream = l2_policy_evict_first();
        for (int c = 0; c < n_chunks; ++c) {
            const int r = c % TC_NRAW, s = c % TC_SA;
            mbar_wait(bar_raw_full(r), (uint32_t)((c / TC_NRAW) & 1));
            if (threadIdx.x == 0) tr(1, c, 0);
            const uint32_t raw = raw_st(r);
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = lds_v4(raw + (uint32_t)(trow * 128 + (((4 * thalf + j) ^ (trow & 7)) << 4)));     // swizzled slot the TMA wrote
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float x[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float h = tf32_hi(x[e]);
                    hi[4 * j + e] = __float_as_uint(h);
                    lo[4 * j + e] = __float_as_uint(x[e] - h);
                }
            }
            mbar_wait(bar_empty(s), (uint32_t)(((c / TC_SA) & 1) ^ 1));      // the MMAs that read this A stage have completed
            if (threadIdx.x == 0) tr(1, c, 1);
            tc_fence_after();
            const uint32_t ta = tmem_base + ((uint32_t)(32 * tq) << 16) + TC_A_COL0 + (uint32_t)(64 * s + 16 * thalf);
            tmem_st16(ta, hi);
            tmem_st16(ta + 32u, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_full_w(s)); mbar_arrive(bar_raw_empty(r)); }
            if (threadIdx.x == 0) tr(1, c, 2);
        }
        // ---- P: TMEM -> registers -> hi/lo -> swizzled smem (B operand of phase B) ----
        const int q = warp & 3, half = warp >> 2;
        const int prow = 32 * q + lane;            // TMEM lane == row of the tile
        mbar_wait(bar_p_full, 0);
        if (threadIdx.x == 0) tr(6, 0, 0);
        tc_fence_after();
        for (int rc = half; rc < n_rc; rc += 2) {
            uint32_t v[32], v2[32];
            tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(rc * 32), v);
            tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(R + rc * 32), v2);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
            const uint32_t hb = p_hi_atom(rc), lb = p_lo_atom(rc);
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                const uint32_t off = (uint32_t)(prow * 128 + ((ch ^ (prow & 7)) << 4));
                const float a = __uint_as_float(v[4 * ch]), b = __uint_as_float(v[4 * ch + 1]);
                const float cc = __uint_as_float(v[4 * ch + 2]), d = __uint_as_float(v[4 * ch + 3]);
                const float ha = tf32_hi(a), hb2 = tf32_hi(b), hc = tf32_hi(cc), hd = tf32_hi(d);
                sts_v4(hb + off, ha, hb2, hc, hd);
                sts_v4(lb + off, a - ha, b - hb2, cc - hc, d - hd);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p_smem);
        // ---- epilogue: W_new = W_old + dW, transposed accumulator (lane = W column, register = W row) ----
        for (int kc = 0; kc < n_kc; ++kc) {
            const int b = kc & 1;
            const int col = kc * 128 + 32 * q + lane;
            const int jrow0 = 64 * half;                     // this warp's 64 tile rows (two TMEM loads of 32 columns)
            // addend W_old[jrow0 .. +64, col]: issued BEFORE waiting for the accumulator so the L2 latency hides under the MMAs
            float w[64];
            const float* wp = w_old + (size_t)jrow0 * K + col;
            if (rows_valid == TC_TILE_M) {
#pragma unroll
                for (int i = 0; i < 64; ++i) w[i] = ldg_f32(wp + (size_t)i * K);
            } else {
#pragma unroll
                for (int i = 0; i < 64; ++i) w[i] = (jrow0 + i < rows_valid) ? ldg_f32(wp + (size_t)i * K) : 0.f;
            }
            if (threadIdx.x == 0) tr(5, kc, 0);
            mbar_wait(bar_acc_full(b), (uint32_t)((kc >> 1) & 1));
            if (threadIdx.x == 0) tr(5, kc, 1);
            tc_fence_after();
            float* op = w_new + (size_t)jrow0 * K + col;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                uint32_t v[32], v2[32];
                tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(256 * b + jrow0 + 32 * g), v);
                tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(256 * b + 128 + jrow0 + 32 * g), v2);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
                if (rows_valid == TC_TILE_M) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) stg_f32_hint(op + (size_t)(32 * g + i) * K, w[32 * g + i] + __uint_as_float(v[i]), pol_stream);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (jrow0 + 32 * g + i < rows_valid) stg_f32_hint(op + (size_t)(32 * g + i) * K, w[32 * g + i] + __uint_as_float(v[i]), pol_stream);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty(b));
            if (threadIdx.x == 0) tr(5, kc, 2);
        }
    } else if (warp == WARP_W_TMA) {
        // =============================== TMA warp 1: raw W chunks, TC_NRAW deep ===============================
        const uint64_t pol_keep = l2_policy_evict_last();     // the tile is read again by the epilogue: keep it in L2
        const CUtensorMap* wm = &wmaps.w[layer];
        for (int c = 0; c < n_chunks; ++c) {
            const int r = c % TC_NRAW;
            mbar_wait(bar_raw_empty(r), (uint32_t)(((c / TC_NRAW) & 1) ^ 1));
            __syncwarp();
            if (elect_one()) {
                tr(0, c, 0);
                mbar_arrive_expect_tx(bar_raw_full(r), 16384u);
                tma_load_2d_hint(raw_st(r), wm, bar_raw_full(r), c * 32, lt * TC_TILE_M, pol_keep);
            }
        }
    } else if (warp == WARP_E_TMA) {
        // =============================== TMA warp 2: E tiles (phase A), Qt tiles (phase B) ===============================
        const uint32_t e_bytes = 2u * (uint32_t)R * 128u;
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % NE;
            mbar_wait(bar_empty_e(s), (uint32_t)(((c / NE) & 1) ^ 1));
            __syncwarp();
            if (elect_one()) {
                tr(2, c, 0);
                mbar_arrive_expect_tx(bar_full_e(s), e_bytes);
                tma_load_2d(stage_e_hi(s), &maps.e_hi, bar_full_e(s), c * 32, 0);
                tma_load_2d(stage_e_lo(s), &maps.e_lo, bar_full_e(s), c * 32, 0);
            }
        }
        // the Qt ring aliases the raw ring and the E ring: both must be drained (all phase A MMAs complete)
        for (int r = 0; r < TC_NRAW; ++r) {
            const int uses = (n_chunks - r + TC_NRAW - 1) / TC_NRAW;      // completed phases of raw_empty[r]
            if (uses > 0) mbar_wait(bar_raw_empty(r), (uint32_t)((uses - 1) & 1));
        }
        mbar_wait(bar_p_full, 0);
        int it = 0;
        for (int kc = 0; kc < n_kc; ++kc)
            for (int rc = 0; rc < n_rc; ++rc, ++it) {
                const int t = it % NQ;
                const uint32_t ph = (uint32_t)((it / NQ) & 1);
                mbar_wait(bar_q_empty(t), ph ^ 1u);
                __syncwarp();
                if (elect_one()) {
                    mbar_arrive_expect_tx(bar_q_full(t), 32768u);
                    tma_load_2d(qt_hi_st(t), &maps.qt_hi, bar_q_full(t), rc * 32, kc * 128);
                    tma_load_2d(qt_lo_st(t), &maps.qt_lo, bar_q_full(t), rc * 32, kc * 128);
                }
            }
    } else {
        // =============================== MMA issuer ===============================
        // the whole warp runs the loops and the barrier waits (converged); ONE elected lane issues the MMAs and commits
        const uint32_t idesc_a = umma_idesc_tf32(128, 2 * R), idesc_a2 = umma_idesc_tf32(128, R);
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % TC_SA, se = c % NE;
            mbar_wait(bar_full_w(s), (uint32_t)((c / TC_SA) & 1));
            mbar_wait(bar_full_e(se), (uint32_t)((c / NE) & 1));
            tc_fence_after();
            __syncwarp();
            if (elect_one()) {
                tr(3, c, 1);
                const uint32_t a_hi = tmem_base + TC_A_COL0 + (uint32_t)(64 * s), a_lo = a_hi + 32u;
                const uint64_t b_hi = umma_desc_sw128(stage_e_hi(se));     // E_hi rows followed by E_lo rows: one 2R-row tile
#pragma unroll
                for (int k = 0; k < 4; ++k) {          // 8 tf32 per step: 8 TMEM columns of A, 32 bytes inside the swizzle atom of B
                    const uint64_t adv = (uint64_t)(k * 2);
                    umma_tf32_ts(tmem_base, a_hi + 8u * k, b_hi + adv, idesc_a, (c | k) != 0);     // N = 2R: [hi.hi | hi.lo]
                    umma_tf32_ts(tmem_base, a_lo + 8u * k, b_hi + adv, idesc_a2, 1);               // N = R : += lo.hi on the first half
                }
                umma_commit(bar_empty(s));
                umma_commit(bar_empty_e(se));
                if (c == n_chunks - 1) umma_commit(bar_p_full);
                tr(3, c, 2);
            }
        }
        // ---- phase B ----
        mbar_wait(bar_p_smem, 0);
        tc_fence_after();
        const uint32_t idesc_b = umma_idesc_tf32(128, 256), idesc_b2 = umma_idesc_tf32(128, 128);
        int it = 0;
        for (int kc = 0; kc < n_kc; ++kc) {
            const int b = kc & 1;
            mbar_wait(bar_acc_empty(b), (uint32_t)(((kc >> 1) & 1) ^ 1));
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + 256u * (uint32_t)b;
            for (int rc = 0; rc < n_rc; ++rc, ++it) {
                const int t = it % NQ;
                const uint32_t ph = (uint32_t)((it / NQ) & 1);
                mbar_wait(bar_q_full(t), ph);
                tc_fence_after();
                __syncwarp();
                if (elect_one()) {
                    if (rc == 0) tr(4, kc, 0);
                    const uint64_t a_hi = umma_desc_sw128(qt_hi_st(t)), a_lo = umma_desc_sw128(qt_lo_st(t));
                    const uint64_t b_hi = umma_desc_sw128(p_hi_atom(rc));      // P_hi rows followed by P_lo rows: 256-row tile
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t adv = (uint64_t)(k * 2);
                        umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc_b, (rc | k) != 0);     // N = 256: [hi.hi | hi.lo]
                        umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc_b2, 1);                // N = 128: += lo.hi
                    }
                    umma_commit(bar_q_empty(t));
                    if (rc == n_rc - 1) { umma_commit(bar_acc_full(b)); tr(4, kc, 1); }
                }
            }
        }
    }
    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// hi = rna_tf32(x), lo = x - hi for the small pre-split operands (E and Qt)
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long n) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float v = x[i], h = tf32_hi(v); hi[i] = h; lo[i] = v - h; }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// row-major [rows, cols] fp32, box [box_rows, 32 cols], 128B swizzle
static int make_map(CUtensorMap* m, const float* ptr, int rows, int cols, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not found"); return UCE_E_STATE; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return UCE_E_STATE; }
    return 0;
}

bool apply_tc_available(const uce_ws* ws) {
    const int R = ws->rank_pad;
    return ws->K % 128 == 0 && R >= 32 && R <= 128 && R % 32 == 0 && !ws->dense && ws->rank > 0 && get_encode() != nullptr;
}

int apply_tc_split_operands(uce_ws* ws, cudaStream_t st, int* launches) {
    const long n = (long)ws->rank_pad * ws->K;
    split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws->E, ws->E_hi, ws->E_lo, n);
    UCE_LAUNCH_CHECK();
    split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws->Qt, ws->Qt_hi, ws->Qt_lo, n);
    UCE_LAUNCH_CHECK();
    *launches += 2;
    return 0;
}

int apply_tc_lowrank(uce_ws* ws, const LayerRef* layers_dev, const LayerRef* layers_host, int n_layers, int total_tiles,
                     cudaStream_t st, int* launches) {
    const int K = ws->K, R = ws->rank_pad;
    if (!apply_tc_available(ws)) { set_error("tcgen05 apply unavailable for K=%d rank_pad=%d dense=%d", K, R, ws->dense); return UCE_E_STATE; }
    if (n_layers > TC_MAX_LAYERS) { set_error("tcgen05 apply takes at most %d projections per call", TC_MAX_LAYERS); return UCE_E_STATE; }
    for (int l = 0; l < n_layers; ++l)
        if (((uintptr_t)layers_host[l].w_old & 15) || ((uintptr_t)layers_host[l].w_new & 15)) { set_error("tcgen05 apply needs 16-byte aligned weights"); return UCE_E_ARG; }
    TcMaps maps;
    static thread_local TcWMaps wmaps;      // kept off the stack, one per host thread (handles are independent); copied into the launch by value
    int rc;
    if ((rc = make_map(&maps.e_hi, ws->E_hi, R, K, R))) return rc;
    if ((rc = make_map(&maps.e_lo, ws->E_lo, R, K, R))) return rc;
    if ((rc = make_map(&maps.qt_hi, ws->Qt_hi, K, R, 128))) return rc;
    if ((rc = make_map(&maps.qt_lo, ws->Qt_lo, K, R, 128))) return rc;
    for (int l = 0; l < n_layers; ++l)
        if ((rc = make_map(&wmaps.w[l], layers_host[l].w_old, layers_host[l].d, K, TC_TILE_M))) return rc;
    const TcSmem L = tc_smem_layout(R);
    const int smem = L.total + 1024;   // slack for the manual 1024-byte alignment
    static thread_local int configured_dev[64] = {0};      // opt-in shared-memory size is a per-device function attribute
    int& configured = configured_dev[ws->device & 63];
    if (configured < smem) {
        UCE_CUDA(cudaFuncSetAttribute(apply_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    long long* trace = nullptr;
    const char* trace_path = getenv("UCE_TC_TRACE");
    if (trace_path) { UCE_CUDA(cudaMalloc(&trace, 7 * 64 * 4 * sizeof(long long))); UCE_CUDA(cudaMemsetAsync(trace, 0, 7 * 64 * 4 * sizeof(long long), st)); }
    apply_tc_kernel<<<total_tiles, TC_THREADS, smem, st>>>(layers_dev, n_layers, K, R, maps, wmaps, trace);
    UCE_LAUNCH_CHECK();
    *launches += 1;
    if (trace) {   // debugging aid: dump the timeline of CTA 0 (synchronises)
        std::vector<long long> h(7 * 64 * 4);
        UCE_CUDA(cudaStreamSynchronize(st));
        UCE_CUDA(cudaMemcpy(h.data(), trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        UCE_CUDA(cudaFree(trace));
        if (FILE* f = fopen(trace_path, "w")) {
            long long t0 = 0;
            for (long long v : h) if (v && (!t0 || v < t0)) t0 = v;
            const char* roles[7] = {"w_tma", "transform", "e_tma", "mma_a", "mma_b", "epilogue", "pconv"};
            for (int r = 0; r < 7; ++r) for (int i = 0; i < 64; ++i) {
                const long long* e = &h[(r * 64 + i) * 4];
                if (e[0] || e[1] || e[2]) fprintf(f, "%s %d %lld %lld %lld\n", roles[r], i, e[0] ? e[0] - t0 : -1, e[1] ? e[1] - t0 : -1, e[2] ? e[2] - t0 : -1);
            }
            fclose(f);
        }
    }
    return 0;
}

}  // namespace uce
