// Fused low-rank apply, second generation: TWO co-resident CTAs per SM (rank_pad <= 64), everything that touches HBM or L2
// moves through TMA.
//
//   W_new[rows,:] = W_old[rows,:] + (W_old[rows,:] E^T) Q            (uce_sd_erase.py:45-82, see apply.cu)
//
// Same algebra, operands and fp32 fidelity (3xTF32, lo.lo dropped) as apply_tc.cu.  What changes is the shape of a CTA:
// apply_tc.cu runs ONE 352-thread CTA per SM with all 512 TMEM columns and 198 KB of shared memory, so the HBM read
// phase (A) and the L2-read / HBM-write phase (B) of a row tile run strictly one after the other on every SM, and the
// 200 row tiles of an SD-1.4 edit take two waves on 148 SMs (ncu: SMs 55 % active, profiles/r01_apply_tc_ncu.txt).
// Here a CTA is half as large — 224 threads, 256 TMEM columns, <= 105 KB of shared memory — so two of them share an SM:
//   * all row tiles of an SD-1.4 edit are resident at once (296 slots): no second wave;
//   * phase A of one tile overlaps phase B of the other: the SM's TMA ingest, tensor pipe and store path stay busy.
//
//   phase A  P[128,R] = W_tile[128,K] . E[R,K]^T in TMEM columns [0,R)
//            * W TMA warp: raw fp32 chunks [tile_rows x 32] into a 4-deep ring (L2 evict_last)
//            * 4 transform warps (thread = tile row = TMEM lane): hi = rna_tf32(w), lo = w - hi, tcgen05.st into one
//              of three 64-column A stages at TMEM columns [64,256)
//            * E TMA warp: [R x 32] tiles of the pre-split E_hi, E_lo (2-deep ring, L2 resident)
//            * MMA warp: per 8-wide k-step  hi.hi + hi.lo + lo.hi  (three N = R MMAs, A from TMEM)
//   phase B  dW[128 rows, 32 cols] = P[128,R] . Q[R, 32 cols] per UNIT of 32 W columns, four 32-column accumulators
//            at TMEM columns [128,256)
//            * P stays in tensor memory: read back, split, stored as P_hi [0,R) | P_lo [R,2R) — the A operand
//            * Qt_hi / Qt_lo [32 x 32] tiles (B operand) by TMA through a ring of 4 KB slots
//            * the W_old addend of a unit arrives by TMA in a [tile_rows x 32] box (ring of 5, L2 hits), the epilogue
//              warps add the accumulator IN PLACE (lane = row; the 128B swizzle makes the 16-byte accesses
//              conflict-free) and the box leaves through ONE TMA store.
//            Why not registers: the first version of this kernel loaded the addend with ld.global into registers, two
//            32-row groups ahead — 16 k cycles per 128 columns (profiles/r01_apply_tc2_timeline.txt): every wait on
//            a load scoreboard waits for ALL loads in flight on it, so register prefetch never overlapped anything.
//            TMA keeps three boxes (48 KB) in flight per CTA with no register or scoreboard involvement.
//
// tile_rows (<= 128, multiple of 8) is a launch parameter: the TMA box, the row stride between tiles and the rows a CTA
// stores; the MMAs always run M = 128 (rows beyond the box are zero and never stored: rows are independent).
#include "tc_apply_common.cuh"
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace uce {
namespace tc2 {
using namespace uce::tc;
using namespace uce::tca;

constexpr int PW = 4;                                   // transform / P-conversion / epilogue warps
constexpr int THREADS = (PW + 3) * 32;                  // + W TMA warp + E/Qt TMA warp + MMA warp
constexpr int NRAW = 5, NSA = 3, NE = 2;
constexpr int NB = 5;                                   // addend / output boxes (16 KB each)
constexpr int NQ = 2;                                   // Qt slots: all hi/lo tiles of one 32-column unit (<= 16 KB) per slot
constexpr int NACC = 4;                                 // 32-column accumulators
constexpr int MAX_LAYERS = 96;                          // two tensor maps per projection travel as kernel parameters
constexpr int WARP_W_TMA = PW, WARP_E_TMA = PW + 1, WARP_MMA = PW + 2;
constexpr uint32_t A_COL0 = 64;                         // A stages {W_hi 32 cols, W_lo 32 cols} at TMEM columns [64,256)
constexpr uint32_t ACC_COL0 = 128;
constexpr uint32_t TMEM_COLS = 256;

struct Maps { CUtensorMap e_hi, e_lo, qt_hi, qt_lo; };
struct WMaps { CUtensorMap in[MAX_LAYERS], out[MAX_LAYERS]; };

// Shared-memory carve-up (bytes), identical on host and device.  The dynamic shared-memory window of a kernel without
// static shared memory starts 1024-byte aligned (checked at run time), so no alignment slack is carried: two CTAs of
// 112.5 KB fit the SM's 228 KB only without it.
//   phase A   [0, 80K) raw ring: NRAW x 16 KB slots (a [tile_rows x 32] fp32 chunk of W each)
//             [80K, 80K + NE * e_stage) E ring: {E_hi R*128 B, E_lo R*128 B} per stage
//   phase B   [0, 80K) NB boxes of 16 KB (addend in, W_new out), then NQ Qt slots of 16 KB   (aliases phase A, used after it is drained)
struct Smem { int e_stage, e_off, q_off, bar_off, total; };
__host__ __device__ inline Smem smem_layout(int R) {
    Smem s;
    s.e_stage = 2 * R * 128;
    s.e_off = NRAW * 16384;
    s.q_off = NB * 16384;
    const int end_a = s.e_off + NE * s.e_stage;
    const int end_b = s.q_off + NQ * 16384;
    s.bar_off = end_a > end_b ? end_a : end_b;
    s.total = s.bar_off + 512;
    return s;
}

__global__ void __launch_bounds__(THREADS, 2)
apply_tc2_kernel(const LayerRef* __restrict__ layers, int n_layers, int K, int R, int tile_rows,
                 const __grid_constant__ Maps maps, const __grid_constant__ WMaps wmaps, long long* __restrict__ trace) {
    // optional timeline of CTA 0 (UCE_TC_TRACE=<file>): trace[(role * 64 + index) * 4 + event] = clock64()
    auto tr = [&](int role, int idx, int ev) {
        if (trace && blockIdx.x == 0 && idx < 64) trace[(role * 64 + idx) * 4 + ev] = clock64();
    };
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);                          // swizzled tiles need 1024-byte alignment
    if (base & 1023u) {
        if (threadIdx.x == 0) printf("uce apply_tc2: dynamic shared memory base %u is not 1024-byte aligned\n", base);
        __trap();
    }
    const Smem L = smem_layout(R);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- barriers ----
    const uint32_t bars = base + L.bar_off;
    auto bar_raw_full  = [&](int r) { return bars + 8u * r; };                 // [0,5)   W TMA -> transform warps
    auto bar_raw_empty = [&](int r) { return bars + 8u * (5 + r); };           // [5,10)
    auto bar_a_full    = [&](int s) { return bars + 8u * (10 + s); };          // [10,13) transform warps -> MMA (A stage written)
    auto bar_a_empty   = [&](int s) { return bars + 8u * (13 + s); };          // [13,16) MMA -> transform warps
    auto bar_e_full    = [&](int s) { return bars + 8u * (16 + s); };          // [16,18) E TMA -> MMA
    auto bar_e_empty   = [&](int s) { return bars + 8u * (18 + s); };          // [18,20)
    const uint32_t bar_p_full = bars + 8u * 20, bar_p_ready = bars + 8u * 21;
    auto bar_q_full    = [&](int t) { return bars + 8u * (22 + t); };          // [22,24) Qt TMA -> MMA (all tiles of a unit)
    auto bar_q_empty   = [&](int t) { return bars + 8u * (24 + t); };          // [24,26)
    auto bar_acc_full  = [&](int a) { return bars + 8u * (26 + a); };          // [26,30) MMA -> epilogue
    auto bar_acc_empty = [&](int a) { return bars + 8u * (30 + a); };          // [30,34)
    auto bar_box_full  = [&](int b) { return bars + 8u * (34 + b); };          // [34,39) addend TMA -> epilogue
    auto bar_box_ready = [&](int b) { return bars + 8u * (39 + b); };          // [39,44) epilogue -> W TMA warp (box holds W_new)
    const uint32_t tmem_slot = bars + 8u * 44;

    const int tile = blockIdx.x;
    const int layer = find_layer(layers, n_layers, tile);
    const LayerRef Lr = layers[layer];
    const int row0 = (tile - Lr.tile_begin) * tile_rows;
    const int rows_valid = min(tile_rows, Lr.d - row0);
    const int n_chunks = K / 32;          // phase A k-chunks (32 fp32 = one swizzle atom row) == phase B units of 32 W columns
    const int n_rc = R / 32;              // r atoms
    const uint32_t box_bytes = (uint32_t)tile_rows * 128u;

    if (threadIdx.x == 0) {
        for (int r = 0; r < NRAW; ++r) { mbar_init(bar_raw_full(r), 1); mbar_init(bar_raw_empty(r), PW); }
        for (int s = 0; s < NSA; ++s) { mbar_init(bar_a_full(s), PW); mbar_init(bar_a_empty(s), 1); }
        for (int s = 0; s < NE; ++s) { mbar_init(bar_e_full(s), 1); mbar_init(bar_e_empty(s), 1); }
        for (int t = 0; t < NQ; ++t) { mbar_init(bar_q_full(t), 1); mbar_init(bar_q_empty(t), 1); }
        mbar_init(bar_p_full, 1); mbar_init(bar_p_ready, PW);
        for (int a = 0; a < NACC; ++a) { mbar_init(bar_acc_full(a), 1); mbar_init(bar_acc_empty(a), PW); }
        for (int b = 0; b < NB; ++b) { mbar_init(bar_box_full(b), 1); mbar_init(bar_box_ready(b), PW); }
        mbar_fence_init();
    }
    if (warp == WARP_MMA) tmem_alloc(tmem_slot, TMEM_COLS);      // half of the SM's tensor memory: the co-resident CTA owns the rest
    if (warp == WARP_E_TMA && lane == 0) {
        tma_prefetch_desc(&wmaps.in[layer]); tma_prefetch_desc(&wmaps.out[layer]);
        tma_prefetch_desc(&maps.e_hi); tma_prefetch_desc(&maps.e_lo); tma_prefetch_desc(&maps.qt_hi); tma_prefetch_desc(&maps.qt_lo);
    }
    fence_before();
    __syncthreads();
    fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    auto raw_st = [&](int r) { return base + (uint32_t)(r * 16384); };
    auto stage_e_hi = [&](int s) { return base + (uint32_t)(L.e_off + s * L.e_stage); };
    auto stage_e_lo = [&](int s) { return base + (uint32_t)(L.e_off + s * L.e_stage + R * 128); };
    auto box_st = [&](int b) { return base + (uint32_t)(b * 16384); };
    auto qt_tile = [&](int t, int i) { return base + (uint32_t)(L.q_off + t * 16384 + i * 4096); };     // tile i = 2 * rc + (0 hi, 1 lo)

    if (warp < PW) {
        // =============================== W transform, then P conversion, then epilogue ===============================
        const int trow = 32 * warp + lane;               // tile row == TMEM lane
        const bool row_live = trow < rows_valid;
        const uint32_t lane_base = tmem_base + ((uint32_t)(32 * warp) << 16);
        const uint32_t row_off = (uint32_t)(trow * 128);
        const uint32_t sw = (uint32_t)(trow & 7);
        for (int c = 0; c < n_chunks; ++c) {
            const int r = c % NRAW, s = c % NSA;
            mbar_wait(bar_raw_full(r), (uint32_t)((c / NRAW) & 1));
            if (threadIdx.x == 0) tr(1, c, 0);
            const uint32_t raw = raw_st(r) + row_off;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row_live) v = lds_v4(raw + (((uint32_t)j ^ sw) << 4));      // swizzled 16-byte slot the TMA wrote
                const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float h = tf32_hi(x[e]);
                    hi[4 * j + e] = __float_as_uint(h);
                    lo[4 * j + e] = __float_as_uint(x[e] - h);
                }
            }
            mbar_wait(bar_a_empty(s), (uint32_t)(((c / NSA) & 1) ^ 1));      // the MMAs that read this A stage have completed
            if (threadIdx.x == 0) tr(1, c, 1);
            fence_after();
            const uint32_t ta = lane_base + A_COL0 + (uint32_t)(64 * s);
            tmem_st32(ta, hi);
            tmem_st32(ta + 32u, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_a_full(s)); mbar_arrive(bar_raw_empty(r)); }
            if (threadIdx.x == 0) tr(1, c, 2);
        }
        // ---- P: TMEM -> registers -> hi | lo -> TMEM columns [0,R) | [R,2R)  (A operand of phase B) ----
        mbar_wait(bar_p_full, 0);
        if (threadIdx.x == 0) tr(6, 0, 0);
        fence_after();
        for (int rc = 0; rc < n_rc; ++rc) {
            uint32_t v[32], lo[32];
            tmem_ld32(lane_base + (uint32_t)(rc * 32), v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float x = __uint_as_float(v[i]), h = tf32_hi(x);
                v[i] = __float_as_uint(h);
                lo[i] = __float_as_uint(x - h);
            }
            tmem_st32(lane_base + (uint32_t)(rc * 32), v);
            tmem_st32(lane_base + (uint32_t)(R + rc * 32), lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");      // before v / lo are written again (next rc)
        }
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p_ready);
        if (threadIdx.x == 0) tr(6, 0, 1);
        // ---- epilogue: per unit of 32 W columns, box += accumulator (in place, swizzled smem); the W TMA warp stores the box ----
        for (int u = 0; u < n_chunks; ++u) {
            const int b = u % NB, a = u % NACC;
            mbar_wait(bar_acc_full(a), (uint32_t)((u / NACC) & 1));
            fence_after();
            uint32_t v[32];
            tmem_ld32(lane_base + ACC_COL0 + (uint32_t)(32 * a), v);
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty(a));
            mbar_wait(bar_box_full(b), (uint32_t)((u / NB) & 1));
            if (threadIdx.x == 0) tr(5, u, 0);
            if (row_live) {
                const uint32_t row = box_st(b) + row_off;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t addr = row + (((uint32_t)j ^ sw) << 4);
                    const float4 w = lds_v4(addr);
                    sts_v4(addr, w.x + __uint_as_float(v[4 * j]), w.y + __uint_as_float(v[4 * j + 1]),
                           w.z + __uint_as_float(v[4 * j + 2]), w.w + __uint_as_float(v[4 * j + 3]));
                }
            }
            fence_proxy_async();                       // generic-proxy writes -> visible to the TMA store
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_box_ready(b));
            if (threadIdx.x == 0) tr(5, u, 1);
        }
    } else if (warp == WARP_W_TMA) {
        // =============================== TMA warp 1: raw W chunks (phase A), addend boxes (phase B) ===============================
        // (all lanes wait, one elected lane issues: see elect_one() in tc_common.cuh)
        const uint64_t pol_keep = l2_evict_last();     // the tile is read again by the epilogue: keep it in L2
        const uint64_t pol_last_use = l2_evict_first();
        const CUtensorMap* wm = &wmaps.in[layer];
        for (int c = 0; c < n_chunks; ++c) {
            const int r = c % NRAW;
            mbar_wait(bar_raw_empty(r), (uint32_t)(((c / NRAW) & 1) ^ 1));
            __syncwarp();
            if (elect_one()) {
                tr(0, c, 0);
                mbar_arrive_expect_tx(bar_raw_full(r), box_bytes);
                tma_load_2d_hint(raw_st(r), wm, bar_raw_full(r), c * 32, row0, pol_keep);
            }
        }
        // the boxes alias the raw / E rings: every phase-A MMA has completed once P is final, and an MMA on an A stage
        // completes only after all four transform warps have read that raw chunk
        mbar_wait(bar_p_full, 0);
        __syncwarp();
        const uint64_t pol_stream = l2_evict_first();
        const CUtensorMap* om = &wmaps.out[layer];
        if (elect_one()) {
            for (int u = 0; u < NB && u < n_chunks; ++u) {
                mbar_arrive_expect_tx(bar_box_full(u), box_bytes);
                tma_load_2d_hint(box_st(u), wm, bar_box_full(u), u * 32, row0, pol_last_use);
            }
        }
        __syncwarp();
        // box u: W_new is complete -> TMA store; once the PREVIOUS store has been read out of shared memory its box takes
        // the addend of unit u - 1 + NB.  (One thread issues every store: bulk async-groups are per thread.)
        for (int u = 0; u < n_chunks; ++u) {
            const int b = u % NB;
            mbar_wait(bar_box_ready(b), (uint32_t)((u / NB) & 1));
            __syncwarp();
            if (elect_one()) {
                tr(0, u, 1);
                tma_store_2d(om, box_st(b), u * 32, row0, pol_stream);
                tma_store_commit();
                const int nu = u - 1 + NB;
                if (u >= 1 && nu < n_chunks) {
                    tma_store_wait_read<1>();
                    const int nb = nu % NB;             // == (u - 1) % NB
                    mbar_arrive_expect_tx(bar_box_full(nb), box_bytes);
                    tma_load_2d_hint(box_st(nb), wm, bar_box_full(nb), nu * 32, row0, pol_last_use);
                }
            }
        }
        __syncwarp();
        if (elect_one()) tma_store_wait_read<0>();     // shared memory must outlive the last store's read
    } else if (warp == WARP_E_TMA) {
        // =============================== TMA warp 2: E tiles (phase A), Qt tiles (phase B) ===============================
        const uint32_t e_bytes = 2u * (uint32_t)R * 128u;
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % NE;
            mbar_wait(bar_e_empty(s), (uint32_t)(((c / NE) & 1) ^ 1));
            __syncwarp();
            if (elect_one()) {
                tr(2, c, 0);
                mbar_arrive_expect_tx(bar_e_full(s), e_bytes);
                tma_load_2d(stage_e_hi(s), &maps.e_hi, bar_e_full(s), c * 32, 0);
                tma_load_2d(stage_e_lo(s), &maps.e_lo, bar_e_full(s), c * 32, 0);
            }
        }
        mbar_wait(bar_p_full, 0);                      // the Qt slots alias the E ring
        const uint32_t q_bytes = (uint32_t)(n_rc * 2) * 4096u;
        for (int u = 0; u < n_chunks; ++u) {
            const int t = u % NQ;
            mbar_wait(bar_q_empty(t), (uint32_t)(((u / NQ) & 1) ^ 1));
            __syncwarp();
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_q_full(t), q_bytes);        // ONE barrier for all hi/lo tiles of the unit
                for (int rc = 0; rc < n_rc; ++rc) {
                    tma_load_2d(qt_tile(t, 2 * rc), &maps.qt_hi, bar_q_full(t), rc * 32, u * 32);
                    tma_load_2d(qt_tile(t, 2 * rc + 1), &maps.qt_lo, bar_q_full(t), rc * 32, u * 32);
                }
            }
        }
    } else {
        // =============================== MMA issuer ===============================
        // the whole warp runs the loops and the barrier waits (converged); ONE elected lane issues the MMAs and commits
        const uint32_t idesc_a = idesc_tf32(128, R);
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % NSA, se = c % NE;
            mbar_wait(bar_e_full(se), (uint32_t)((c / NE) & 1));
            mbar_wait(bar_a_full(s), (uint32_t)((c / NSA) & 1));
            fence_after();
            __syncwarp();
            if (elect_one()) {
                tr(3, c, 1);
                const uint32_t a_hi = tmem_base + A_COL0 + (uint32_t)(64 * s), a_lo = a_hi + 32u;
                const uint64_t b_hi = umma_desc_sw128(stage_e_hi(se)), b_lo = umma_desc_sw128(stage_e_lo(se));
#pragma unroll
                for (int k = 0; k < 4; ++k) {          // 8 tf32 per step: 8 TMEM columns of A, 32 bytes inside the swizzle atom of B
                    const uint64_t adv = (uint64_t)(k * 2);
                    umma_tf32_ts(tmem_base, a_hi + 8u * k, b_hi + adv, idesc_a, (c | k) != 0);
                    umma_tf32_ts(tmem_base, a_hi + 8u * k, b_lo + adv, idesc_a, 1);
                    umma_tf32_ts(tmem_base, a_lo + 8u * k, b_hi + adv, idesc_a, 1);
                }
                umma_commit(bar_a_empty(s));
                umma_commit(bar_e_empty(se));
                if (c == n_chunks - 1) umma_commit(bar_p_full);
                tr(3, c, 2);
            }
        }
        // ---- phase B: D[128 rows, 32 cols] = P_hi Qt_hi^T + P_lo Qt_hi^T + P_hi Qt_lo^T, A from tensor memory ----
        mbar_wait(bar_p_ready, 0);
        fence_after();
        const uint32_t idesc_b = idesc_tf32(128, 32);
        const uint32_t p_hi = tmem_base, p_lo = tmem_base + (uint32_t)R;
        for (int u = 0; u < n_chunks; ++u) {
            const int a = u % NACC, t = u % NQ;
            mbar_wait(bar_acc_empty(a), (uint32_t)(((u / NACC) & 1) ^ 1));
            mbar_wait(bar_q_full(t), (uint32_t)((u / NQ) & 1));
            fence_after();
            __syncwarp();
            if (elect_one()) {
                tr(4, u, 0);
                const uint32_t d_tmem = tmem_base + ACC_COL0 + 32u * (uint32_t)a;
                for (int rc = 0; rc < n_rc; ++rc) {
                    const uint64_t bq_hi = umma_desc_sw128(qt_tile(t, 2 * rc)), bq_lo = umma_desc_sw128(qt_tile(t, 2 * rc + 1));
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t adv = (uint64_t)(k * 2);
                        const uint32_t col = (uint32_t)(rc * 32 + 8 * k);
                        umma_tf32_ts(d_tmem, p_hi + col, bq_hi + adv, idesc_b, (rc | k) != 0);      // hi.hi
                        umma_tf32_ts(d_tmem, p_lo + col, bq_hi + adv, idesc_b, 1);                  // lo.hi
                        umma_tf32_ts(d_tmem, p_hi + col, bq_lo + adv, idesc_b, 1);                  // hi.lo
                    }
                }
                umma_commit(bar_q_empty(t));
                umma_commit(bar_acc_full(a));
                tr(4, u, 1);
            }
        }
    }
    // ---- teardown ----
    fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// row-major [rows, cols] fp32, box [box_rows, 32 cols], 128B swizzle
static int make_map(CUtensorMap* m, const float* ptr, int rows, int cols, int box_rows) {
    EncodeTiledFn enc = tensor_map_encoder();
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

}  // namespace tc2

bool apply_tc2_available(const uce_ws* ws) {
    const int R = ws->rank_pad;
    return ws->K % 128 == 0 && (R == 32 || R == 64) && !ws->dense && ws->rank > 0 && tensor_map_encoder() != nullptr;
}

// Rows per tile: UCE_TC2_TILE_ROWS (multiple of 8 in [8,128], read on every call) or 128.
int apply_tc2_tile_rows() {
    if (const char* e = getenv("UCE_TC2_TILE_ROWS")) {
        const int t = atoi(e);
        if (t >= 8 && t <= 128 && t % 8 == 0) return t;
    }
    return 128;
}

int apply_tc2_lowrank(uce_ws* ws, const LayerRef* layers_dev, const LayerRef* layers_host, int n_layers, int total_tiles,
                      int tile_rows, cudaStream_t st, int* launches) {
    using namespace tc2;
    const int K = ws->K, R = ws->rank_pad;
    if (!apply_tc2_available(ws)) { set_error("two-CTA tcgen05 apply unavailable for K=%d rank_pad=%d dense=%d", K, R, ws->dense); return UCE_E_STATE; }
    if (n_layers > MAX_LAYERS) { set_error("tcgen05 apply takes at most %d projections per call", MAX_LAYERS); return UCE_E_STATE; }
    for (int l = 0; l < n_layers; ++l)
        if (((uintptr_t)layers_host[l].w_old & 15) || ((uintptr_t)layers_host[l].w_new & 15)) { set_error("tcgen05 apply needs 16-byte aligned weights"); return UCE_E_ARG; }
    Maps maps;
    static thread_local WMaps wmaps;      // kept off the stack, one per host thread (handles are independent); copied into the launch by value
    int rc;
    if ((rc = make_map(&maps.e_hi, ws->E_hi, R, K, R))) return rc;
    if ((rc = make_map(&maps.e_lo, ws->E_lo, R, K, R))) return rc;
    if ((rc = make_map(&maps.qt_hi, ws->Qt_hi, K, R, 32))) return rc;
    if ((rc = make_map(&maps.qt_lo, ws->Qt_lo, K, R, 32))) return rc;
    for (int l = 0; l < n_layers; ++l) {
        if ((rc = make_map(&wmaps.in[l], layers_host[l].w_old, layers_host[l].d, K, tile_rows))) return rc;
        if ((rc = make_map(&wmaps.out[l], layers_host[l].w_new, layers_host[l].d, K, tile_rows))) return rc;
    }
    const Smem L = smem_layout(R);
    const int smem = L.total;          // no alignment slack: see smem_layout
    static thread_local int configured_dev[64] = {0};      // opt-in shared-memory size is a per-device function attribute
    int& configured = configured_dev[ws->device & 63];
    if (configured < smem) {
        UCE_CUDA(cudaFuncSetAttribute(apply_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        UCE_CUDA(cudaFuncSetAttribute(apply_tc2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured = smem;
    }
    long long* trace = nullptr;
    const char* trace_path = getenv("UCE_TC_TRACE");
    if (trace_path) { UCE_CUDA(cudaMalloc(&trace, 7 * 64 * 4 * sizeof(long long))); UCE_CUDA(cudaMemsetAsync(trace, 0, 7 * 64 * 4 * sizeof(long long), st)); }
    apply_tc2_kernel<<<total_tiles, THREADS, smem, st>>>(layers_dev, n_layers, K, R, tile_rows, maps, wmaps, trace);
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
