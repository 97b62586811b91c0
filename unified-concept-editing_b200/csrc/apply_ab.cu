// Low-rank apply (rank pad <= 64) as TWO tcgen05 kernels with the K dimension split over CTAs:
//
//   kernel A   Ppart_j[rows, R] = W_old[rows, cols_j] . E[:, cols_j]^T          cols_j = j-th slice of K, j < ks
//              (the reference's guide outputs v* = W_old c folded with the edit rows, uce_sd_erase.py:45-53)
//   kernel B   W_new[rows, cols_j] = W_old[rows, cols_j] + (sum_j' Ppart_j'[rows, :]) . Q[:, cols_j]
//              (mat1 @ inverse(mat2), uce_sd_erase.py:61-82)
//
// Why two kernels and a K split (measurements of the fused one-kernel form, apply_tc3.cu, profiles/r01_apply_tc3_*):
//   * every CTA of the fused kernel streams ALL of E_hi|E_lo and Qt_hi|Qt_lo (2 x 393 KB at R = 64, K = 768) next to 2 x 0.52 MB of
//     W: the SM's TMA ingest, not HBM, bounds it.  A CTA that owns a 1/ks slice of K needs only that slice of E and Qt: with
//     ks = 2 the low-rank operands cost 0.2 MB per kernel and CTA instead of 0.39.
//   * the price of the split is that P must be summed over the slices: the partial products go through an L2-resident scratch
//     (ks x 6.4 MB for SD-1.4) and kernel B adds them in a fixed order while it loads them — no atomics, no cluster exchange.
//   * a CTA carries THREE row blocks (TMEM: 3 x 64 accumulator columns + a 5-deep ring of A stages in kernel A; 3 x 128 columns
//     of P_hi|P_lo + 4 accumulators in kernel B) of <= 128 rows, planned on the host so that one wave of CTAs covers the edit:
//     SD-1.4 -> 222 blocks of 104-120 rows on 74 x 2 CTAs (the fused kernel ran 288 blocks of 80-96 rows through M = 128 MMAs).
//   * kernel A needs only E = G_e - C_e, not Q: the caller may run it on a second stream WHILE the factor computes Q
//     (uce_edit_dev_f32); only kernel B is on the critical path behind the factor.
//   * tensor-core accumulators round toward zero, so the error of a sum grows with the length of the in-TMEM accumulation
//     chain (measured: 7e-9 x K relative to the update for the fused kernel).  The split shortens the chain to K / ks / 8 steps
//     of three MMAs; ks is chosen so that a slice has at most 512 columns.
//
// Both kernels: 8 transform / epilogue warps (thread = row = TMEM lane), one W TMA warp, one E / Qt TMA warp, and the MMA issue —
// kernel A: ONE warp (all MMAs into an accumulator must come from one thread: only then are they ordered), every work item goes
// through all eight transform warps; kernel B: two issuers and two epilogue sets, every other item each (an accumulator belongs
// to one item).  fp32 fidelity through 3xTF32 (hi.hi + hi.lo + lo.hi; W and P split on the fly into tensor
// memory, E and Qt pre-split by the factor).  Every mbarrier wait carries the clock watchdog of tc_common.cuh.
#include "tc_apply_common.cuh"
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace uce {
namespace ab {
using namespace uce::tc;
using namespace uce::tca;

constexpr int NBLK = 3;                                 // row blocks per CTA
constexpr int TW = 8;                                   // transform / epilogue warps (two sets of four)
constexpr int WARP_W = TW, WARP_B = TW + 1, WARP_MMA = TW + 2;
constexpr int THREADS_A = (TW + 3) * 32;                // + W TMA warp + E TMA warp + ONE MMA issuer
constexpr int THREADS = (TW + 4) * 32;                  // kernel B: + W TMA warp + Qt TMA warp + two MMA issuers
constexpr int MAX_LAYERS = 96;
constexpr uint32_t TMEM_COLS = 512;
// kernel A
constexpr int NRAW = 8, NA = 5, NE = 4;
constexpr uint32_t A_COL0 = 64u * NBLK;                 // accumulators [64 g, 64 g + R), A stages {W_hi 32, W_lo 32} behind them
// kernel B
constexpr int NBX = 10, NQ = 3, NACC = 4;
constexpr uint32_t PLO_COL0 = 64u * NBLK, ACC_COL0 = 128u * NBLK;

struct Slot { int layer, row0, rows, h; };               // one row block: rows [row0, row0 + rows) of a projection, TMA box height h
struct EMaps { CUtensorMap hi, lo; };
struct WIn { CUtensorMap in[MAX_LAYERS]; };
struct WIo { CUtensorMap in[MAX_LAYERS], out[MAX_LAYERS]; };

__host__ __device__ inline int smem_a(int R) { return NRAW * 16384 + NE * (2 * R * 128) + 1024; }
__host__ __device__ inline int smem_b() { return NBX * 16384 + NQ * 16384 + 1024; }

struct Cta {
    int q, j, n_act, ncs, c0, rot, n_items;
    Slot sl[NBLK];
};
__device__ __forceinline__ Cta cta_setup(const Slot* __restrict__ slots, int n_slots, int K, int ks) {
    Cta c;
    c.q = (int)blockIdx.x / ks; c.j = (int)blockIdx.x % ks;
    const int s0 = c.q * NBLK;
    c.n_act = min(NBLK, n_slots - s0);
#pragma unroll
    for (int g = 0; g < NBLK; ++g) c.sl[g] = (g < c.n_act) ? slots[s0 + g] : Slot{0, 0, 0, 8};
    c.ncs = K / 32 / ks;                                 // 32-column chunks (kernel A) / units (kernel B) of this CTA's slice of K
    c.c0 = c.j * c.ncs;
    // every CTA of a slice streams the SAME E / Qt tiles: CTA groups walk the slice from different offsets so that they do not
    // all ask the same L2 lines at the same time
    c.rot = (int)(((unsigned)c.q * 5u) % (unsigned)c.ncs);
    c.n_items = c.ncs * c.n_act;
    return c;
}
__device__ __forceinline__ int col_of(const Cta& c, int ci) { int x = ci + c.rot; if (x >= c.ncs) x -= c.ncs; return (c.c0 + x) * 32; }

// ------------------------------------------------------------------------------------------------ kernel A
// Barrier discipline (both kernels): mbarrier waits test the PARITY of a phase, so a waiter must never be two phases away from
// the barrier it tests — every barrier here has one fixed set of waiters that observes every one of its phases.  (A first version
// let two transform sets and two MMA issuers take every other work item over a 5-deep ring: a waiter then skipped every other
// phase of a slot's barriers, could pass a wait one phase early, read a half-written A stage and double-arrive — wrong rows,
// run-to-run differences and hangs on hardware.)  Kernel A therefore runs every work item through ALL eight transform warps
// (each takes its row's 16 columns of the 32-column chunk) and ONE MMA-issuing warp; all MMAs into an accumulator come from
// one thread, which also orders them.
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

__global__ void __launch_bounds__(THREADS_A, 1)
apply_p_kernel(const Slot* __restrict__ slots, int n_slots, int K, int R, int ks, float* __restrict__ Pbuf,
               const __grid_constant__ EMaps emaps, const __grid_constant__ WIn wmaps, long long* __restrict__ trace) {
    // optional timeline of CTA 0 (UCE_AB_TRACE=<file>): trace[(role * 64 + index) * 4 + event] = clock64()
    auto tr = [&](int role, int idx, int ev) {
        if (trace && blockIdx.x == 0 && idx < 64) trace[(role * 64 + idx) * 4 + ev] = clock64();
    };
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if (threadIdx.x == 0) tr(6, 2, 0);
    const uint32_t base = smem_u32(smem_raw);
    if (base & 1023u) {
        if (threadIdx.x == 0) printf("uce apply_p: dynamic shared memory base %u is not 1024-byte aligned\n", base);
        __trap();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Cta c = cta_setup(slots, n_slots, K, ks);
    const int e_stage = 2 * R * 128;
    const uint32_t bars = base + (uint32_t)(NRAW * 16384 + NE * e_stage);
    auto bar_raw_full  = [&](int r) { return bars + 8u * r; };                 // [0,8)   W TMA -> transform warps
    auto bar_raw_empty = [&](int r) { return bars + 8u * (8 + r); };           // [8,16)
    auto bar_a_full    = [&](int a) { return bars + 8u * (16 + a); };          // [16,21) transform warps -> MMA issuer
    auto bar_a_empty   = [&](int a) { return bars + 8u * (21 + a); };          // [21,26)
    auto bar_e_full    = [&](int s) { return bars + 8u * (26 + s); };          // [26,30) E TMA -> MMA issuer
    auto bar_e_empty   = [&](int s) { return bars + 8u * (30 + s); };          // [30,34)
    const uint32_t bar_p_full = bars + 8u * 34, tmem_slot = bars + 8u * 35;
    auto raw_st = [&](int r) { return base + (uint32_t)(r * 16384); };
    auto e_hi_st = [&](int s) { return base + (uint32_t)(NRAW * 16384 + s * e_stage); };
    auto e_lo_st = [&](int s) { return base + (uint32_t)(NRAW * 16384 + s * e_stage + R * 128); };

    if (threadIdx.x == 0) {
        for (int r = 0; r < NRAW; ++r) { mbar_init(bar_raw_full(r), 1); mbar_init(bar_raw_empty(r), TW); }
        for (int a = 0; a < NA; ++a) { mbar_init(bar_a_full(a), TW); mbar_init(bar_a_empty(a), 1); }
        for (int s = 0; s < NE; ++s) { mbar_init(bar_e_full(s), 1); mbar_init(bar_e_empty(s), 1); }
        mbar_init(bar_p_full, 1);
        mbar_fence_init();
    }
    if (warp == WARP_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
    if (warp == WARP_B && lane == 0) {
        tma_prefetch_desc(&emaps.hi); tma_prefetch_desc(&emaps.lo);
        for (int g = 0; g < c.n_act; ++g) tma_prefetch_desc(&wmaps.in[c.sl[g].layer]);
    }
    fence_before();
    __syncthreads();
    fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (threadIdx.x == 0) tr(6, 2, 1);

    if (warp < TW) {
        // =============================== W transform: every warp on every item (its rows x 16 of the chunk's 32 columns) ===============================
        const int half = warp >> 2, wq = warp & 3;
        const int trow = 32 * wq + lane;                 // row of the block == TMEM lane
        const uint32_t lane_base = tmem_base + ((uint32_t)(32 * wq) << 16);
        const uint32_t row_off = (uint32_t)(trow * 128);
        const uint32_t sw = (uint32_t)(trow & 7);
        int g = 0;
        for (int i = 0; i < c.n_items; ++i) {
            const int r = i % NRAW, a = i % NA;
            mbar_wait(bar_raw_full(r), (uint32_t)((i / NRAW) & 1));
            if (threadIdx.x == 0) tr(1, i, 0);
            const bool row_live = trow < c.sl[g].rows;
            const uint32_t raw = raw_st(r) + row_off;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int jv = 0; jv < 4; ++jv) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row_live) v = lds_v4(raw + (((uint32_t)(4 * half + jv) ^ sw) << 4));      // swizzled 16-byte slot the TMA wrote
                tf32_split(v.x, hi[4 * jv], lo[4 * jv]);         tf32_split(v.y, hi[4 * jv + 1], lo[4 * jv + 1]);
                tf32_split(v.z, hi[4 * jv + 2], lo[4 * jv + 2]); tf32_split(v.w, hi[4 * jv + 3], lo[4 * jv + 3]);
            }
            mbar_wait(bar_a_empty(a), (uint32_t)(((i / NA) & 1) ^ 1));       // the MMAs that read this A stage have completed
            if (threadIdx.x == 0) tr(1, i, 1);
            fence_after();
            __syncwarp();
            const uint32_t ta = lane_base + A_COL0 + 64u * (uint32_t)a + 16u * (uint32_t)half;
            tmem_st16(ta, hi);
            tmem_st16(ta + 32u, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_a_full(a)); mbar_arrive(bar_raw_empty(r)); }
            if (threadIdx.x == 0) tr(1, i, 2);
            if (++g == c.n_act) g = 0;
        }
        // ---- the partial products of this K slice: TMEM -> registers -> scratch [slice][slot][128 rows][R]; a warp takes half the columns ----
        mbar_wait(bar_p_full, 0);
        if (threadIdx.x == 0) tr(6, 0, 0);
        fence_after();
        __syncwarp();
        // Scratch layout [slice][slot][R / 4 column quads][128 rows][4]: lane = row, so the 32 lanes of a store instruction write 512
        // consecutive bytes (row-major [row][R] cost one L1 line per LANE: 32 cycles per store instruction, 6 k cycles for this drain,
        // and 15 k cycles for the matching loads in kernel B).
        const int cw = R / 2;                                        // columns per warp half: 32 (R = 64) or 16 (R = 32)
        for (int gb = 0; gb < c.n_act; ++gb) {
            float4* dst = reinterpret_cast<float4*>(Pbuf) + (((size_t)c.j * n_slots + (size_t)(c.q * NBLK + gb)) * (R / 4) + (size_t)(half * cw / 4)) * 128 + trow;
            const bool live = trow < c.sl[gb].rows;
            uint32_t v[32];
            if (cw == 32) {
                tmem_ld32(lane_base + 64u * (uint32_t)gb + 32u * (uint32_t)half, v);
            } else {
                uint32_t w[16];
                tmem_ld16(lane_base + 64u * (uint32_t)gb + 16u * (uint32_t)half, w);
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = w[e];
            }
            if (live) {
#pragma unroll
                for (int jv = 0; jv < 8; ++jv)
                    if (4 * jv < cw)
                        dst[(size_t)jv * 128] = make_float4(__uint_as_float(v[4 * jv]), __uint_as_float(v[4 * jv + 1]), __uint_as_float(v[4 * jv + 2]), __uint_as_float(v[4 * jv + 3]));
            }
            __syncwarp();                              // tcgen05.ld is .sync.aligned: the warp reconverges before the next one
        }
    } else if (warp == WARP_W) {
        // =============================== TMA warp 1: raw W boxes, one per (chunk, block) ===============================
        const uint64_t pol_keep = l2_evict_last();     // kernel B reads the same rows again as the addend: keep them in L2
        int ci = 0, g = 0;
        for (int i = 0; i < c.n_items; ++i) {
            const int r = i % NRAW;
            mbar_wait(bar_raw_empty(r), (uint32_t)(((i / NRAW) & 1) ^ 1));
            __syncwarp();
            if (elect_one()) {
                tr(0, i, 0);
                mbar_arrive_expect_tx(bar_raw_full(r), (uint32_t)c.sl[g].h * 128u);
                tma_load_2d_hint(raw_st(r), &wmaps.in[c.sl[g].layer], bar_raw_full(r), col_of(c, ci), c.sl[g].row0, pol_keep);
            }
            __syncwarp();
            if (++g == c.n_act) { g = 0; ++ci; }
        }
    } else if (warp == WARP_B) {
        // =============================== TMA warp 2: E_hi | E_lo tiles of the slice ===============================
        for (int ci = 0; ci < c.ncs; ++ci) {
            const int s = ci % NE;
            mbar_wait(bar_e_empty(s), (uint32_t)(((ci / NE) & 1) ^ 1));
            __syncwarp();
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_e_full(s), (uint32_t)e_stage);
                tma_load_2d(e_hi_st(s), &emaps.hi, bar_e_full(s), col_of(c, ci), 0);
                tma_load_2d(e_lo_st(s), &emaps.lo, bar_e_full(s), col_of(c, ci), 0);
            }
            __syncwarp();
        }
    } else if (warp == WARP_MMA) {
        // =============================== the MMA issuer (one warp, converged; one elected lane issues) ===============================
        const uint32_t idesc = idesc_tf32(128, R);
        int i = 0;
        for (int ci = 0; ci < c.ncs; ++ci) {
            const int se = ci % NE;
            mbar_wait(bar_e_full(se), (uint32_t)((ci / NE) & 1));
            for (int g = 0; g < c.n_act; ++g, ++i) {
                const int a = i % NA;
                mbar_wait(bar_a_full(a), (uint32_t)((i / NA) & 1));
                fence_after();
                __syncwarp();
                if (elect_one()) {
                    tr(3, i, 0);
                    const uint64_t b_hi = umma_desc_sw128(e_hi_st(se)), b_lo = umma_desc_sw128(e_lo_st(se));
                    const uint32_t a_hi = tmem_base + A_COL0 + 64u * (uint32_t)a, a_lo = a_hi + 32u;
                    const uint32_t d_tmem = tmem_base + 64u * (uint32_t)g;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {      // 8 tf32 per step: 8 TMEM columns of A, 32 bytes inside the swizzle atom of B
                        const uint64_t adv = (uint64_t)(k * 2);
                        umma_tf32_ts(d_tmem, a_hi + 8u * k, b_hi + adv, idesc, (ci | k) != 0);
                        umma_tf32_ts(d_tmem, a_hi + 8u * k, b_lo + adv, idesc, 1);
                        umma_tf32_ts(d_tmem, a_lo + 8u * k, b_hi + adv, idesc, 1);
                    }
                    umma_commit(bar_a_empty(a));
                    if (g == c.n_act - 1) umma_commit(bar_e_empty(se));      // every block has consumed this E stage
                    if (g == c.n_act - 1 && ci == c.ncs - 1) umma_commit(bar_p_full);
                    tr(3, i, 1);
                }
                __syncwarp();
            }
        }
    }
    fence_before();
    __syncthreads();
    if (threadIdx.x == 0) tr(6, 2, 2);
    if (warp == WARP_MMA) {
        fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ kernel B
__global__ void __launch_bounds__(THREADS, 1)
apply_w_kernel(const Slot* __restrict__ slots, int n_slots, int K, int R, int ks, const float* __restrict__ Pbuf,
               const __grid_constant__ EMaps qmaps, const __grid_constant__ WIo wmaps, long long* __restrict__ trace) {
    auto tr = [&](int role, int idx, int ev) {
        if (trace && blockIdx.x == 0 && idx < 64) trace[(role * 64 + idx) * 4 + ev] = clock64();
    };
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    pdl_launch();
    if (threadIdx.x == 0) tr(6, 2, 0);
    const uint32_t base = smem_u32(smem_raw);
    if (base & 1023u) {
        if (threadIdx.x == 0) printf("uce apply_w: dynamic shared memory base %u is not 1024-byte aligned\n", base);
        __trap();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Cta c = cta_setup(slots, n_slots, K, ks);
    const int n_rc = R / 32;
    const uint32_t bars = base + (uint32_t)((NBX + NQ) * 16384);
    auto bar_box_full  = [&](int b) { return bars + 8u * b; };                 // [0,10)  addend TMA -> epilogue set
    auto bar_box_ready = [&](int b) { return bars + 8u * (10 + b); };          // [10,20) epilogue set -> W TMA warp (box holds W_new)
    auto bar_q_full    = [&](int t) { return bars + 8u * (20 + t); };          // [20,23) Qt TMA -> MMA issuers (all tiles of a unit)
    auto bar_q_empty   = [&](int t) { return bars + 8u * (23 + t); };          // [23,26) one arrival per block of the unit
    auto bar_acc_full  = [&](int a) { return bars + 8u * (26 + a); };          // [26,30) MMA issuer -> epilogue set
    auto bar_acc_empty = [&](int a) { return bars + 8u * (30 + a); };          // [30,34)
    const uint32_t bar_p_ready = bars + 8u * 34, tmem_slot = bars + 8u * 35;
    auto box_st = [&](int b) { return base + (uint32_t)(b * 16384); };
    auto qt_tile = [&](int t, int i) { return base + (uint32_t)(NBX * 16384 + t * 16384 + i * 4096); };     // tile i = 2 * rc + (0 hi, 1 lo)

    if (threadIdx.x == 0) {
        for (int b = 0; b < NBX; ++b) { mbar_init(bar_box_full(b), 1); mbar_init(bar_box_ready(b), 4); }
        for (int t = 0; t < NQ; ++t) { mbar_init(bar_q_full(t), 1); mbar_init(bar_q_empty(t), (uint32_t)c.n_act); }
        for (int a = 0; a < NACC; ++a) { mbar_init(bar_acc_full(a), 1); mbar_init(bar_acc_empty(a), 4); }
        mbar_init(bar_p_ready, TW);
        mbar_fence_init();
    }
    if (warp == WARP_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
    if (warp == WARP_B && lane == 0) {
        tma_prefetch_desc(&qmaps.hi); tma_prefetch_desc(&qmaps.lo);
        for (int g = 0; g < c.n_act; ++g) { tma_prefetch_desc(&wmaps.in[c.sl[g].layer]); tma_prefetch_desc(&wmaps.out[c.sl[g].layer]); }
    }
    fence_before();
    __syncthreads();
    fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (threadIdx.x == 0) tr(6, 2, 1);

    if (warp < TW) {
        // =============================== P = sum of the slices' partials -> hi | lo in tensor memory; then the epilogue ===============================
        const int set = warp >> 2, wq = warp & 3;
        const int trow = 32 * wq + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(32 * wq) << 16);
        const uint32_t row_off = (uint32_t)(trow * 128);
        const uint32_t sw = (uint32_t)(trow & 7);
        // every warp on every block: a set takes half of the rank columns (32 at R = 64, 16 at R = 32); both slices' partials of a
        // block are requested before the first add (scratch layout: see kernel A — coalesced 16-byte loads, lane = row)
        const int cw = R / 2;
        for (int gb = 0; gb < c.n_act; ++gb) {
            const bool live = trow < c.sl[gb].rows;
            float acc[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) acc[e] = 0.f;
            if (live) {
                const size_t part = (size_t)n_slots * (R / 4) * 128;      // float4 elements between two slices' partials
                const float4* s0 = reinterpret_cast<const float4*>(Pbuf) + ((size_t)(c.q * NBLK + gb) * (R / 4) + (size_t)(set * cw / 4)) * 128 + trow;
                for (int p = 0; p < ks; p += 2) {         // fixed order: bit-reproducible; two slices' loads in flight together
                    const bool two = p + 1 < ks;
                    float4 v[8], w[8];
#pragma unroll
                    for (int jv = 0; jv < 8; ++jv)
                        if (4 * jv < cw) {
                            v[jv] = s0[(size_t)p * part + (size_t)jv * 128];
                            w[jv] = two ? s0[(size_t)(p + 1) * part + (size_t)jv * 128] : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                    for (int jv = 0; jv < 8; ++jv)
                        if (4 * jv < cw) {
                            acc[4 * jv] += v[jv].x; acc[4 * jv + 1] += v[jv].y; acc[4 * jv + 2] += v[jv].z; acc[4 * jv + 3] += v[jv].w;
                            if (two) { acc[4 * jv] += w[jv].x; acc[4 * jv + 1] += w[jv].y; acc[4 * jv + 2] += w[jv].z; acc[4 * jv + 3] += w[jv].w; }
                        }
                }
            }
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) tf32_split(acc[e], hi[e], lo[e]);
            __syncwarp();                              // tcgen05.st is .sync.aligned
            const uint32_t t_hi = lane_base + 64u * (uint32_t)gb + (uint32_t)(set * cw), t_lo = t_hi + PLO_COL0;
            if (cw == 32) {
                tmem_st32(t_hi, hi);
                tmem_st32(t_lo, lo);
            } else {
                uint32_t h16[16], l16[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) { h16[e] = hi[e]; l16[e] = lo[e]; }
                tmem_st16(t_hi, h16);
                tmem_st16(t_lo, l16);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p_ready);
        if (threadIdx.x == 0) tr(6, 0, 1);
        // ---- epilogue: per (unit, block) box += accumulator (in place, swizzled smem); the W TMA warp stores the box ----
        int g = set;
        while (g >= c.n_act) g -= c.n_act;
        for (int i = set; i < c.n_items; i += 2) {
            const int b = i % NBX, a = i % NACC;
            mbar_wait(bar_acc_full(a), (uint32_t)((i / NACC) & 1));
            if (threadIdx.x == 0) tr(5, i, 0);
            fence_after();
            uint32_t v[32];
            tmem_ld32(lane_base + ACC_COL0 + 32u * (uint32_t)a, v);
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty(a));
            mbar_wait(bar_box_full(b), (uint32_t)((i / NBX) & 1));
            if (threadIdx.x == 0) tr(5, i, 1);
            if (trow < c.sl[g].rows) {
                const uint32_t row = box_st(b) + row_off;
#pragma unroll
                for (int jv = 0; jv < 8; ++jv) {
                    const uint32_t addr = row + (((uint32_t)jv ^ sw) << 4);
                    const float4 w = lds_v4(addr);
                    sts_v4(addr, w.x + __uint_as_float(v[4 * jv]), w.y + __uint_as_float(v[4 * jv + 1]),
                           w.z + __uint_as_float(v[4 * jv + 2]), w.w + __uint_as_float(v[4 * jv + 3]));
                }
            }
            fence_proxy_async();                       // generic-proxy writes -> visible to the TMA store
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_box_ready(b));
            if (threadIdx.x == 0) tr(5, i, 2);
            g += 2;
            while (g >= c.n_act) g -= c.n_act;
        }
    } else if (warp == WARP_W) {
        // =============================== TMA warp 1: addend boxes in, W_new boxes out ===============================
        const uint64_t pol_in = l2_evict_first(), pol_out = l2_evict_first();
        auto load_box = [&](int item) {
            const int u = item / c.n_act, g = item - u * c.n_act, bx = item % NBX;
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_box_full(bx), (uint32_t)c.sl[g].h * 128u);
                tma_load_2d_hint(box_st(bx), &wmaps.in[c.sl[g].layer], bar_box_full(bx), col_of(c, u), c.sl[g].row0, pol_in);
            }
            __syncwarp();
        };
        for (int i = 0; i < NBX && i < c.n_items; ++i) load_box(i);
        // item i: W_new is complete -> TMA store; once the PREVIOUS item's store has been read out of shared memory its box takes the
        // addend of item i - 1 + NBX.  (One thread issues every store: bulk async-groups are per thread.)
        int u = 0, g = 0;
        for (int i = 0; i < c.n_items; ++i) {
            const int b = i % NBX;
            mbar_wait(bar_box_ready(b), (uint32_t)((i / NBX) & 1));
            __syncwarp();
            const int ni = i - 1 + NBX;
            const bool reload = i >= 1 && ni < c.n_items;
            if (elect_one()) {
                tr(0, i, 0);
                tma_store_2d(&wmaps.out[c.sl[g].layer], box_st(b), col_of(c, u), c.sl[g].row0, pol_out);
                tma_store_commit();
                if (reload) tma_store_wait_read<1>();   // the previous item's box has been read out
            }
            __syncwarp();
            if (reload) load_box(ni);                   // ni % NBX == (i - 1) % NBX
            if (++g == c.n_act) { g = 0; ++u; }
        }
        __syncwarp();
        if (elect_one()) tma_store_wait_read<0>();     // shared memory must outlive the last store's read
        __syncwarp();
    } else if (warp == WARP_B) {
        // =============================== TMA warp 2: Qt_hi | Qt_lo tiles of a unit under ONE barrier ===============================
        // Programmatic dependent launch: this kernel is launched while the factor's last kernel (which writes Qt) still runs; everything
        // above and beside this warp — barriers, tensor memory, the partial products (kernel A finished before the factor's solve was
        // even launched) and the addend boxes of W_old — needs nothing from it.  Only the Qt tiles do: wait here.
        pdl_wait();
        const uint32_t q_bytes = (uint32_t)(n_rc * 2) * 4096u;
        for (int u = 0; u < c.ncs; ++u) {
            const int t = u % NQ;
            mbar_wait(bar_q_empty(t), (uint32_t)(((u / NQ) & 1) ^ 1));
            __syncwarp();
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_q_full(t), q_bytes);
                for (int rc = 0; rc < n_rc; ++rc) {
                    tma_load_2d(qt_tile(t, 2 * rc), &qmaps.hi, bar_q_full(t), rc * 32, col_of(c, u));
                    tma_load_2d(qt_tile(t, 2 * rc + 1), &qmaps.lo, bar_q_full(t), rc * 32, col_of(c, u));
                }
            }
        }
    } else {
        // =============================== MMA issuers: D[128 rows, 32 cols] = P_hi Qt_hi^T + P_lo Qt_hi^T + P_hi Qt_lo^T ===============================
        // An accumulator belongs to one work item and all MMAs of an item come from one thread (issuer = item parity; items on the
        // same accumulator slot are NACC = 4 apart: the same parity, the same issuer, the same epilogue set).  Both issuers wait
        // for EVERY unit's Qt barrier, also for a unit in which they have no item (a CTA with one block): see "barrier discipline".
        const int t = warp - WARP_MMA;
        const uint32_t idesc = idesc_tf32(128, 32);
        mbar_wait(bar_p_ready, 0);
        fence_after();
        int i = 0;
        for (int u = 0; u < c.ncs; ++u) {
            const int tq = u % NQ;
            mbar_wait(bar_q_full(tq), (uint32_t)((u / NQ) & 1));
            for (int g = 0; g < c.n_act; ++g, ++i) {
                if ((i & 1) != t) continue;
                const int a = i % NACC;
                mbar_wait(bar_acc_empty(a), (uint32_t)(((i / NACC) & 1) ^ 1));
                fence_after();
                __syncwarp();
                if (elect_one()) {
                    tr(4, i, 0);
                    const uint32_t d_tmem = tmem_base + ACC_COL0 + 32u * (uint32_t)a;
                    const uint32_t p_hi = tmem_base + 64u * (uint32_t)g, p_lo = tmem_base + PLO_COL0 + 64u * (uint32_t)g;
                    for (int rc = 0; rc < n_rc; ++rc) {
                        const uint64_t bq_hi = umma_desc_sw128(qt_tile(tq, 2 * rc)), bq_lo = umma_desc_sw128(qt_tile(tq, 2 * rc + 1));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adv = (uint64_t)(k * 2);
                            const uint32_t col = (uint32_t)(rc * 32 + 8 * k);
                            umma_tf32_ts(d_tmem, p_hi + col, bq_hi + adv, idesc, (rc | k) != 0);      // hi.hi
                            umma_tf32_ts(d_tmem, p_lo + col, bq_hi + adv, idesc, 1);                  // lo.hi
                            umma_tf32_ts(d_tmem, p_hi + col, bq_lo + adv, idesc, 1);                  // hi.lo
                        }
                    }
                    umma_commit(bar_q_empty(tq));
                    umma_commit(bar_acc_full(a));
                    tr(4, i, 1);
                }
                __syncwarp();
            }
        }
    }
    fence_before();
    __syncthreads();
    if (threadIdx.x == 0) tr(6, 2, 2);
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

}  // namespace ab

// Number of K slices: a slice has at most 512 columns (accumulation chain, see the header) and at least 4 chunks of 32.
int apply_ab_ksplit(int K) {
    if (const char* e = getenv("UCE_AB_KSPLIT")) {
        const int t = atoi(e);
        if (t >= 1 && t <= 8 && K % (32 * t) == 0) return t;
    }
    int ks = 1;
    while (ks < 8 && K % (64 * ks) == 0 && (K / ks > 512 || ks < 2) && K / (2 * ks) >= 128) ks *= 2;
    return ks;
}

bool apply_ab_available(const uce_ws* ws, int n_layers) {
    const int R = ws->rank_pad;
    return ws->K % 32 == 0 && ws->K >= 128 && (R == 32 || R == 64) && !ws->dense && ws->rank > 0 && n_layers <= ab::MAX_LAYERS &&
           tensor_map_encoder() != nullptr;
}

// Row-block plan: block height per projection (multiple of 8, <= 128) and the number of blocks.  When the whole edit fits one
// wave — NBLK blocks on each of sm_count / ks CTA groups — the height is the SMALLEST one >= 64 that still fits (the most even
// spread of rows over the SMs), evened out inside each projection; otherwise blocks are 128 rows.  UCE_AB_BLOCK_ROWS overrides.
int apply_ab_plan(int sm_count, int ks, const int* d, int n_layers, int* block_rows, int* first_block) {
    auto count = [&](int H) { long t = 0; for (int l = 0; l < n_layers; ++l) t += ceil_div(d[l], H); return t; };
    const int sms = sm_count > 0 ? sm_count : 148;
    const long cap = (long)ab::NBLK * (sms / ks > 0 ? sms / ks : 1);
    int H = 128, forced = 0;
    if (const char* e = getenv("UCE_AB_BLOCK_ROWS")) {
        const int t = atoi(e);
        if (t >= 8 && t <= 128 && t % 8 == 0) { H = t; forced = 1; }
    }
    if (!forced && count(128) <= cap)
        for (int t = 64; t <= 128; t += 8)
            if (count(t) <= cap) { H = t; break; }
    int blocks = 0;
    for (int l = 0; l < n_layers; ++l) {
        const int nb = ceil_div(d[l], H);
        int hl = forced ? H : round_up(ceil_div(d[l], nb), 8);
        if (hl > 128) hl = 128;
        block_rows[l] = hl;
        first_block[l] = blocks;
        blocks += ceil_div(d[l], hl);
    }
    return blocks;
}

// stage: 0 both kernels on `st`; 1 only kernel A (partial products; needs E only); 2 only kernel B (needs Q and the partials).
// `slots_dev` holds `n_slots` planned blocks (host copy in `slots_host`); `layers_host` gives the weights of each projection.
int apply_ab_lowrank(uce_ws* ws, const void* slots_dev, const void* slots_host, int n_slots, const LayerRef* layers_host, int n_layers,
                     cudaStream_t st, int stage, int* launches, cudaEvent_t ev_mid) {
    using namespace ab;
    const int K = ws->K, R = ws->rank_pad;
    if (!apply_ab_available(ws, n_layers)) { set_error("K-split tcgen05 apply unavailable for K=%d rank_pad=%d dense=%d layers=%d", K, R, ws->dense, n_layers); return UCE_E_STATE; }
    const int ks = apply_ab_ksplit(K);
    for (int l = 0; l < n_layers; ++l)
        if (((uintptr_t)layers_host[l].w_old & 15) || ((uintptr_t)layers_host[l].w_new & 15)) { set_error("tcgen05 apply needs 16-byte aligned weights"); return UCE_E_ARG; }
    (void)slots_host;
    const size_t need = (size_t)ks * n_slots * 128 * R;
    if (ws->P_off + need > ws->P_cap) { set_error("apply scratch too small (%zu > %zu floats)", ws->P_off + need, ws->P_cap); return UCE_E_STATE; }
    float* Pscr = ws->P + ws->P_off;
    static thread_local WIo wmaps;        // kept off the stack, one per host thread; copied into the launches by value
    int rc;
    for (int l = 0; l < n_layers; ++l) {
        if ((rc = make_map(&wmaps.in[l], layers_host[l].w_old, layers_host[l].d, K, layers_host[l].tile_rows))) return rc;
        if ((rc = make_map(&wmaps.out[l], layers_host[l].w_new, layers_host[l].d, K, layers_host[l].tile_rows))) return rc;
    }
    static thread_local int conf_a[64] = {0}, conf_b[64] = {0};      // opt-in shared-memory size is a per-device function attribute
    const int grid = ceil_div(n_slots, NBLK) * ks;
    long long* trace = nullptr;                                       // debugging aid: timeline of CTA 0 of both kernels (synchronises)
    const char* trace_path = getenv("UCE_AB_TRACE");
    constexpr int TRN = 2 * 7 * 64 * 4;
    if (trace_path && stage == 0) { UCE_CUDA(cudaMalloc(&trace, TRN * sizeof(long long))); UCE_CUDA(cudaMemsetAsync(trace, 0, TRN * sizeof(long long), st)); }
    if (stage == 0 || stage == 1) {
        EMaps em;
        if ((rc = make_map(&em.hi, ws->E_hi, R, K, R))) return rc;
        if ((rc = make_map(&em.lo, ws->E_lo, R, K, R))) return rc;
        const int smem = smem_a(R);
        int& cf = conf_a[ws->device & 63];
        if (cf < smem) { UCE_CUDA(cudaFuncSetAttribute(apply_p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); cf = smem; }
        apply_p_kernel<<<grid, THREADS_A, smem, st>>>((const Slot*)slots_dev, n_slots, K, R, ks, Pscr, em, *reinterpret_cast<const WIn*>(&wmaps), trace);
        UCE_LAUNCH_CHECK();
        *launches += 1;
    }
    if (stage == 0 && ev_mid) UCE_CUDA(cudaEventRecord(ev_mid, st));
    if (stage == 0 || stage == 2) {
        EMaps qm;
        if ((rc = make_map(&qm.hi, ws->Qt_hi, K, R, 32))) return rc;
        if ((rc = make_map(&qm.lo, ws->Qt_lo, K, R, 32))) return rc;
        const int smem = smem_b();
        int& cf = conf_b[ws->device & 63];
        if (cf < smem) { UCE_CUDA(cudaFuncSetAttribute(apply_w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); cf = smem; }
        UCE_CUDA(launch_k(apply_w_kernel, dim3(grid), dim3(THREADS), (size_t)smem, st, 1, (const Slot*)slots_dev, n_slots, K, R, ks, (const float*)Pscr, qm, wmaps,
                          trace ? trace + TRN / 2 : (long long*)nullptr));
        *launches += 1;
    }
    if (trace) {
        std::vector<long long> hbuf(TRN);
        UCE_CUDA(cudaStreamSynchronize(st));
        UCE_CUDA(cudaMemcpy(hbuf.data(), trace, hbuf.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        UCE_CUDA(cudaFree(trace));
        if (FILE* f = fopen(trace_path, "w")) {
            const char* roles[7] = {"w_tma", "transform", "e_tma", "mma_a", "mma_b", "epilogue", "p"};
            for (int kz = 0; kz < 2; ++kz) {
                long long t0 = 0;
                for (int x = 0; x < TRN / 2; ++x) { const long long v = hbuf[kz * TRN / 2 + x]; if (v && (!t0 || v < t0)) t0 = v; }
                for (int r = 0; r < 7; ++r) for (int i = 0; i < 64; ++i) {
                    const long long* e = &hbuf[kz * TRN / 2 + (r * 64 + i) * 4];
                    if (e[0] || e[1] || e[2]) fprintf(f, "%c %s %d %lld %lld %lld\n", kz ? 'B' : 'A', roles[r], i, e[0] ? e[0] - t0 : -1, e[1] ? e[1] - t0 : -1, e[2] ? e[2] - t0 : -1);
                }
            }
            fclose(f);
        }
    }
    return 0;
}

size_t apply_ab_slot_bytes() { return sizeof(ab::Slot); }
void apply_ab_fill_slots(void* slots_host, const LayerRef* layers_host, int n_layers) {
    ab::Slot* s = (ab::Slot*)slots_host;
    int n = 0;
    for (int l = 0; l < n_layers; ++l)
        for (int r0 = 0; r0 < layers_host[l].d; r0 += layers_host[l].tile_rows) {
            const int rows = layers_host[l].d - r0 < layers_host[l].tile_rows ? layers_host[l].d - r0 : layers_host[l].tile_rows;
            s[n++] = ab::Slot{l, r0, rows, layers_host[l].tile_rows};
        }
}

}  // namespace uce
