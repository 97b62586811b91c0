// Fused low-rank apply (rank pad <= 64): ONE CTA per SM that carries TWO row blocks through a single E / Qt stream.
//
//   W_new[rows,:] = W_old[rows,:] + (W_old[rows,:] E^T) Q            (uce_sd_erase.py:45-82, see apply.cu)
//
// fp32 fidelity through 3xTF32 (hi.hi + hi.lo + lo.hi, lo.lo dropped).  Why this shape (two earlier ones — one 128-row tile per
// CTA, then two co-resident CTAs per SM — were measured and retired, profiles/r01_apply_history.txt):
// their timelines and ncu captures show that an SM moves ~35-45 B/clk through TMA whatever
// the ring depths are, and that a row tile of 128 rows costs 1.57 MB of TMA ingest: its W rows twice (phase A, then the
// addend of phase B; 2 x 393 KB) plus the WHOLE of E_hi|E_lo and Qt_hi|Qt_lo (2 x 393 KB) — the low-rank operands are
// as large as the tile.  Two co-resident 128-row CTAs therefore stream E and Qt twice per SM, and with 200 tiles on
// 148 SMs a third of the SMs carry twice the bytes of the others.  Here:
//   * a CTA owns TWO row blocks of `h` rows (h <= 128, chosen per projection by the host so that ALL CTAs of an edit
//     form one balanced wave: SD-1.4 -> 144 CTAs of 160-192 rows); both blocks share every E / Qt tile in shared memory:
//     ingest per SM falls from 3.14 MB to ~1.9 MB;
//   * 512 TMEM columns and 225 KB of shared memory belong to one CTA: 5-deep rings of block pairs.
//
//   phase A  P_g[128,R] = W_g[128,K] . E[R,K]^T for both blocks g, TMEM columns [64g, 64g + R)
//            * W TMA warp: per 32-column chunk one [h x 32] box per block (L2 evict_last), 5-deep ring
//            * 8 transform warps (group g = warp / 4 works on block g; thread = row = TMEM lane): hi / lo split (integer
//              round + mask + one subtraction: no conversion-pipe instruction), tcgen05.st into A stage (s, g)
//            * E TMA warp: [R x 32] tiles of E_hi, E_lo, 3-deep ring
//            * MMA warp (converged, one elected lane issues): per block and 8-wide k-step hi.hi + hi.lo + lo.hi (N = R, A from TMEM)
//   phase B  dW_g[128 rows, 32 cols] = P_g . Q[:, 32 cols] per unit of 32 W columns
//            * P_g stays in tensor memory: hi in place, lo in freed A-stage columns; A operand of the phase-B MMAs
//            * Qt_hi / Qt_lo [32 x 32] tiles of a unit under ONE barrier, 2 slots
//            * per unit and block a [h x 32] box: W_old addend in by TMA, accumulator added in place (swizzled smem, lane =
//              row), W_new out by TMA store; the W TMA warp issues stores and reloads (6-deep ring of box pairs)
#include "tc_apply_common.cuh"
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace uce {
namespace tc3 {
using namespace uce::tc;
using namespace uce::tca;

constexpr int NBLK = 2;                                 // row blocks per CTA
constexpr int PW = 4 * NBLK;                            // transform / P-conversion / epilogue warps (4 per block)
constexpr int THREADS = (PW + 2 + NBLK) * 32;           // + W TMA warp + E/Qt TMA warp + one MMA warp per block
constexpr int NRAW = 5, NSA = 3, NE = 3;
constexpr int NB = 6;                                   // addend / output box pairs (2 x 16 KB each)
constexpr int NACC = 4;                                 // accumulator pairs (2 x 32 columns)
constexpr int NQ = 2;                                   // Qt slots (all hi/lo tiles of one unit, <= 16 KB)
constexpr int MAX_LAYERS = 96;                          // two tensor maps per projection travel as kernel parameters
constexpr int WARP_W_TMA = PW, WARP_E_TMA = PW + 1, WARP_MMA = PW + 2;      // MMA warps: WARP_MMA + g issues block g
constexpr uint32_t TMEM_COLS = 512;
// tensor-memory columns
__host__ __device__ constexpr uint32_t col_p(int g) { return 64u * g; }                       // phase A accumulator / phase B P_hi
__host__ __device__ constexpr uint32_t col_stage(int s, int g) { return 128u + 128u * s + 64u * g; }   // {W_hi 32, W_lo 32}
__host__ __device__ constexpr uint32_t col_plo(int g) { return 128u + 64u * g; }              // phase B P_lo (freed A stage 0)
__host__ __device__ constexpr uint32_t col_acc(int a, int g) { return 256u + 64u * a + 32u * g; }

struct Maps { CUtensorMap e_hi, e_lo, qt_hi, qt_lo; };
struct WMaps { CUtensorMap in[MAX_LAYERS], out[MAX_LAYERS]; };

// Shared-memory carve-up (bytes), identical on host and device; the dynamic window starts 1024-byte aligned (checked).
//   phase A   [0, 160K) raw ring: NRAW stages x NBLK slots of 16 KB     [160K, + NE * e_stage) E ring {E_hi | E_lo}
//   phase B   [0, 192K) NB box pairs of 2 x 16 KB                        [192K, + NQ * 16K) Qt slots     (aliases phase A)
struct Smem { int e_stage, e_off, q_off, bar_off, total; };
__host__ __device__ inline Smem smem_layout(int R) {
    Smem s;
    s.e_stage = 2 * R * 128;
    s.e_off = NRAW * NBLK * 16384;
    s.q_off = NB * NBLK * 16384;
    const int end_a = s.e_off + NE * s.e_stage;
    const int end_b = s.q_off + NQ * 16384;
    s.bar_off = end_a > end_b ? end_a : end_b;
    s.total = s.bar_off + 512;
    return s;
}

__device__ __forceinline__ int find_layer(const LayerRef* layers, int n_layers, int tile) {
    int lo = 0, hi = n_layers - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (layers[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(THREADS, 1)
apply_tc3_kernel(const LayerRef* __restrict__ layers, int n_layers, int K, int R, int addend_ldgsts,
                 const __grid_constant__ Maps maps, const __grid_constant__ WMaps wmaps, long long* __restrict__ trace) {
    // optional timeline of CTA 0 (UCE_TC_TRACE=<file>): trace[(role * 64 + index) * 4 + event] = clock64()
    auto tr = [&](int role, int idx, int ev) {
        if (trace && blockIdx.x == 0 && idx < 64) trace[(role * 64 + idx) * 4 + ev] = clock64();
    };
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);                          // swizzled tiles need 1024-byte alignment
    if (base & 1023u) {
        if (threadIdx.x == 0) printf("uce apply_tc3: dynamic shared memory base %u is not 1024-byte aligned\n", base);
        __trap();
    }
    const Smem L = smem_layout(R);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- barriers ----
    const uint32_t bars = base + L.bar_off;
    auto bar_raw_full  = [&](int r) { return bars + 8u * r; };                 // [0,5)   W TMA -> transform warps
    auto bar_raw_empty = [&](int r) { return bars + 8u * (5 + r); };           // [5,10)
    auto bar_a_full    = [&](int s) { return bars + 8u * (10 + s); };          // [10,13) transform warps -> MMA (A stages of both blocks written)
    auto bar_a_empty   = [&](int s) { return bars + 8u * (13 + s); };          // [13,16) MMA -> transform warps
    auto bar_e_full    = [&](int s) { return bars + 8u * (16 + s); };          // [16,19) E TMA -> MMA
    auto bar_e_empty   = [&](int s) { return bars + 8u * (19 + s); };          // [19,22)
    const uint32_t bar_p_full = bars + 8u * 22, bar_p_ready = bars + 8u * 23;
    auto bar_q_full    = [&](int t) { return bars + 8u * (24 + t); };          // [24,26) Qt TMA -> MMA (all tiles of a unit)
    auto bar_q_empty   = [&](int t) { return bars + 8u * (26 + t); };          // [26,28)
    auto bar_acc_full  = [&](int a, int g) { return bars + 8u * (28 + 2 * a + g); };     // [28,36) MMA warp g -> epilogue group g
    auto bar_acc_empty = [&](int a, int g) { return bars + 8u * (36 + 2 * a + g); };     // [36,44)
    auto bar_box_full  = [&](int b) { return bars + 8u * (44 + b); };          // [44,50) addend TMA -> epilogue
    auto bar_box_ready = [&](int b) { return bars + 8u * (50 + b); };          // [50,56) epilogue -> W TMA warp (boxes hold W_new)
    const uint32_t tmem_slot = bars + 8u * 56;

    const int tile = blockIdx.x;
    const int layer = find_layer(layers, n_layers, tile);
    const LayerRef Lr = layers[layer];
    const int h = Lr.tile_rows;                                  // rows per block (multiple of 8, <= 128)
    const int row0 = (tile - Lr.tile_begin) * NBLK * h;
    const int rows_valid0 = max(0, min(h, Lr.d - row0)), rows_valid1 = max(0, min(h, Lr.d - (row0 + h)));
    const int n_act = rows_valid1 > 0 ? 2 : 1;                   // active blocks (block 0 always has rows)
    const int n_chunks = K / 32;          // phase A k-chunks (32 fp32 = one swizzle atom row) == phase B units of 32 W columns
    const int n_rc = R / 32;              // r atoms
    const uint32_t box_bytes = (uint32_t)h * 128u;
    // Every CTA streams the SAME E and Qt tiles; started together they would all ask the same few L2 slices for the same
    // lines at the same time.  CTA i walks the K dimension (phase A chunks, phase B units) starting at a different offset.
    const int rot = (int)((blockIdx.x * 7u) % (unsigned)n_chunks);
    auto col_of = [&](int c) { int x = c + rot; if (x >= n_chunks) x -= n_chunks; return x * 32; };

    if (threadIdx.x == 0) {
        for (int r = 0; r < NRAW; ++r) { mbar_init(bar_raw_full(r), 1); mbar_init(bar_raw_empty(r), PW); }
        // every consumer-release barrier of the MMA side counts BOTH MMA warps (an idle block's warp arrives without MMAs)
        for (int s = 0; s < NSA; ++s) { mbar_init(bar_a_full(s), PW); mbar_init(bar_a_empty(s), NBLK); }
        for (int s = 0; s < NE; ++s) { mbar_init(bar_e_full(s), 1); mbar_init(bar_e_empty(s), NBLK); }
        for (int t = 0; t < NQ; ++t) { mbar_init(bar_q_full(t), 1); mbar_init(bar_q_empty(t), NBLK); }
        mbar_init(bar_p_full, NBLK); mbar_init(bar_p_ready, PW);
        for (int a = 0; a < NACC; ++a)
            for (int g = 0; g < NBLK; ++g) { mbar_init(bar_acc_full(a, g), 1); mbar_init(bar_acc_empty(a, g), 4); }
        // addend boxes: one expect_tx arrival (TMA) or the 32 cp.async completion arrivals of the W TMA warp's lanes
        for (int b = 0; b < NB; ++b) { mbar_init(bar_box_full(b), addend_ldgsts ? 32 : 1); mbar_init(bar_box_ready(b), PW); }
        mbar_fence_init();
    }
    if (warp == WARP_MMA) tmem_alloc(tmem_slot, TMEM_COLS);      // all of the SM's tensor memory: one CTA per SM
    if (warp == WARP_E_TMA && lane == 0) {
        tma_prefetch_desc(&wmaps.in[layer]); tma_prefetch_desc(&wmaps.out[layer]);
        tma_prefetch_desc(&maps.e_hi); tma_prefetch_desc(&maps.e_lo); tma_prefetch_desc(&maps.qt_hi); tma_prefetch_desc(&maps.qt_lo);
    }
    fence_before();
    __syncthreads();
    fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    auto raw_st = [&](int r, int g) { return base + (uint32_t)((r * NBLK + g) * 16384); };
    auto stage_e_hi = [&](int s) { return base + (uint32_t)(L.e_off + s * L.e_stage); };
    auto stage_e_lo = [&](int s) { return base + (uint32_t)(L.e_off + s * L.e_stage + R * 128); };
    auto box_st = [&](int b, int g) { return base + (uint32_t)((b * NBLK + g) * 16384); };
    auto qt_tile = [&](int t, int i) { return base + (uint32_t)(L.q_off + t * 16384 + i * 4096); };     // tile i = 2 * rc + (0 hi, 1 lo)

    if (warp < PW) {
        // =============================== W transform, then P conversion, then epilogue (group g = block g) ===============================
        const int g = warp >> 2, wq = warp & 3;
        const int trow = 32 * wq + lane;                 // row of the block == TMEM lane
        const bool row_live = trow < (g ? rows_valid1 : rows_valid0);
        const uint32_t lane_base = tmem_base + ((uint32_t)(32 * wq) << 16);
        const uint32_t row_off = (uint32_t)(trow * 128);
        const uint32_t sw = (uint32_t)(trow & 7);
        // The stores into tensor memory are waited for (tcgen05.wait::st) BEFORE the registers they read are written again.  A
        // software-pipelined variant that left them in flight across the next chunk's split was measured neutral in time, and the
        // same pattern in apply_gemm3x.cu intermittently garbled one chunk of one warp's rows: the source registers of an
        // asynchronous tcgen05.st are not safe to overwrite before the wait, just as tcgen05.ld results are not safe to read.
        for (int c = 0; c < n_chunks; ++c) {
            const int r = c % NRAW, s = c % NSA;
            mbar_wait(bar_raw_full(r), (uint32_t)((c / NRAW) & 1));
            if (threadIdx.x == 0) tr(1, c, 0);
            const uint32_t raw = raw_st(r, g) + row_off;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row_live) v = lds_v4(raw + (((uint32_t)j ^ sw) << 4));      // swizzled 16-byte slot the TMA wrote
                tf32_split(v.x, hi[4 * j], lo[4 * j]);         tf32_split(v.y, hi[4 * j + 1], lo[4 * j + 1]);
                tf32_split(v.z, hi[4 * j + 2], lo[4 * j + 2]); tf32_split(v.w, hi[4 * j + 3], lo[4 * j + 3]);
            }
            mbar_wait(bar_a_empty(s), (uint32_t)(((c / NSA) & 1) ^ 1));      // the MMAs that read this A stage have completed
            if (threadIdx.x == 0) tr(1, c, 1);
            fence_after();
            const uint32_t ta = lane_base + col_stage(s, g);
            tmem_st32(ta, hi);
            tmem_st32(ta + 32u, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_a_full(s)); mbar_arrive(bar_raw_empty(r)); }
            if (threadIdx.x == 0) tr(1, c, 2);
        }
        // ---- P_g: TMEM -> registers -> hi (in place) | lo (freed A-stage columns): the A operand of phase B ----
        mbar_wait(bar_p_full, 0);
        if (threadIdx.x == 0) tr(6, 0, 0);
        fence_after();
        for (int rc = 0; rc < n_rc; ++rc) {
            uint32_t v[32], lo[32];
            tmem_ld32(lane_base + col_p(g) + (uint32_t)(rc * 32), v);
#pragma unroll
            for (int i = 0; i < 32; ++i) tf32_split(__uint_as_float(v[i]), v[i], lo[i]);
            tmem_st32(lane_base + col_p(g) + (uint32_t)(rc * 32), v);
            tmem_st32(lane_base + col_plo(g) + (uint32_t)(rc * 32), lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");      // before v / lo are written again (next rc)
        }
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p_ready);
        if (threadIdx.x == 0) tr(6, 0, 1);
        // ---- epilogue: per unit of 32 W columns, box += accumulator (in place, swizzled smem); the W TMA warp stores the box ----
        for (int u = 0; u < n_chunks; ++u) {
            const int b = u % NB, a = u % NACC;
            mbar_wait(bar_acc_full(a, g), (uint32_t)((u / NACC) & 1));
            fence_after();
            uint32_t v[32];
            tmem_ld32(lane_base + col_acc(a, g), v);
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty(a, g));
            mbar_wait(bar_box_full(b), (uint32_t)((u / NB) & 1));
            if (threadIdx.x == 0) tr(5, u, 0);
            if (row_live) {
                const uint32_t row = box_st(b, g) + row_off;
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
        // =============================== TMA warp 1: raw W chunks (phase A), addend boxes in / W_new boxes out (phase B) ===============================
        // (all lanes wait, one elected lane issues: see elect_one() in tc_common.cuh)
        const uint64_t pol_keep = l2_evict_last();     // the rows are read again as the addend: keep them in L2
        const uint64_t pol_last_use = l2_evict_first();
        const CUtensorMap* wm = &wmaps.in[layer];
        const uint32_t pair_bytes = box_bytes * (uint32_t)n_act;
        for (int c = 0; c < n_chunks; ++c) {
            const int r = c % NRAW;
            mbar_wait(bar_raw_empty(r), (uint32_t)(((c / NRAW) & 1) ^ 1));
            __syncwarp();
            if (elect_one()) {
                tr(0, c, 0);
                mbar_arrive_expect_tx(bar_raw_full(r), pair_bytes);
                for (int g = 0; g < n_act; ++g) tma_load_2d_hint(raw_st(r, g), wm, bar_raw_full(r), col_of(c), row0 + g * h, pol_keep);
            }
        }
        // Box pair b < NRAW occupies exactly the bytes of raw stage b: it takes the addend of unit b as soon as the transform
        // warps have released that stage for the last time — the first addends arrive while phase A drains and P is converted.
        // Box pairs >= NRAW alias the E ring: they wait until every phase-A MMA has completed (P final).
        const uint64_t pol_stream = l2_evict_first();
        const CUtensorMap* om = &wmaps.out[layer];
        // The addend of unit `unit` into box pair `bx`.  Two routes: TMA (one elected lane), or — addend_ldgsts — 16-byte
        // cp.async copies issued by all 32 lanes straight into the swizzled slots, completion counted on the same mbarrier:
        // that traffic then goes through the LSU / L1 path instead of the SM's TMA unit, which is what bounds this kernel.
        auto load_box = [&](int unit, int bx) {
            if (!addend_ldgsts) {
                if (elect_one()) {
                    mbar_arrive_expect_tx(bar_box_full(bx), pair_bytes);
                    for (int g = 0; g < n_act; ++g) tma_load_2d_hint(box_st(bx, g), wm, bar_box_full(bx), col_of(unit), row0 + g * h, pol_last_use);
                }
            } else {
                const int col = col_of(unit);
                for (int g = 0; g < n_act; ++g) {
                    const int rows = g ? rows_valid1 : rows_valid0;
                    const float* src0 = Lr.w_old + (size_t)(row0 + g * h) * K + col;
                    const uint32_t dst0 = box_st(bx, g);
                    for (int idx = lane; idx < rows * 8; idx += 32) {            // 8 lanes = the 128 bytes of one row: coalesced
                        const int row = idx >> 3, j = idx & 7;
                        cp_async_16(dst0 + (uint32_t)(row * 128 + ((j ^ (row & 7)) << 4)), src0 + (size_t)row * K + 4 * j);
                    }
                }
                cp_async_mbar_arrive(bar_box_full(bx));
            }
            __syncwarp();
        };
        for (int u = 0; u < NB && u < n_chunks; ++u) {
            if (u < NRAW) {
                const int uses = (n_chunks - u + NRAW - 1) / NRAW;            // chunks that went through raw stage u (>= 1)
                mbar_wait(bar_raw_empty(u), (uint32_t)((uses - 1) & 1));
            } else {
                mbar_wait(bar_p_full, 0);
            }
            __syncwarp();
            load_box(u, u);
        }
        __syncwarp();
        // unit u: W_new is complete -> TMA stores; once the PREVIOUS unit's stores have been read out of shared memory its
        // boxes take the addend of unit u - 1 + NB.  (One thread issues every store: bulk async-groups are per thread.)
        for (int u = 0; u < n_chunks; ++u) {
            const int b = u % NB;
            mbar_wait(bar_box_ready(b), (uint32_t)((u / NB) & 1));
            __syncwarp();
            const int nu = u - 1 + NB;
            const bool reload = u >= 1 && nu < n_chunks;
            if (elect_one()) {
                tr(0, u, 1);
                for (int g = 0; g < n_act; ++g) tma_store_2d(om, box_st(b, g), col_of(u), row0 + g * h, pol_stream);
                tma_store_commit();
                if (reload) tma_store_wait_read<1>();   // the previous unit's boxes have been read out
            }
            __syncwarp();
            if (reload) load_box(nu, nu % NB);          // nu % NB == (u - 1) % NB
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
                tma_load_2d(stage_e_hi(s), &maps.e_hi, bar_e_full(s), col_of(c), 0);
                tma_load_2d(stage_e_lo(s), &maps.e_lo, bar_e_full(s), col_of(c), 0);
            }
        }
        // Qt slot 1 lies beyond the E ring (nothing of phase A lives there): unit 0 goes there and is fetched right away;
        // slot 0 aliases the last E stage and waits until every phase-A MMA has completed.
        const uint32_t q_bytes = (uint32_t)(n_rc * 2) * 4096u;
        for (int u = 0; u < n_chunks; ++u) {
            const int t = (u + 1) % NQ;
            if (u == 1) mbar_wait(bar_p_full, 0);
            mbar_wait(bar_q_empty(t), (uint32_t)(((u / NQ) & 1) ^ 1));
            __syncwarp();
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_q_full(t), q_bytes);        // ONE barrier for all hi/lo tiles of the unit
                for (int rc = 0; rc < n_rc; ++rc) {
                    tma_load_2d(qt_tile(t, 2 * rc), &maps.qt_hi, bar_q_full(t), rc * 32, col_of(u));
                    tma_load_2d(qt_tile(t, 2 * rc + 1), &maps.qt_lo, bar_q_full(t), rc * 32, col_of(u));
                }
            }
        }
    } else {
        // =============================== MMA issuers: warp WARP_MMA + g owns row block g ===============================
        // Each warp runs its loops and barrier waits converged; ONE elected lane issues its block's MMAs and commits.  Two
        // issuers because one could not keep the tensor pipe busy: it needs ~400 cycles of barrier round trips between
        // chunks during which its queue runs dry (profiles/r01_apply_tc3_timeline.txt: 1 300 cycles per chunk for 770 cycles
        // of MMAs); with two, one block's MMAs execute while the other block's issuer waits.
        const int g = warp - WARP_MMA;
        const bool act = g < n_act;
        const uint32_t idesc_a = idesc_tf32(128, R);
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % NSA, se = c % NE;
            mbar_wait(bar_e_full(se), (uint32_t)((c / NE) & 1));
            mbar_wait(bar_a_full(s), (uint32_t)((c / NSA) & 1));
            fence_after();
            __syncwarp();
            if (elect_one()) {
                if (act) {
                    if (g == 0) tr(3, c, 1);
                    const uint64_t b_hi = umma_desc_sw128(stage_e_hi(se)), b_lo = umma_desc_sw128(stage_e_lo(se));
                    const uint32_t a_hi = tmem_base + col_stage(s, g), a_lo = a_hi + 32u;
                    const uint32_t d_tmem = tmem_base + col_p(g);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {      // 8 tf32 per step: 8 TMEM columns of A, 32 bytes inside the swizzle atom of B
                        const uint64_t adv = (uint64_t)(k * 2);
                        umma_tf32_ts(d_tmem, a_hi + 8u * k, b_hi + adv, idesc_a, (c | k) != 0);
                        umma_tf32_ts(d_tmem, a_hi + 8u * k, b_lo + adv, idesc_a, 1);
                        umma_tf32_ts(d_tmem, a_lo + 8u * k, b_hi + adv, idesc_a, 1);
                    }
                    umma_commit(bar_a_empty(s));
                    umma_commit(bar_e_empty(se));
                    if (c == n_chunks - 1) umma_commit(bar_p_full);
                    if (g == 0) tr(3, c, 2);
                } else {
                    mbar_arrive(bar_a_empty(s));
                    mbar_arrive(bar_e_empty(se));
                    if (c == n_chunks - 1) mbar_arrive(bar_p_full);
                }
            }
        }
        // ---- phase B: D_g[128 rows, 32 cols] = P_hi Qt_hi^T + P_lo Qt_hi^T + P_hi Qt_lo^T, A from tensor memory ----
        mbar_wait(bar_p_ready, 0);
        fence_after();
        const uint32_t idesc_b = idesc_tf32(128, 32);
        const uint32_t p_hi = tmem_base + col_p(g), p_lo = tmem_base + col_plo(g);
        for (int u = 0; u < n_chunks; ++u) {
            const int a = u % NACC, t = (u + 1) % NQ;
            mbar_wait(bar_acc_empty(a, g), (uint32_t)(((u / NACC) & 1) ^ 1));
            mbar_wait(bar_q_full(t), (uint32_t)((u / NQ) & 1));
            fence_after();
            __syncwarp();
            if (elect_one()) {
                if (act) {
                    if (g == 0) tr(4, u, 0);
                    const uint32_t d_tmem = tmem_base + col_acc(a, g);
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
                    umma_commit(bar_acc_full(a, g));
                    if (g == 0) tr(4, u, 1);
                } else {
                    mbar_arrive(bar_q_empty(t));
                    mbar_arrive(bar_acc_full(a, g));
                }
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

}  // namespace tc3

bool apply_tc3_available(const uce_ws* ws, int n_layers) {
    const int R = ws->rank_pad;
    return ws->K % 128 == 0 && (R == 32 || R == 64) && !ws->dense && ws->rank > 0 && n_layers <= tc3::MAX_LAYERS &&
           tensor_map_encoder() != nullptr;
}

// Tile plan: rows per block for every projection (tile_rows[l]; a CTA owns two blocks) and the total number of CTAs.
// When the whole edit fits one wave of CTAs (one per SM) the block height is the SMALLEST one >= 64 that still fits — the
// most even spread of rows, hence of TMA bytes, over the SMs — and is then evened out inside each projection; otherwise
// blocks are 128 rows.  UCE_TC3_BLOCK_ROWS (multiple of 8 in [8,128]) overrides the search.
int apply_tc3_plan(int sm_count, const int* d, int n_layers, int* tile_rows, int* tile_begin) {
    auto count = [&](int H) { long t = 0; for (int l = 0; l < n_layers; ++l) t += ceil_div(d[l], 2 * H); return t; };
    int H = 128;
    int forced = 0;
    if (const char* e = getenv("UCE_TC3_BLOCK_ROWS")) {
        const int t = atoi(e);
        if (t >= 8 && t <= 128 && t % 8 == 0) { H = t; forced = 1; }
    }
    const int sms = sm_count > 0 ? sm_count : 148;
    // never below 64 rows per block: every CTA streams all of E and Qt (16 R K bytes, as much as 2 x 43 rows of W traffic at
    // R = 64) whatever its height; short blocks multiply that stream (measured on the host path, whose launches cover 4
    // projections: 8-row blocks cost 0.6 ms of the 2.6 ms end-to-end edit through L2 contention with the PCIe copies)
    if (!forced && count(128) <= sms)
        for (int t = 64; t <= 128; t += 8)
            if (count(t) <= sms) { H = t; break; }
    int tiles = 0;
    for (int l = 0; l < n_layers; ++l) {
        const int nt = ceil_div(d[l], 2 * H);
        int hl = forced ? H : round_up(ceil_div(d[l], 2 * nt), 8);       // even split of the projection's rows over its CTAs
        if (hl > 128) hl = 128;
        tile_rows[l] = hl;
        tile_begin[l] = tiles;
        tiles += ceil_div(d[l], 2 * hl);
    }
    return tiles;
}

int apply_tc3_lowrank(uce_ws* ws, const LayerRef* layers_dev, const LayerRef* layers_host, int n_layers, int total_tiles,
                      cudaStream_t st, int* launches, cudaEvent_t ev_begin) {
    using namespace tc3;
    const int K = ws->K, R = ws->rank_pad;
    if (!apply_tc3_available(ws, n_layers)) { set_error("two-block tcgen05 apply unavailable for K=%d rank_pad=%d dense=%d layers=%d", K, R, ws->dense, n_layers); return UCE_E_STATE; }
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
        if ((rc = make_map(&wmaps.in[l], layers_host[l].w_old, layers_host[l].d, K, layers_host[l].tile_rows))) return rc;
        if ((rc = make_map(&wmaps.out[l], layers_host[l].w_new, layers_host[l].d, K, layers_host[l].tile_rows))) return rc;
    }
    const Smem L = smem_layout(R);
    const int smem = L.total;
    static thread_local int configured_dev[64] = {0};      // opt-in shared-memory size is a per-device function attribute
    int& configured = configured_dev[ws->device & 63];
    if (configured < smem) {
        UCE_CUDA(cudaFuncSetAttribute(apply_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    long long* trace = nullptr;
    const char* trace_path = getenv("UCE_TC_TRACE");
    if (trace_path) { UCE_CUDA(cudaMalloc(&trace, 7 * 64 * 4 * sizeof(long long))); UCE_CUDA(cudaMemsetAsync(trace, 0, 7 * 64 * 4 * sizeof(long long), st)); }
    // profiling: the begin event goes right before the launch — recorded earlier it would also time whatever the GPU idles while
    // the host encodes the tensor maps above (visible when the factor before it is short: cfg3 measured 65 us for a 33 us kernel)
    if (ev_begin) UCE_CUDA(cudaEventRecord(ev_begin, st));
    const char* am = getenv("UCE_TC3_ADDEND");          // "tma" (default) or "ldgsts": route of the phase-B addend (see load_box)
    const int addend_ldgsts = (am && am[0] == 'l') ? 1 : 0;
    apply_tc3_kernel<<<total_tiles, THREADS, smem, st>>>(layers_dev, n_layers, K, R, addend_ldgsts, maps, wmaps, trace);
    UCE_LAUNCH_CHECK();
    *launches += 1;
    if (trace) {   // debugging aid: dump the timeline of CTA 0 (synchronises)
        std::vector<long long> hbuf(7 * 64 * 4);
        UCE_CUDA(cudaStreamSynchronize(st));
        UCE_CUDA(cudaMemcpy(hbuf.data(), trace, hbuf.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        UCE_CUDA(cudaFree(trace));
        if (FILE* f = fopen(trace_path, "w")) {
            long long t0 = 0;
            for (long long v : hbuf) if (v && (!t0 || v < t0)) t0 = v;
            const char* roles[7] = {"w_tma", "transform", "e_tma", "mma_a", "mma_b", "epilogue", "pconv"};
            for (int r = 0; r < 7; ++r) for (int i = 0; i < 64; ++i) {
                const long long* e = &hbuf[(r * 64 + i) * 4];
                if (e[0] || e[1] || e[2]) fprintf(f, "%s %d %lld %lld %lld\n", roles[r], i, e[0] ? e[0] - t0 : -1, e[1] ? e[1] - t0 : -1, e[2] ? e[2] - t0 : -1);
            }
            fclose(f);
        }
    }
    return 0;
}

}  // namespace uce
