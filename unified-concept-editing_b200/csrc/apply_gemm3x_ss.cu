// High-rank apply on the tensor cores, SS form (rank pad > 64): the same two-launch 3xTF32 GEMM as apply_gemm3x.cu
//
//   P = W_old E^T  [d, R]   (uce_sd_erase.py:45-53)        W_new = W_old + P Q  [d, K]   (uce_sd_erase.py:61-82)
//
// but with BOTH MMA operands in shared memory and nothing staged in tensor memory:
//
//   * A_hi is the raw fp32 [128 x 32] tile where the TMA delivered it (128B swizzle = the canonical K-major UMMA layout): a
//     kind::tf32 MMA ignores the low 13 mantissa bits, i.e. it reads hi = trunc(x).  No copy (for now the truncated value is also
//     written back in place, so that correctness does not rest on that property of the hardware).
//   * A_lo = x - trunc(x) (exact in fp32) is written by 4 warps (thread = row) into a second [128 x 32] tile of the same stage,
//     published to the tensor core with fence.proxy.async + mbarrier (the pattern of apply_tc.cu's phase B).
//   * the stage (hi + lo) is released by the MMA warp's tcgen05.commit; B (E_hi|E_lo, Qt_hi|Qt_lo, pre-split with rounding by the
//     factor), the [128 x BN] accumulator in tensor memory, and the TMA-staged epilogue are those of apply_gemm3x.cu.
//
// Precision: truncation leaves |lo| <= 2^-10 |x| (rounding: 2^-11) and the tensor core keeps 11 of lo's 13 significant bits, so the
// dropped terms are 2^-21 relative per product instead of 2^-22 — 1e-6, 20x inside the 2e-5 gate of tests/test_solver_gpu.py.
//
// Why it exists: apply_gemm3x.cu (A split into TENSOR MEMORY by the transform warps) showed an intermittent error in the rows of one
// warp on hardware (profiles/r01_gemm3x_diag_*.txt) whose ring protocol is sound (scripts/protocol_sim.py).  This variant removes
// the tensor-memory staging altogether.  STATUS (round 1): written after the GPU budget was spent — opt-in only, apply impl 6
// (uce_ws_set_apply_impl / UCE_APPLY_IMPL=6), never chosen automatically; tests/test_solver_gpu.py::test_highrank_tcgen05_apply runs
// it next to impl 5 when UCE_TEST_GEMM3X=1.
#include "tc_apply_common.cuh"
#include <cstdint>
#include <cstdlib>

namespace uce {
namespace g3s {
using namespace uce::tc;
using namespace uce::tca;

constexpr int PW = 4;                                   // transform / epilogue warps
constexpr int THREADS = (PW + 3) * 32;                  // + A TMA warp + B TMA warp + MMA warp
constexpr int NST = 3, NBS = 2;                         // A stages {raw = hi 16 KB | lo 16 KB}, B stages {hi | lo}
constexpr int BN_MAX = 256;
constexpr int MAX_LAYERS = 96;
constexpr int WARP_A_TMA = PW, WARP_B_TMA = PW + 1, WARP_MMA = PW + 2;
constexpr uint32_t TMEM_COLS = 256;                     // accumulator [0, BN) only

struct Maps { CUtensorMap b_hi, b_lo, p; };
struct WMaps { CUtensorMap in[MAX_LAYERS], out[MAX_LAYERS]; };

// shared memory: [0, 96K) A ring, stage s = {raw (hi) 16 KB | lo 16 KB} | B ring NBS x {hi | lo} of BN x 128 B | barriers.  The
// epilogue's BN/32 boxes of 16 KB alias the front of this space once every MMA has completed.
struct Smem { int b_stage, b_off, bar_off, total; };
__host__ __device__ inline Smem smem_layout(int BN) {
    Smem s;
    s.b_stage = 2 * BN * 128;
    s.b_off = NST * 32768;
    const int end_main = s.b_off + NBS * s.b_stage;
    const int end_epi = (BN / 32) * 16384;
    s.bar_off = end_main > end_epi ? end_main : end_epi;
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
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(PW * 32) : "memory"); }

// mode 0:  C = P scratch rows [tile * 128, +128),  A = W_old rows of the tile            (P = W_old E^T)
// mode 1:  C = W_new rows of the tile,  A = P scratch rows,  Addend = W_old rows         (W_new = W_old + P Q)
__global__ void __launch_bounds__(THREADS, 1)
gemm3x_ss_kernel(const LayerRef* __restrict__ layers, int n_layers, int mode, int Kd, int N, int BN,
              const __grid_constant__ Maps maps, const __grid_constant__ WMaps wmaps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    if (base & 1023u) {
        if (threadIdx.x == 0) printf("uce gemm3x_ss: dynamic shared memory base %u is not 1024-byte aligned\n", base);
        __trap();
    }
    const Smem L = smem_layout(BN);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bars = base + L.bar_off;
    auto bar_raw_full  = [&](int r) { return bars + 8u * r; };                 // [0,3)  TMA -> lo warps
    auto bar_lo_full   = [&](int r) { return bars + 8u * (3 + r); };           // [3,6)  lo warps -> MMA (implies raw_full)
    auto bar_st_empty  = [&](int r) { return bars + 8u * (6 + r); };           // [6,9)  MMA commit -> TMA (hi and lo tiles free)
    auto bar_b_full    = [&](int s) { return bars + 8u * (9 + s); };           // [9,11)
    auto bar_b_empty   = [&](int s) { return bars + 8u * (11 + s); };          // [11,13)
    const uint32_t bar_acc_full = bars + 8u * 13, bar_add_full = bars + 8u * 14, tmem_slot = bars + 8u * 15;

    const int tile = blockIdx.y;
    const int layer = find_layer(layers, n_layers, tile);
    const LayerRef Lr = layers[layer];
    const int row0 = (tile - Lr.tile_begin) * 128;                // first row of the tile inside its projection
    const int n0 = blockIdx.x * BN;                               // first output column
    const int n_cols = min(BN, N - n0);
    const int n_boxes = (n_cols + 31) / 32;
    const int n_chunks = Kd / 32;

    if (threadIdx.x == 0) {
        for (int r = 0; r < NST; ++r) { mbar_init(bar_raw_full(r), 1); mbar_init(bar_lo_full(r), PW); mbar_init(bar_st_empty(r), 1); }
        for (int s = 0; s < NBS; ++s) { mbar_init(bar_b_full(s), 1); mbar_init(bar_b_empty(s), 1); }
        mbar_init(bar_acc_full, 1); mbar_init(bar_add_full, 1);
        mbar_fence_init();
    }
    if (warp == WARP_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
    if (warp == WARP_B_TMA && lane == 0) {
        tma_prefetch_desc(&maps.b_hi); tma_prefetch_desc(&maps.b_lo); tma_prefetch_desc(&maps.p);
        tma_prefetch_desc(&wmaps.in[layer]); tma_prefetch_desc(&wmaps.out[layer]);
    }
    fence_before();
    __syncthreads();
    fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    auto raw_st = [&](int r) { return base + (uint32_t)(r * 32768); };
    auto lo_st = [&](int r) { return base + (uint32_t)(r * 32768 + 16384); };
    auto b_hi_st = [&](int s) { return base + (uint32_t)(L.b_off + s * L.b_stage); };
    auto b_lo_st = [&](int s) { return base + (uint32_t)(L.b_off + s * L.b_stage + BN * 128); };
    auto box_st = [&](int b) { return base + (uint32_t)(b * 16384); };
    const CUtensorMap* a_map = mode ? &maps.p : &wmaps.in[layer];
    const int a_row = mode ? tile * 128 : row0;

    if (warp < PW) {
        // =============================== A_lo tiles, then epilogue ===============================
        const int trow = 32 * warp + lane;               // tile row == TMEM lane of the accumulator
        const uint32_t lane_base = tmem_base + ((uint32_t)(32 * warp) << 16);
        const uint32_t row_off = (uint32_t)(trow * 128);
        const uint32_t sw = (uint32_t)(trow & 7);
        for (int c = 0; c < n_chunks; ++c) {
            const int r = c % NST;
            mbar_wait(bar_raw_full(r), (uint32_t)((c / NST) & 1));          // the stage is free (the TMA warp waited for it) and loaded
            const uint32_t raw = raw_st(r) + row_off, lo = lo_st(r) + row_off;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t off = ((uint32_t)j ^ sw) << 4;
                const float4 v = lds_v4(raw + off);                            // rows beyond the tensor were zero-filled by TMA
                const float hx = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u), hy = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                const float hz = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u), hw = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                sts_v4(lo + off, v.x - hx, v.y - hy, v.z - hz, v.w - hw);
                // the truncated hi is written back in place, so the result does not depend on what the tensor core does with the low
                // 13 bits of a tf32 operand (documented as ignored); drop this store once that has been confirmed on hardware
                sts_v4(raw + off, hx, hy, hz, hw);
            }
            fence_proxy_async();                                               // generic writes -> visible to the tensor core's reads
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_lo_full(r));
        }
        // ---- epilogue: accumulator (+ addend box) -> swizzled boxes -> TMA stores ----
        mbar_wait(bar_acc_full, 0);
        fence_after();
        if (mode) mbar_wait(bar_add_full, 0);
        for (int b = 0; b < n_boxes; ++b) {
            uint32_t v[32];
            tmem_ld32(lane_base + (uint32_t)(32 * b), v);
            const uint32_t row = box_st(b) + row_off;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t addr = row + (((uint32_t)j ^ sw) << 4);
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                if (mode) w = lds_v4(addr);
                sts_v4(addr, w.x + __uint_as_float(v[4 * j]), w.y + __uint_as_float(v[4 * j + 1]),
                       w.z + __uint_as_float(v[4 * j + 2]), w.w + __uint_as_float(v[4 * j + 3]));
            }
        }
        fence_proxy_async();
        epi_bar_sync();
        if (warp == 0) {
            if (elect_one()) {
                const uint64_t pol = l2_evict_first();
                const CUtensorMap* om = mode ? &wmaps.out[layer] : &maps.p;
                const int o_row = mode ? row0 : tile * 128;
                for (int b = 0; b < n_boxes; ++b) tma_store_2d(om, box_st(b), n0 + 32 * b, o_row, pol);
                tma_store_commit();
                tma_store_wait_read<0>();              // shared memory must outlive the stores' reads
            }
            __syncwarp();
        }
    } else if (warp == WARP_A_TMA) {
        // =============================== TMA warp 1: raw A chunks, then (mode 1) the addend boxes ===============================
        const uint64_t pol = mode ? l2_evict_first() : l2_evict_last();       // W_old is read again by the second launch
        for (int c = 0; c < n_chunks; ++c) {
            const int r = c % NST;
            mbar_wait(bar_st_empty(r), (uint32_t)(((c / NST) & 1) ^ 1));     // the MMAs that read this stage (hi and lo) have completed
            __syncwarp();
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_raw_full(r), 16384u);
                tma_load_2d_hint(raw_st(r), a_map, bar_raw_full(r), c * 32, a_row, pol);
            }
        }
        if (mode) {                                    // the boxes alias the rings: every MMA must have completed
            mbar_wait(bar_acc_full, 0);
            __syncwarp();
            if (elect_one()) {
                const uint64_t pol_last = l2_evict_first();
                mbar_arrive_expect_tx(bar_add_full, (uint32_t)n_boxes * 16384u);
                for (int b = 0; b < n_boxes; ++b) tma_load_2d_hint(box_st(b), &wmaps.in[layer], bar_add_full, n0 + 32 * b, row0, pol_last);
            }
        }
    } else if (warp == WARP_B_TMA) {
        // =============================== TMA warp 2: B_hi | B_lo tiles ===============================
        const uint32_t b_bytes = 2u * (uint32_t)BN * 128u;
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % NBS;
            mbar_wait(bar_b_empty(s), (uint32_t)(((c / NBS) & 1) ^ 1));
            __syncwarp();
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_b_full(s), b_bytes);
                tma_load_2d(b_hi_st(s), &maps.b_hi, bar_b_full(s), c * 32, n0);
                tma_load_2d(b_lo_st(s), &maps.b_lo, bar_b_full(s), c * 32, n0);
            }
        }
    } else {
        // =============================== MMA issuer ===============================
        const uint32_t idesc = idesc_tf32(128, BN);
        for (int c = 0; c < n_chunks; ++c) {
            const int r = c % NST, sb = c % NBS;
            mbar_wait(bar_b_full(sb), (uint32_t)((c / NBS) & 1));
            mbar_wait(bar_lo_full(r), (uint32_t)((c / NST) & 1));
            fence_after();
            __syncwarp();
            if (elect_one()) {
                const uint64_t a_hi = umma_desc_sw128(raw_st(r)), a_lo = umma_desc_sw128(lo_st(r));
                const uint64_t b_hi = umma_desc_sw128(b_hi_st(sb)), b_lo = umma_desc_sw128(b_lo_st(sb));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t adv = (uint64_t)(k * 2);                   // 8 tf32 = 32 bytes along K, >> 4
                    umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, (c | k) != 0);
                    umma_tf32(tmem_base, a_hi + adv, b_lo + adv, idesc, 1);
                    umma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc, 1);
                }
                umma_commit(bar_st_empty(r));
                umma_commit(bar_b_empty(sb));
                if (c == n_chunks - 1) umma_commit(bar_acc_full);
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

static int make_map(CUtensorMap* m, const float* ptr, long rows, int cols, int box_rows) {
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

}  // namespace g3s

bool apply_gemm3x_ss_available(const uce_ws* ws, int n_layers) {
    return ws->K % 32 == 0 && ws->rank_pad % 32 == 0 && ws->rank > 0 && !ws->dense && n_layers <= g3s::MAX_LAYERS && tensor_map_encoder() != nullptr;
}

// P scratch: total_tiles * 128 rows of rank_pad floats (ws->P, sized by the caller); tiles of 128 rows (LayerRef.tile_begin).
int apply_gemm3x_ss_highrank(uce_ws* ws, const LayerRef* layers_dev, const LayerRef* layers_host, int n_layers, int total_tiles,
                          cudaStream_t st, int* launches) {
    using namespace g3s;
    const int K = ws->K, R = ws->rank_pad;
    if (!apply_gemm3x_ss_available(ws, n_layers)) { set_error("tcgen05 high-rank apply (SS form) unavailable for K=%d rank_pad=%d dense=%d layers=%d", K, R, ws->dense, n_layers); return UCE_E_STATE; }
    for (int l = 0; l < n_layers; ++l)
        if (((uintptr_t)layers_host[l].w_old & 15) || ((uintptr_t)layers_host[l].w_new & 15)) { set_error("tcgen05 apply needs 16-byte aligned weights"); return UCE_E_ARG; }
    static thread_local WMaps wmaps;      // kept off the stack, one per host thread (handles are independent); copied into the launch by value
    int rc;
    for (int l = 0; l < n_layers; ++l) {
        if ((rc = make_map(&wmaps.in[l], layers_host[l].w_old, layers_host[l].d, K, 128))) return rc;
        if ((rc = make_map(&wmaps.out[l], layers_host[l].w_new, layers_host[l].d, K, 128))) return rc;
    }
    static thread_local int configured_dev[64] = {0};      // opt-in shared-memory size is a per-device function attribute
    int& configured = configured_dev[ws->device & 63];
    for (int pass = 0; pass < 2; ++pass) {
        // pass 0: P[M, R] = W_old[M, K] . E[R, K]^T        pass 1: W_new[M, K] = W_old + P[M, R] . Qt[K, R]^T
        const int N = pass ? K : R, Kd = pass ? R : K;
        const int BN = N < BN_MAX ? N : BN_MAX;
        Maps maps;
        if ((rc = make_map(&maps.b_hi, pass ? ws->Qt_hi : ws->E_hi, N, Kd, BN))) return rc;
        if ((rc = make_map(&maps.b_lo, pass ? ws->Qt_lo : ws->E_lo, N, Kd, BN))) return rc;
        if ((rc = make_map(&maps.p, ws->P, (long)total_tiles * 128, R, 128))) return rc;
        const int smem = smem_layout(BN).total;
        if (configured < smem) {
            UCE_CUDA(cudaFuncSetAttribute(gemm3x_ss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured = smem;
        }
        gemm3x_ss_kernel<<<dim3((unsigned)ceil_div(N, BN), (unsigned)total_tiles), THREADS, smem, st>>>(layers_dev, n_layers, pass, Kd, N, BN, maps, wmaps);
        UCE_LAUNCH_CHECK();
        *launches += 1;
    }
    return 0;
}

}  // namespace uce
