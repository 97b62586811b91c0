// High-rank apply on the tensor cores (rank pad > 64, e.g. BASELINE config 4: SDXL, 1 000 erase concepts, K = 2 048):
//
//   P     = W_old E^T                 [d, R]   (guide outputs folded with the edit rows, uce_sd_erase.py:45-53)
//   W_new = W_old + P Q               [d, K]   (mat1 @ inverse(mat2), uce_sd_erase.py:61-82)
//
// as TWO launches of one fp32-fidelity GEMM kernel  C[M,N] = A[M,Kd] . B[N,Kd]^T (+ Addend):  3xTF32 (hi.hi + hi.lo + lo.hi,
// lo.lo dropped) on tcgen05 with the accumulator in tensor memory.  P does not fit tensor memory here (R columns > 512 - stages),
// so unlike apply_tc3.cu it travels through an HBM scratch between the two launches.
//
//   * A (W_old, then P) arrives as raw fp32 [128 x 32] chunks by TMA; 4 transform warps (thread = row = TMEM lane) split every value
//     into hi / lo and store both into one of four 64-column A stages in TENSOR MEMORY (same transform as apply_tc3.cu)
//   * B (E_hi|E_lo [R,K], then Qt_hi|Qt_lo [K,R]) is pre-split by the factor; [BN x 32] tiles of both by TMA, 2 stages
//   * MMA warp (converged, one elected lane issues): per 8-wide k-step three N = BN MMAs into the [128 x BN] accumulator
//   * epilogue: accumulator -> registers -> 128B-swizzled [128 x 32] boxes in shared memory (+ the W_old addend, which the TMA warp
//     fetched into the same boxes) -> TMA stores
//
// STATUS (round 1): opt-in only — apply impl 5 (uce_ws_set_apply_impl / UCE_APPLY_IMPL=5), never chosen automatically; the SIMT
// kernels remain the high-rank path.  Three short hardware runs at the very end of the round (scripts/gemm3x_diag.py,
// profiles/r01_gemm3x_diag_*.txt): no hang, no trap, and the update agrees with the SIMT apply to 5e-6 for rank pads 64 / 96 / 224 /
// 256 in most launches — but NOT in all of them: a 200-row projection came out 4e-2..1e-1 off in the first 32 rows of both of its
// row tiles in one process and exact in the next one; a 264-row projection was off in its last 8 rows (the first rows of its third
// tile).  Always the rows of warp 0 (TMEM lanes 0-31) of a tile, and not reproducible per shape: a race, not an indexing error.
// Warp 0 differs from warps 1-3 only in the epilogue (it issues the TMA stores).  tests/test_solver_gpu.py::
// test_highrank_tcgen05_apply stays skipped (UCE_TEST_GEMM3X=1 runs it).  Round 2 starts here.
#include "tc_apply_common.cuh"
#include <cstdint>
#include <cstdlib>

namespace uce {
namespace g3 {
using namespace uce::tc;
using namespace uce::tca;

constexpr int PW = 4;                                   // transform / epilogue warps
constexpr int THREADS = (PW + 3) * 32;                  // + A TMA warp + B TMA warp + MMA warp
constexpr int NRAW = 4, NSA = 4, NBS = 2;
constexpr int BN_MAX = 256;
constexpr int MAX_LAYERS = 96;
constexpr int WARP_A_TMA = PW, WARP_B_TMA = PW + 1, WARP_MMA = PW + 2;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t A_COL0 = 256;                        // accumulator [0, BN), A stages {hi 32, lo 32} at [256 + 64 s, +64)

struct Maps { CUtensorMap b_hi, b_lo, p; };
struct WMaps { CUtensorMap in[MAX_LAYERS], out[MAX_LAYERS]; };

// shared memory: [0, 64K) raw A ring | [64K, 64K + NBS * 2 * BN * 128) B ring {hi | lo} | barriers.  The epilogue's BN/32 boxes of
// 16 KB alias the front of this space once every MMA has completed.
struct Smem { int b_stage, b_off, bar_off, total; };
__host__ __device__ inline Smem smem_layout(int BN) {
    Smem s;
    s.b_stage = 2 * BN * 128;
    s.b_off = NRAW * 16384;
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
gemm3x_kernel(const LayerRef* __restrict__ layers, int n_layers, int mode, int Kd, int N, int BN,
              const __grid_constant__ Maps maps, const __grid_constant__ WMaps wmaps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    if (base & 1023u) {
        if (threadIdx.x == 0) printf("uce gemm3x: dynamic shared memory base %u is not 1024-byte aligned\n", base);
        __trap();
    }
    const Smem L = smem_layout(BN);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bars = base + L.bar_off;
    auto bar_raw_full  = [&](int r) { return bars + 8u * r; };                 // [0,4)
    auto bar_raw_empty = [&](int r) { return bars + 8u * (4 + r); };           // [4,8)
    auto bar_a_full    = [&](int s) { return bars + 8u * (8 + s); };           // [8,12)
    auto bar_a_empty   = [&](int s) { return bars + 8u * (12 + s); };          // [12,16)
    auto bar_b_full    = [&](int s) { return bars + 8u * (16 + s); };          // [16,18)
    auto bar_b_empty   = [&](int s) { return bars + 8u * (18 + s); };          // [18,20)
    const uint32_t bar_acc_full = bars + 8u * 20, bar_add_full = bars + 8u * 21, tmem_slot = bars + 8u * 22;

    const int tile = blockIdx.y;
    const int layer = find_layer(layers, n_layers, tile);
    const LayerRef Lr = layers[layer];
    const int row0 = (tile - Lr.tile_begin) * 128;                // first row of the tile inside its projection
    const int n0 = blockIdx.x * BN;                               // first output column
    const int n_cols = min(BN, N - n0);
    const int n_boxes = (n_cols + 31) / 32;
    const int n_chunks = Kd / 32;

    if (threadIdx.x == 0) {
        for (int r = 0; r < NRAW; ++r) { mbar_init(bar_raw_full(r), 1); mbar_init(bar_raw_empty(r), PW); }
        for (int s = 0; s < NSA; ++s) { mbar_init(bar_a_full(s), PW); mbar_init(bar_a_empty(s), 1); }
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

    auto raw_st = [&](int r) { return base + (uint32_t)(r * 16384); };
    auto b_hi_st = [&](int s) { return base + (uint32_t)(L.b_off + s * L.b_stage); };
    auto b_lo_st = [&](int s) { return base + (uint32_t)(L.b_off + s * L.b_stage + BN * 128); };
    auto box_st = [&](int b) { return base + (uint32_t)(b * 16384); };
    const CUtensorMap* a_map = mode ? &maps.p : &wmaps.in[layer];
    const int a_row = mode ? tile * 128 : row0;

    if (warp < PW) {
        // =============================== A transform, then epilogue ===============================
        const int trow = 32 * warp + lane;               // tile row == TMEM lane
        const uint32_t lane_base = tmem_base + ((uint32_t)(32 * warp) << 16);
        const uint32_t row_off = (uint32_t)(trow * 128);
        const uint32_t sw = (uint32_t)(trow & 7);
        // The stores into tensor memory are waited for (tcgen05.wait::st) BEFORE the registers they read are written again.  A
        // software-pipelined variant that left them in flight across the next chunk's split was measured neutral in time, and the
        // same pattern in apply_gemm3x.cu intermittently garbled one chunk of one warp's rows: the source registers of an
        // asynchronous tcgen05.st are not safe to overwrite before the wait, just as tcgen05.ld results are not safe to read.
        for (int c = 0; c < n_chunks; ++c) {
            const int r = c % NRAW, s = c % NSA;
            mbar_wait(bar_raw_full(r), (uint32_t)((c / NRAW) & 1));
            const uint32_t raw = raw_st(r) + row_off;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 v = lds_v4(raw + (((uint32_t)j ^ sw) << 4));       // rows beyond the tensor were zero-filled by TMA
                tf32_split(v.x, hi[4 * j], lo[4 * j]);         tf32_split(v.y, hi[4 * j + 1], lo[4 * j + 1]);
                tf32_split(v.z, hi[4 * j + 2], lo[4 * j + 2]); tf32_split(v.w, hi[4 * j + 3], lo[4 * j + 3]);
            }
            mbar_wait(bar_a_empty(s), (uint32_t)(((c / NSA) & 1) ^ 1));      // the MMAs that read this A stage have completed
            fence_after();
            __syncwarp();                                                    // lane 0 took the `if (lane == 0)` arrive path last iteration: .sync.aligned needs the warp converged
            const uint32_t ta = lane_base + A_COL0 + (uint32_t)(64 * s);
            tmem_st32(ta, hi);
            tmem_st32(ta + 32u, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_a_full(s)); mbar_arrive(bar_raw_empty(r)); }
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
            const int r = c % NRAW;
            mbar_wait(bar_raw_empty(r), (uint32_t)(((c / NRAW) & 1) ^ 1));
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
            const int s = c % NSA, sb = c % NBS;
            mbar_wait(bar_b_full(sb), (uint32_t)((c / NBS) & 1));
            mbar_wait(bar_a_full(s), (uint32_t)((c / NSA) & 1));
            fence_after();
            __syncwarp();
            if (elect_one()) {
                const uint32_t a_hi = tmem_base + A_COL0 + (uint32_t)(64 * s), a_lo = a_hi + 32u;
                const uint64_t b_hi = umma_desc_sw128(b_hi_st(sb)), b_lo = umma_desc_sw128(b_lo_st(sb));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t adv = (uint64_t)(k * 2);
                    umma_tf32_ts(tmem_base, a_hi + 8u * k, b_hi + adv, idesc, (c | k) != 0);
                    umma_tf32_ts(tmem_base, a_hi + 8u * k, b_lo + adv, idesc, 1);
                    umma_tf32_ts(tmem_base, a_lo + 8u * k, b_hi + adv, idesc, 1);
                }
                umma_commit(bar_a_empty(s));
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

}  // namespace g3

bool apply_gemm3x_available(const uce_ws* ws, int n_layers) {
    return ws->K % 32 == 0 && ws->rank_pad % 32 == 0 && ws->rank > 0 && !ws->dense && n_layers <= g3::MAX_LAYERS && tensor_map_encoder() != nullptr;
}

// P scratch: total_tiles * 128 rows of rank_pad floats (ws->P, sized by the caller); tiles of 128 rows (LayerRef.tile_begin).
// stage: 0 both passes on `st`; 1 only pass 1 (P = W_old E^T: needs E only — uce_edit_dev_f32 runs it beside the factor); 2 only pass 2
int apply_gemm3x_highrank(uce_ws* ws, const LayerRef* layers_dev, const LayerRef* layers_host, int n_layers, int total_tiles,
                          cudaStream_t st, int* launches, int stage) {
    using namespace g3;
    const int K = ws->K, R = ws->rank_pad;
    if (!apply_gemm3x_available(ws, n_layers)) { set_error("tcgen05 high-rank apply unavailable for K=%d rank_pad=%d dense=%d layers=%d", K, R, ws->dense, n_layers); return UCE_E_STATE; }
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
    float* Pscr = ws->P + ws->P_off;
    for (int pass = 0; pass < 2; ++pass) {
        if ((stage == 1 && pass == 1) || (stage == 2 && pass == 0)) continue;
        // pass 0: P[M, R] = W_old[M, K] . E[R, K]^T        pass 1: W_new[M, K] = W_old + P[M, R] . Qt[K, R]^T
        const int N = pass ? K : R, Kd = pass ? R : K;
        const int BN = N < BN_MAX ? N : BN_MAX;
        Maps maps;
        if ((rc = make_map(&maps.b_hi, pass ? ws->Qt_hi : ws->E_hi, N, Kd, BN))) return rc;
        if ((rc = make_map(&maps.b_lo, pass ? ws->Qt_lo : ws->E_lo, N, Kd, BN))) return rc;
        if ((rc = make_map(&maps.p, Pscr, (long)total_tiles * 128, R, 128))) return rc;
        const int smem = smem_layout(BN).total;
        if (configured < smem) {
            UCE_CUDA(cudaFuncSetAttribute(gemm3x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured = smem;
        }
        gemm3x_kernel<<<dim3((unsigned)ceil_div(N, BN), (unsigned)total_tiles), THREADS, smem, st>>>(layers_dev, n_layers, pass, Kd, N, BN, maps, wmaps);
        UCE_LAUNCH_CHECK();
        *launches += 1;
    }
    return 0;
}

}  // namespace uce
