// Phase 2 of the UCE edit: per-projection apply, batched over all projections.
//
//   low rank :  P = W_old E^T  [d, r]      (the reference's guide outputs v* = W_old c, uce_sd_erase.py:45-53,
//                                            folded with the edit rows into E = G_e - C_e)
//               W_new = W_old + P Q         (mat1 @ inverse(mat2), uce_sd_erase.py:61-82)
//   dense    :  W_new = W_old + W_old D,  D = E^T Q  [K, K]   (when r > K/2)
//
// This file holds the fp32 SIMT implementation (grouped over projections through a flattened
// row-tile list) and the dispatch; apply_tc3.cu / apply_gemm3x.cu hold the tcgen05 3xTF32 implementations validated against it.
#include "uce_ws.h"
#include "gemm_simt.cuh"
#include <algorithm>
#include <cstring>
#include <vector>

namespace uce {

bool apply_gemm3x_available(const uce_ws* ws, int n_layers);   // apply_gemm3x.cu
int apply_gemm3x_highrank(uce_ws* ws, const LayerRef* layers_dev, const LayerRef* layers_host, int n_layers, int total_tiles,
                          cudaStream_t st, int* launches, int stage);
bool apply_ab_available(const uce_ws* ws, int n_layers);    // apply_ab.cu
int apply_ab_ksplit(int K);
int apply_ab_plan(int sm_count, int ks, const int* d, int n_layers, int* block_rows, int* first_block);
int apply_ab_lowrank(uce_ws* ws, const void* slots_dev, const void* slots_host, int n_slots, const LayerRef* layers_host, int n_layers,
                     cudaStream_t st, int stage, int* launches, cudaEvent_t ev_mid);
size_t apply_ab_slot_bytes();
void apply_ab_fill_slots(void* slots_host, const LayerRef* layers_host, int n_layers);
bool apply_tc3_available(const uce_ws* ws, int n_layers);   // apply_tc3.cu
int apply_tc3_plan(int sm_count, const int* d, int n_layers, int* tile_rows, int* tile_begin);
int apply_tc3_lowrank(uce_ws* ws, const LayerRef* layers_dev, const LayerRef* layers_host, int n_layers, int total_tiles,
                      cudaStream_t st, int* launches, cudaEvent_t ev_begin);

// Small host tables (row-block slots, layer references) reach the device as KERNEL PARAMETERS, not through a copy engine: inside the
// host-buffer call the host-to-device engine is busy with multi-megabyte weight uploads and serves copies one at a time, so a 1 KB table
// copy waited for a whole upload group and delayed every apply — and with it every download — by one group (measured: 2.9 ms per call at
// 4 groups where the copies alone overlap in 1.9).  Parameters are baked into a captured graph, so a replay rewrites the same table.
struct TableChunk { unsigned int w[4096]; };                  // 16 KB per launch (kernel parameters may hold 32 KB; the K-split kernels pass 24 KB of tensor maps)
__global__ void __launch_bounds__(256) table_write_kernel(unsigned int* dst, const __grid_constant__ TableChunk c, int n_words) {
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = c.w[i];
}
int table_upload(void* dst, const void* src_host, size_t bytes, cudaStream_t st, int* launches) {
    const size_t words = (bytes + 3) / 4;
    for (size_t w0 = 0; w0 < words; w0 += 4096) {
        TableChunk c;
        const size_t n = std::min<size_t>(4096, words - w0);
        memcpy(c.w, (const unsigned int*)src_host + w0, std::min(n * 4, bytes - w0 * 4));
        table_write_kernel<<<1, 256, 0, st>>>((unsigned int*)dst + w0, c, (int)n);
        UCE_LAUNCH_CHECK();
        if (launches) ++*launches;
    }
    return 0;
}

__device__ __forceinline__ int find_layer(const LayerRef* layers, int n_layers, int tile) {
    int lo = 0, hi = n_layers - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (layers[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// stage 1: P[tile rows, :r_pad] = W_old[rows, :] . E^T
__global__ void __launch_bounds__(SG_THREADS) apply_simt_p_kernel(const LayerRef* layers, int n_layers, int K, int r_pad,
                                                                  const float* E, float* P) {
    __shared__ float As[SG_BK][SG_BM + 4];
    __shared__ float Bs[SG_BK][SG_BN + 4];
    const int tile = blockIdx.y;
    const LayerRef L = layers[find_layer(layers, n_layers, tile)];
    const int lt = tile - L.tile_begin;
    SimtGemmArgs<float, float, float, float> g;
    g.M = min(SG_BM, L.d - lt * SG_BM); g.N = r_pad; g.Kd = K;
    g.A = L.w_old + (long)lt * SG_BM * K; g.sa_m = K; g.sa_k = 1;
    g.B = E; g.sb_n = K; g.sb_k = 1;
    g.C = P + (long)tile * SG_BM * r_pad; g.ldc = r_pad;
    g.Add = nullptr; g.ldadd = 0; g.alpha = 1.f; g.beta = 0.f; g.lower_only = 0;
    simt_gemm_tile<float, float, float, float>(g, 0, blockIdx.x, As, Bs);
}

// stage 2: W_new[rows, :] = W_old[rows, :] + P[rows, :r_pad] . Qt^T   (Qt [K, r_pad])
__global__ void __launch_bounds__(SG_THREADS) apply_simt_w_kernel(const LayerRef* layers, int n_layers, int K, int r_pad,
                                                                  const float* Qt, const float* P) {
    __shared__ float As[SG_BK][SG_BM + 4];
    __shared__ float Bs[SG_BK][SG_BN + 4];
    const int tile = blockIdx.y;
    const LayerRef L = layers[find_layer(layers, n_layers, tile)];
    const int lt = tile - L.tile_begin;
    SimtGemmArgs<float, float, float, float> g;
    g.M = min(SG_BM, L.d - lt * SG_BM); g.N = K; g.Kd = r_pad;
    g.A = P + (long)tile * SG_BM * r_pad; g.sa_m = r_pad; g.sa_k = 1;
    g.B = Qt; g.sb_n = r_pad; g.sb_k = 1;
    g.C = L.w_new + (long)lt * SG_BM * K; g.ldc = K;
    g.Add = L.w_old + (long)lt * SG_BM * K; g.ldadd = K; g.alpha = 1.f; g.beta = 1.f; g.lower_only = 0;
    simt_gemm_tile<float, float, float, float>(g, 0, blockIdx.x, As, Bs);
}

// dense: out[rows, :] = (add ? W_old : 0) + W_old[rows, :] . Dt^T ; out = W_new or scratch
__global__ void __launch_bounds__(SG_THREADS) apply_simt_dense_kernel(const LayerRef* layers, int n_layers, int K, const float* Dt,
                                                                      float* scratch) {
    __shared__ float As[SG_BK][SG_BM + 4];
    __shared__ float Bs[SG_BK][SG_BN + 4];
    const int tile = blockIdx.y;
    const LayerRef L = layers[find_layer(layers, n_layers, tile)];
    const int lt = tile - L.tile_begin;
    SimtGemmArgs<float, float, float, float> g;
    g.M = min(SG_BM, L.d - lt * SG_BM); g.N = K; g.Kd = K;
    g.A = L.w_old + (long)lt * SG_BM * K; g.sa_m = K; g.sa_k = 1;
    g.B = Dt; g.sb_n = K; g.sb_k = 1;
    if (scratch) {   // in-place edit: delta goes to scratch first (other CTAs still read these rows of W_old)
        g.C = scratch + (long)tile * SG_BM * K; g.ldc = K; g.Add = nullptr; g.ldadd = 0; g.beta = 0.f;
    } else {
        g.C = L.w_new + (long)lt * SG_BM * K; g.ldc = K; g.Add = g.A; g.ldadd = K; g.beta = 1.f;
    }
    g.alpha = 1.f; g.lower_only = 0;
    simt_gemm_tile<float, float, float, float>(g, 0, blockIdx.x, As, Bs);
}

__global__ void add_scratch_kernel(const LayerRef* layers, int n_layers, int K, const float* scratch) {
    const int tile = blockIdx.x;
    const LayerRef L = layers[find_layer(layers, n_layers, tile)];
    const int lt = tile - L.tile_begin;
    const int rows = min(SG_BM, L.d - lt * SG_BM);
    const float* s = scratch + (long)tile * SG_BM * K;
    const float* wo = L.w_old + (long)lt * SG_BM * K;
    float* wn = L.w_new + (long)lt * SG_BM * K;
    for (long i = threadIdx.x; i < (long)rows * K; i += blockDim.x) wn[i] = wo[i] + s[i];
}

__global__ void copy_layers_kernel(const LayerRef* layers, int n_layers, int K) {
    const int tile = blockIdx.x;
    const LayerRef L = layers[find_layer(layers, n_layers, tile)];
    if (L.w_new == L.w_old) return;
    const int lt = tile - L.tile_begin;
    const int rows = min(SG_BM, L.d - lt * SG_BM);
    const float* wo = L.w_old + (long)lt * SG_BM * K;
    float* wn = L.w_new + (long)lt * SG_BM * K;
    for (long i = threadIdx.x; i < (long)rows * K; i += blockDim.x) wn[i] = wo[i];
}

static bool choose_ab(const uce_ws* ws, int n_layers) {
    const bool lowrank = !ws->dense && ws->rank > 0;
    return lowrank && (ws->apply_impl == 7 || ws->apply_impl == 0) && apply_ab_available(ws, n_layers);
}
static bool choose_g3(const uce_ws* ws, int n_layers) {
    const bool lowrank = !ws->dense && ws->rank > 0;
    if (!lowrank || choose_ab(ws, n_layers)) return false;
    if (ws->apply_impl == 4 || (ws->apply_impl == 0 && apply_tc3_available(ws, n_layers))) return false;
    return (ws->apply_impl == 5 || ws->apply_impl == 0) && apply_gemm3x_available(ws, n_layers);
}
// both two-kernel tcgen05 forms have a first kernel (W_old E^T) that needs only E: it can run beside the factor
bool apply_stage_split(const uce_ws* ws, int n_layers) {
    const int nl = n_layers > 96 ? 96 : n_layers;
    return ws->mode != 0 && (choose_ab(ws, nl) || choose_g3(ws, nl));
}
// scratch floats one call of `nl` projections needs (two-kernel forms; 0 otherwise)
static size_t staged_scratch(const uce_ws* ws, const int* d, int nl) {
    if (choose_ab(ws, nl)) {
        std::vector<int> a(nl), b(nl);
        const int ks = apply_ab_ksplit(ws->K);
        return (size_t)ks * apply_ab_plan(ws->sm_count, ks, d, nl, a.data(), b.data()) * 128 * ws->rank_pad;
    }
    if (choose_g3(ws, nl)) { size_t t = 0; for (int l = 0; l < nl; ++l) t += ceil_div(d[l], 128); return t * 128 * ws->rank_pad; }
    return 0;
}

int apply_dev(uce_ws* ws, const float* const* W_old, float* const* W_new, const int* d, int n_layers, cudaStream_t st,
              bool no_profile, int stage) {
    int slot_begin = 0;
    const int K = ws->K;
    if (ws->mode == 0) { set_error("uce_apply before uce_factor"); return UCE_E_STATE; }
    // the two-block tcgen05 apply carries two tensor maps per projection as kernel parameters (96 projections per launch):
    // longer lists (SDXL: 140 projections) go through it in slices
    constexpr int TC3_MAX = 96;
    if (stage != 0 && !apply_stage_split(ws, n_layers)) { set_error("staged apply needs one of the two-kernel tcgen05 paths"); return UCE_E_STATE; }
    if (n_layers > TC3_MAX && !ws->dense && ws->rank > 0 && ws->apply_impl != 1 &&
        (apply_ab_available(ws, TC3_MAX) || apply_tc3_available(ws, TC3_MAX) || apply_gemm3x_available(ws, TC3_MAX))) {
        // slices of 96 projections.  A staged call (stage 1 for every slice, later stage 2 for every slice) needs every slice's partial
        // products at the same time: each slice gets its own part of the scratch (P_off); unstaged calls reuse one part.
        size_t total_need = 0, max_need = 0;
        for (int l0 = 0; l0 < n_layers; l0 += TC3_MAX) {
            const size_t nd = staged_scratch(ws, d + l0, std::min(TC3_MAX, n_layers - l0));
            total_need += nd; max_need = std::max(max_need, nd);
        }
        const size_t want = stage != 0 ? total_need : max_need;
        if (want > ws->P_cap) {
            if (stage == 2) { set_error("apply scratch changed between the stages"); return UCE_E_STATE; }
            if (ws->P) { UCE_CUDA(cudaStreamSynchronize(st)); UCE_CUDA(cudaFree(ws->P)); ws->P = nullptr; }
            ws->P_cap = want + want / 4;
            UCE_CUDA(cudaMalloc(&ws->P, ws->P_cap * sizeof(float)));
        }
        int total_launches = 0;
        const bool prof = ws->profile && !no_profile;
        if (prof) UCE_CUDA(cudaEventRecord(ws->pev[2], st));
        size_t off = 0;
        for (int l0 = 0; l0 < n_layers; l0 += TC3_MAX) {
            const int nl = std::min(TC3_MAX, n_layers - l0);
            ws->P_off = stage != 0 ? off : 0;
            const int rc = apply_dev(ws, W_old + l0, W_new + l0, d + l0, nl, st, true, stage);
            ws->P_off = 0;
            if (rc) return rc;
            total_launches += ws->launches_apply;
            off += staged_scratch(ws, d + l0, nl);
        }
        ws->launches_apply = total_launches;
        ws->pev_mid = 0;
        if (prof) UCE_CUDA(cudaEventRecord(ws->pev[4], st));
        return 0;
    }
    if (n_layers > ws->layers_cap / 4) { set_error("too many layers per call (%d > %d)", n_layers, ws->layers_cap / 4); return UCE_E_STATE; }
    // The layer table is staged in a ring of pinned slots: the async H2D copy below reads the slot when it
    // EXECUTES (also on every replay of a captured graph), so consecutive calls must not share a slot.
    if (ws->ring_pos + n_layers > ws->layers_cap) ws->ring_pos = 0;
    slot_begin = ws->ring_pos;
    ws->ring_pos += n_layers;
    LayerRef* hl = ws->h_layers + slot_begin;
    // apply_impl: 0 auto (K-split two-kernel tcgen05 apply for rank pads <= 64, the two-GEMM tcgen05 apply for every other low-rank
    // edit, fp32 SIMT for the dense K x K form), 1 fp32 SIMT (validation twin), 4 apply_tc3.cu (fused one-kernel form), 5 apply_gemm3x.cu,
    // 7 apply_ab.cu
    const bool lowrank = !ws->dense && ws->rank > 0;
    const bool use_ab = choose_ab(ws, n_layers);
    const bool use_tc3 = !use_ab && lowrank && (ws->apply_impl == 4 || (ws->apply_impl == 0 && apply_tc3_available(ws, n_layers)));
    const bool use_g3 = !use_tc3 && lowrank && (ws->apply_impl == 5 || (ws->apply_impl == 0 && apply_gemm3x_available(ws, n_layers)));
    const int tile_rows = use_g3 ? 128 : SG_BM;
    int tiles = 0; bool inplace = false;
    for (int l = 0; l < n_layers; ++l) {
        if (!W_old[l] || !W_new[l] || d[l] <= 0) { set_error("layer %d: null pointer or d <= 0", l); return UCE_E_ARG; }
        inplace |= (W_old[l] == W_new[l]);
    }
    const int ab_ks = use_ab ? apply_ab_ksplit(K) : 1;
    void* slots_h = nullptr; void* slots_d = nullptr;
    int table_launches = 0;
    if (use_ab) {
        std::vector<int> trows(n_layers), tbeg(n_layers);
        tiles = apply_ab_plan(ws->sm_count, ab_ks, d, n_layers, trows.data(), tbeg.data());      // tiles = row blocks
        for (int l = 0; l < n_layers; ++l) hl[l] = LayerRef{W_old[l], W_new[l], d[l], tbeg[l], trows[l]};
        const size_t sb = apply_ab_slot_bytes();
        if (!ws->slots_dev) {
            ws->slots_cap = 32768;
            UCE_CUDA(cudaMalloc(&ws->slots_dev, (size_t)ws->slots_cap * sb));
            UCE_CUDA(cudaMallocHost(&ws->h_slots, (size_t)ws->slots_cap * sb));
        }
        if (tiles > ws->slots_cap) { set_error("too many row blocks per call (%d > %d)", tiles, ws->slots_cap); return UCE_E_STATE; }
        if (stage != 2) {
            if (ws->slots_pos + tiles > ws->slots_cap) ws->slots_pos = 0;      // ring, for the same reason as the layer table
            slots_h = (char*)ws->h_slots + (size_t)ws->slots_pos * sb;
            slots_d = (char*)ws->slots_dev + (size_t)ws->slots_pos * sb;
            apply_ab_fill_slots(slots_h, hl, n_layers);
            { int rc = table_upload(slots_d, slots_h, (size_t)tiles * sb, st, &table_launches); if (rc) return rc; }
            if (stage == 1) ws->slots_staged.push_back(ws->slots_pos);           // stage 2 of the same slice picks this table up (FIFO)
            ws->slots_pos += tiles;
        } else {
            if (ws->slots_staged.empty()) { set_error("apply stage 2 without a matching stage 1"); return UCE_E_STATE; }
            slots_d = (char*)ws->slots_dev + (size_t)ws->slots_staged.front() * sb;
            ws->slots_staged.erase(ws->slots_staged.begin());
        }
    } else if (use_tc3) {
        std::vector<int> trows(n_layers), tbeg(n_layers);
        tiles = apply_tc3_plan(ws->sm_count, d, n_layers, trows.data(), tbeg.data());
        for (int l = 0; l < n_layers; ++l) hl[l] = LayerRef{W_old[l], W_new[l], d[l], tbeg[l], trows[l]};
    } else {
        for (int l = 0; l < n_layers; ++l) {
            hl[l] = LayerRef{W_old[l], W_new[l], d[l], tiles, 0};
            tiles += ceil_div(d[l], tile_rows);
        }
    }
    LayerRef* dl = ws->layers_dev + slot_begin;
    int launches = table_launches;
    if (!use_ab) { int rc = table_upload(dl, hl, n_layers * sizeof(LayerRef), st, &launches); if (rc) return rc; }    // the K-split kernels read slots + tensor maps only
    const int r_pad = ws->rank_pad;
    const bool prof = ws->profile && !no_profile;
    ws->pev_mid = 0;
    if (prof && !use_tc3) UCE_CUDA(cudaEventRecord(ws->pev[2], st));      // apply_tc3 records it itself, right before its launch
    if (ws->rank == 0) {   // no active edit rows: the edit is the identity
        copy_layers_kernel<<<tiles, 256, 0, st>>>(dl, n_layers, K);
        UCE_LAUNCH_CHECK(); ++launches;
        ws->launches_apply = launches;
        if (prof) UCE_CUDA(cudaEventRecord(ws->pev[4], st));
        return 0;
    }
    const size_t need = use_ab ? (size_t)ab_ks * tiles * 128 * r_pad : use_g3 ? (size_t)tiles * 128 * r_pad : use_tc3 ? 0 : (ws->dense ? (inplace ? (size_t)tiles * SG_BM * K : 0) : (size_t)tiles * SG_BM * r_pad);
    // P scratch is shared by successive apply calls; they are ordered on one stream (host path: s_compute)
    if (ws->P_off + need > ws->P_cap) {
        if (ws->P_off != 0 || stage == 2) { set_error("apply scratch too small for a staged / sliced call"); return UCE_E_STATE; }
        if (ws->P) { UCE_CUDA(cudaStreamSynchronize(st)); UCE_CUDA(cudaFree(ws->P)); ws->P = nullptr; }
        ws->P_cap = need + need / 4;
        UCE_CUDA(cudaMalloc(&ws->P, ws->P_cap * sizeof(float)));
    }
    if (!ws->dense) {
        if (use_ab) {
            int rc = apply_ab_lowrank(ws, slots_d, slots_h, tiles, hl, n_layers, st, stage, &launches, prof ? ws->pev[3] : nullptr);
            if (rc) return rc;
            if (prof && stage == 0) ws->pev_mid = 1;
        } else if (use_g3) {
            int rc = apply_gemm3x_highrank(ws, dl, hl, n_layers, tiles, st, &launches, stage);
            if (rc) return rc;
        } else if (use_tc3) {
            int rc = apply_tc3_lowrank(ws, dl, hl, n_layers, tiles, st, &launches, prof ? ws->pev[2] : nullptr);
            if (rc) return rc;
        } else {
            apply_simt_p_kernel<<<dim3(ceil_div(r_pad, SG_BN), tiles), SG_THREADS, 0, st>>>(dl, n_layers, K, r_pad, ws->E, ws->P);
            UCE_LAUNCH_CHECK(); ++launches;
            if (prof) { UCE_CUDA(cudaEventRecord(ws->pev[3], st)); ws->pev_mid = 1; }
            apply_simt_w_kernel<<<dim3(ceil_div(K, SG_BN), tiles), SG_THREADS, 0, st>>>(dl, n_layers, K, r_pad, ws->Qt, ws->P);
            UCE_LAUNCH_CHECK(); ++launches;
        }
    } else {
        apply_simt_dense_kernel<<<dim3(ceil_div(K, SG_BN), tiles), SG_THREADS, 0, st>>>(dl, n_layers, K, ws->Dt, inplace ? ws->P : nullptr);
        UCE_LAUNCH_CHECK(); ++launches;
        if (inplace) {
            add_scratch_kernel<<<tiles, 256, 0, st>>>(dl, n_layers, K, ws->P);
            UCE_LAUNCH_CHECK(); ++launches;
        }
    }
    ws->launches_apply = launches;
    if (prof) UCE_CUDA(cudaEventRecord(ws->pev[4], st));
    return 0;
}

}  // namespace uce
