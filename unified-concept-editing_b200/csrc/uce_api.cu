// C ABI of libuce_b200 (include/uce_b200.h).
#include "../../include/uce_b200.h"
#include "uce_ws.h"
#include <chrono>
#include <cstdarg>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <new>

namespace uce {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int apply_tc3_plan(int sm_count, const int* d, int n_layers, int* tile_rows, int* tile_begin);   // apply_tc3.cu
}  // namespace uce
using namespace uce;

extern "C" {

int uce_abi_version(void) { return UCE_B200_ABI_VERSION; }
const char* uce_last_error(void) { return g_err; }

int uce_ws_create(int device, int K, int max_rows, uce_ws** out) {
    if (!out || K <= 0 || max_rows <= 0) { set_error("uce_ws_create: bad argument"); return UCE_E_ARG; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        set_error("uce_ws_create: CUDA device %d not available (%d devices)", device, ndev);
        return UCE_E_NO_DEVICE;
    }
    cudaDeviceProp prop;
    UCE_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("uce_ws_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return UCE_E_NO_DEVICE;
    }
    UCE_CUDA(cudaSetDevice(device));
    uce_ws* ws = new (std::nothrow) uce_ws();
    if (!ws) { set_error("out of host memory"); return UCE_E_STATE; }
    ws->device = device; ws->K = K; ws->max_rows = std::max(max_rows, UCE_RANK_PAD);
    ws->sm_count = prop.multiProcessorCount;
    ws->sys_max = round_up(K, UCE_NB);
    const int mr = ws->max_rows;
    const int rp = round_up(mr, UCE_RANK_PAD);
    ws->layers_cap = 16384;
    if (const char* e = getenv("UCE_APPLY_IMPL")) ws->apply_impl = atoi(e);   // 0 auto, 1 SIMT, 2 tcgen05 (debugging aid)
#define WS_ALLOC(ptr, bytes) do { cudaError_t e_ = cudaMalloc((void**)&(ptr), (bytes)); if (e_ != cudaSuccess) { \
        set_error("cudaMalloc(%zu) failed: %s", (size_t)(bytes), cudaGetErrorString(e_)); uce_ws_destroy(ws); return (int)e_; } } while (0)
#define WS_HOST(ptr, bytes) do { cudaError_t e_ = cudaMallocHost((void**)&(ptr), (bytes)); if (e_ != cudaSuccess) { \
        set_error("cudaMallocHost(%zu) failed: %s", (size_t)(bytes), cudaGetErrorString(e_)); uce_ws_destroy(ws); return (int)e_; } } while (0)
    WS_ALLOC(ws->Cp, (size_t)mr * K * sizeof(float));
    WS_ALLOC(ws->E, (size_t)rp * K * sizeof(float));
    WS_ALLOC(ws->Q, (size_t)rp * K * sizeof(float));
    WS_ALLOC(ws->Qt, (size_t)rp * K * sizeof(float));
    WS_ALLOC(ws->E_hi, (size_t)rp * K * sizeof(float));
    WS_ALLOC(ws->E_lo, (size_t)rp * K * sizeof(float));
    WS_ALLOC(ws->Qt_hi, (size_t)rp * K * sizeof(float));
    WS_ALLOC(ws->Qt_lo, (size_t)rp * K * sizeof(float));
    WS_ALLOC(ws->H, (size_t)ws->sys_max * ws->sys_max * sizeof(double));
    WS_ALLOC(ws->Linv, (size_t)ws->sys_max * UCE_NB * sizeof(double));
    WS_ALLOC(ws->X, (size_t)ws->sys_max * mr * sizeof(double));
    WS_ALLOC(ws->Lsmall, (size_t)(15 * 1024 + 160) * sizeof(double));
    WS_ALLOC(ws->src_idx, (size_t)mr * sizeof(int));
    WS_ALLOC(ws->diag_add, (size_t)mr * sizeof(double));
    WS_ALLOC(ws->flag, sizeof(int));
    WS_ALLOC(ws->layers_dev, (size_t)ws->layers_cap * sizeof(LayerRef));
    WS_HOST(ws->h_src_idx, (size_t)mr * sizeof(int));
    WS_HOST(ws->h_diag_add, (size_t)mr * sizeof(double));
    WS_HOST(ws->h_layers, (size_t)ws->layers_cap * sizeof(LayerRef));
    WS_HOST(ws->h_flag, sizeof(int));
    UCE_CUDA(cudaMemset(ws->flag, 0, sizeof(int)));
    UCE_CUDA(cudaEventCreateWithFlags(&ws->ev_stage, cudaEventDisableTiming));
    *out = ws;
    return 0;
}

int uce_ws_destroy(uce_ws* ws) {
    if (!ws) return 0;
    cudaSetDevice(ws->device);
    cudaDeviceSynchronize();
    void* dev[] = {ws->Cp, ws->Cs64, ws->E, ws->Q, ws->Qt, ws->Dt, ws->E_hi, ws->E_lo, ws->Qt_hi, ws->Qt_lo, ws->H, ws->Hcopy, ws->Linv, ws->Lsmall, ws->X, ws->src_idx,
                   ws->diag_add, ws->flag, ws->P, ws->layers_dev, ws->hostpath_C, ws->hostpath_G, ws->hostpath_W};
    for (void* p : dev) if (p) cudaFree(p);
    if (ws->slots_dev) cudaFree(ws->slots_dev);
    void* host[] = {ws->h_src_idx, ws->h_diag_add, ws->h_layers, ws->h_flag, ws->h_slots};
    for (void* p : host) if (p) cudaFreeHost(p);
    for (auto e : ws->pev) if (e) cudaEventDestroy(e);
    if (ws->ev_stage) cudaEventDestroy(ws->ev_stage);
    for (auto e : ws->ev_h2d) cudaEventDestroy(e);
    for (auto e : ws->ev_done) cudaEventDestroy(e);
    if (ws->s_side) cudaStreamDestroy(ws->s_side);
    for (cudaEvent_t e : {ws->ev_fork, ws->ev_E, ws->ev_A}) if (e) cudaEventDestroy(e);
    if (ws->s_compute) cudaStreamDestroy(ws->s_compute);
    if (ws->s_h2d) cudaStreamDestroy(ws->s_h2d);
    if (ws->s_d2h) cudaStreamDestroy(ws->s_d2h);
    delete ws;
    return 0;
}

int uce_ws_set_apply_impl(uce_ws* ws, int impl) {
    if (!ws) return UCE_E_ARG;
    int prev = ws->apply_impl;
    ws->apply_impl = impl;
    return prev;
}

// Host-only: the row-block plan of the two-block tcgen05 apply for `n_layers` projections of d[l] rows on `sm_count` SMs.
int uce_plan_row_blocks(int sm_count, const int* d, int n_layers, int* block_rows, int* first_cta) {
    if (!d || !block_rows || !first_cta || n_layers <= 0) { set_error("uce_plan_row_blocks: bad argument"); return UCE_E_ARG; }
    for (int l = 0; l < n_layers; ++l) if (d[l] <= 0) { set_error("uce_plan_row_blocks: d[%d] <= 0", l); return UCE_E_ARG; }
    return apply_tc3_plan(sm_count, d, n_layers, block_rows, first_cta);
}

int uce_ws_set_factor_impl(uce_ws* ws, int impl) {
    if (!ws) return UCE_E_ARG;
    int prev = ws->force_general;
    ws->force_general = (impl == 1);
    return prev;
}

int uce_ws_set_debug(uce_ws* ws, int on) {
    if (!ws) return UCE_E_ARG;
    int prev = ws->debug;
    ws->debug = on;
    return prev;
}

int uce_ws_set_profile(uce_ws* ws, int on) {
    if (!ws) return UCE_E_ARG;
    int prev = ws->profile;
    if (on && !ws->pev[0]) {
        UCE_CUDA(cudaSetDevice(ws->device));
        for (auto& e : ws->pev) UCE_CUDA(cudaEventCreate(&e));
    }
    ws->profile = on;
    return prev;
}

int uce_ws_timings(uce_ws* ws, float ms[3]) {
    if (!ws || !ms || !ws->pev[0]) { set_error("uce_ws_timings: profiling was never enabled"); return UCE_E_STATE; }
    UCE_CUDA(cudaSetDevice(ws->device));
    ms[0] = ms[1] = ms[2] = 0.f;
    UCE_CUDA(cudaEventSynchronize(ws->pev[4]));
    UCE_CUDA(cudaEventElapsedTime(&ms[0], ws->pev[0], ws->pev[1]));
    if (ws->pev_mid) {
        UCE_CUDA(cudaEventElapsedTime(&ms[1], ws->pev[2], ws->pev[3]));
        UCE_CUDA(cudaEventElapsedTime(&ms[2], ws->pev[3], ws->pev[4]));
    } else {
        UCE_CUDA(cudaEventElapsedTime(&ms[1], ws->pev[2], ws->pev[4]));
    }
    return 0;
}

static int check_factor_args(uce_ws* ws, const void* C, const void* G, const float* scales, int n_rows, int n_edit) {
    if (!ws || !C || !scales || n_rows <= 0 || n_edit < 0 || n_edit > n_rows || (n_edit > 0 && !G)) {
        set_error("uce_factor: bad argument (n_rows=%d n_edit=%d)", n_rows, n_edit);
        return UCE_E_ARG;
    }
    if (n_rows > ws->max_rows) {
        set_error("uce_factor: n_rows=%d exceeds workspace max_rows=%d", n_rows, ws->max_rows);
        return UCE_E_STATE;
    }
    return 0;
}

int uce_factor_dev_f32(uce_ws* ws, const float* C, const float* G, const float* scales_host, int n_rows, int n_edit,
                       float lamb, void* stream) {
    int rc = check_factor_args(ws, C, G, scales_host, n_rows, n_edit);
    if (rc) return rc;
    UCE_CUDA(cudaSetDevice(ws->device));
    if (ws->profile) UCE_CUDA(cudaEventRecord(ws->pev[0], (cudaStream_t)stream));
    rc = factor_dev(ws, C, G, scales_host, n_rows, n_edit, lamb, (cudaStream_t)stream);
    if (ws->profile) UCE_CUDA(cudaEventRecord(ws->pev[1], (cudaStream_t)stream));
    return rc;
}

int uce_apply_dev_f32(uce_ws* ws, const float* const* W_old, float* const* W_new, const int* d, int n_layers, void* stream) {
    if (!ws || !W_old || !W_new || !d || n_layers <= 0) { set_error("uce_apply: bad argument"); return UCE_E_ARG; }
    UCE_CUDA(cudaSetDevice(ws->device));
    return apply_dev(ws, W_old, W_new, d, n_layers, (cudaStream_t)stream, false);
}

// Fork / join of the K-split apply's first kernel around the factor (see the header).
struct Stage1Hook { uce_ws* ws; const float* const* W_old; float* const* W_new; const int* d; int n_layers; int launches; int ran; };
static int stage1_hook(void* p) {
    Stage1Hook* h = (Stage1Hook*)p;
    uce_ws* ws = h->ws;
    if (!(ws->ev_E_recorded && ws->rank > 0 && apply_stage_split(ws, h->n_layers))) return 0;
    UCE_CUDA(cudaStreamWaitEvent(ws->s_side, ws->ev_fork, 0));
    UCE_CUDA(cudaStreamWaitEvent(ws->s_side, ws->ev_E, 0));
    int rc = apply_dev(ws, h->W_old, h->W_new, h->d, h->n_layers, ws->s_side, true, 1);
    if (rc) return rc;
    h->launches = ws->launches_apply;
    UCE_CUDA(cudaEventRecord(ws->ev_A, ws->s_side));
    ws->hook_done = ws->ev_A;
    h->ran = 1;
    return 0;
}

static int edit_dev(uce_ws* ws, const float* C, const float* G, const float* scales, int n_rows, int n_edit, float lamb,
                    const float* const* W_old, float* const* W_new, const int* d, int n_layers, cudaStream_t st, bool no_profile) {
    if (!ws->s_side) {
        UCE_CUDA(cudaStreamCreateWithFlags(&ws->s_side, cudaStreamNonBlocking));
        UCE_CUDA(cudaEventCreateWithFlags(&ws->ev_fork, cudaEventDisableTiming));
        UCE_CUDA(cudaEventCreateWithFlags(&ws->ev_E, cudaEventDisableTiming));
        UCE_CUDA(cudaEventCreateWithFlags(&ws->ev_A, cudaEventDisableTiming));
    }
    const bool prof = ws->profile && !no_profile;
    // profiling brackets each kernel with events on `stream`: everything stays serial there, so that the durations are per kernel
    const bool overlap = !prof && getenv("UCE_NO_OVERLAP") == nullptr;
    if (overlap) UCE_CUDA(cudaEventRecord(ws->ev_fork, st));      // W_old is ready here (stream order of the caller)
    ws->want_ev_E = overlap ? 1 : 0;
    ws->ev_E_recorded = 0;
    ws->slots_staged.clear();
    if (prof) UCE_CUDA(cudaEventRecord(ws->pev[0], st));
    // single-CTA factor: the factor itself calls back (stage1_hook) once its long kernel is launched, so that the apply's first kernel is
    // on its way early and this stream waits for it before the factor's LAST kernels — the apply's second kernel then follows solve_emit
    // directly and is launched programmatically beside it
    Stage1Hook hook{ws, W_old, W_new, d, n_layers, 0, 0};
    if (overlap) { ws->hook_after_E = stage1_hook; ws->hook_ctx = &hook; }
    ws->hook_done = nullptr;
    int rc = factor_dev(ws, C, G, scales, n_rows, n_edit, lamb, st);
    ws->want_ev_E = 0;
    ws->hook_after_E = nullptr; ws->hook_ctx = nullptr; ws->hook_done = nullptr;
    if (prof) UCE_CUDA(cudaEventRecord(ws->pev[1], st));
    if (rc) return rc;
    if (hook.ran) {
        rc = apply_dev(ws, W_old, W_new, d, n_layers, st, true, 2);
        ws->launches_apply += hook.launches;
        ws->pev_mid = 0;
        return rc;
    }
    if (overlap && ws->ev_E_recorded && ws->rank > 0 && apply_stage_split(ws, n_layers)) {
        UCE_CUDA(cudaStreamWaitEvent(ws->s_side, ws->ev_fork, 0));
        UCE_CUDA(cudaStreamWaitEvent(ws->s_side, ws->ev_E, 0));
        rc = apply_dev(ws, W_old, W_new, d, n_layers, ws->s_side, true, 1);
        const int la = ws->launches_apply;
        if (rc) return rc;
        UCE_CUDA(cudaEventRecord(ws->ev_A, ws->s_side));
        UCE_CUDA(cudaStreamWaitEvent(st, ws->ev_A, 0));
        if (prof) UCE_CUDA(cudaEventRecord(ws->pev[2], st));
        rc = apply_dev(ws, W_old, W_new, d, n_layers, st, true, 2);
        ws->launches_apply += la;
        ws->pev_mid = 0;
        if (prof) UCE_CUDA(cudaEventRecord(ws->pev[4], st));
        return rc;
    }
    return apply_dev(ws, W_old, W_new, d, n_layers, st, no_profile, 0);      // records pev[2..4] itself when profiling
}

int uce_edit_dev_f32(uce_ws* ws, const float* C, const float* G, const float* scales_host, int n_rows, int n_edit, float lamb,
                     const float* const* W_old, float* const* W_new, const int* d, int n_layers, void* stream) {
    int rc = check_factor_args(ws, C, G, scales_host, n_rows, n_edit);
    if (rc) return rc;
    if (!W_old || !W_new || !d || n_layers <= 0) { set_error("uce_edit_dev: bad layer arguments"); return UCE_E_ARG; }
    UCE_CUDA(cudaSetDevice(ws->device));
    return edit_dev(ws, C, G, scales_host, n_rows, n_edit, lamb, W_old, W_new, d, n_layers, (cudaStream_t)stream, false);
}

int uce_ws_check(uce_ws* ws, void* stream) {
    if (!ws) return UCE_E_ARG;
    UCE_CUDA(cudaSetDevice(ws->device));
    cudaStream_t st = (cudaStream_t)stream;
    UCE_CUDA(cudaMemcpyAsync(ws->h_flag, ws->flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    UCE_CUDA(cudaStreamSynchronize(st));
    if (*ws->h_flag != 0) {
        set_error("Cholesky failed in block %d: lamb*I + C^T S C is not positive definite", *ws->h_flag - 1);
        return UCE_E_NOT_SPD;
    }
    return 0;
}

int uce_ws_info(uce_ws* ws, int* mode, int* rank, int* dense, int* sys_n, int* launches_factor, int* launches_apply) {
    if (!ws) return UCE_E_ARG;
    if (mode) *mode = ws->mode;
    if (rank) *rank = ws->rank;
    if (dense) *dense = ws->dense;
    if (sys_n) *sys_n = ws->sys_n;
    if (launches_factor) *launches_factor = ws->launches_factor;
    if (launches_apply) *launches_apply = ws->launches_apply;
    return 0;
}

int uce_ws_debug_read(uce_ws* ws, int which, void* out, size_t cap) {
    if (!ws || !out) return UCE_E_ARG;
    UCE_CUDA(cudaSetDevice(ws->device));
    UCE_CUDA(cudaDeviceSynchronize());
    const size_t n = ws->sys_n, K = ws->K;
    const void* src = nullptr; size_t bytes = 0;
    switch (which) {
        case 0: src = ws->Hcopy; bytes = n * n * sizeof(double); break;
        case 1: src = ws->H; bytes = n * n * sizeof(double); break;
        case 2: src = ws->Q; bytes = (size_t)ws->rank * K * sizeof(float); break;
        case 3: src = ws->E; bytes = (size_t)ws->rank * K * sizeof(float); break;
        case 4: src = ws->Dt; bytes = K * K * sizeof(float); break;
        default: set_error("debug_read: unknown selector %d", which); return UCE_E_ARG;
    }
    if (!src) { set_error("debug_read: buffer %d not populated (set debug before factor?)", which); return UCE_E_STATE; }
    if (bytes > cap) { set_error("debug_read: need %zu bytes, have %zu", bytes, cap); return UCE_E_ARG; }
    UCE_CUDA(cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

// Whole edit from host buffers: copies pipelined against compute in groups of layers.
int uce_edit_host_f32(uce_ws* ws, const float* C, const float* G, const float* scales, int n_rows, int n_edit, float lamb,
                      const float* const* W_old, float* const* W_new, const int* d, int n_layers) {
    int rc = check_factor_args(ws, C, G, scales, n_rows, n_edit);
    if (rc) return rc;
    if (!W_old || !W_new || !d || n_layers <= 0 || n_layers > ws->layers_cap / 4) { set_error("uce_edit_host: bad layer arguments"); return UCE_E_ARG; }
    UCE_CUDA(cudaSetDevice(ws->device));
    const int K = ws->K;
    if (!ws->s_compute) {
        UCE_CUDA(cudaStreamCreateWithFlags(&ws->s_compute, cudaStreamNonBlocking));
        UCE_CUDA(cudaStreamCreateWithFlags(&ws->s_h2d, cudaStreamNonBlocking));
        UCE_CUDA(cudaStreamCreateWithFlags(&ws->s_d2h, cudaStreamNonBlocking));
        UCE_CUDA(cudaMalloc(&ws->hostpath_C, (size_t)ws->max_rows * K * sizeof(float)));
        UCE_CUDA(cudaMalloc(&ws->hostpath_G, (size_t)ws->max_rows * K * sizeof(float)));
    }
    size_t total = 0;
    for (int l = 0; l < n_layers; ++l) {
        if (d[l] <= 0 || !W_old[l] || !W_new[l]) { set_error("uce_edit_host: layer %d invalid", l); return UCE_E_ARG; }
        total += (size_t)d[l] * K;
    }
    if (total > ws->hostpath_W_cap) {
        if (ws->hostpath_W) { UCE_CUDA(cudaDeviceSynchronize()); UCE_CUDA(cudaFree(ws->hostpath_W)); ws->hostpath_W = nullptr; }
        UCE_CUDA(cudaMalloc(&ws->hostpath_W, total * sizeof(float)));
        ws->hostpath_W_cap = total;
    }
    // groups of layers of roughly equal bytes: enough groups to overlap H2D(g+1) / apply(g) / D2H(g-1)
    // (8 or 24 by default, see below; UCE_HOST_GROUPS overrides it for measurements: more groups shorten the pipeline's fill and drain — the first upload
    // and the last download run alone on the link — at the price of more, smaller apply launches)
    // (measured, cfg2: weights in one pinned arena per direction — one copy per group — 6 or 8 groups 2.07 ms, 24 groups 2.47; one pinned
    //  tensor per projection — the copies are per projection anyway — 8 groups 2.33 ms, 24 groups 2.18: profiles/r02_e2e_pipeline.txt)
    bool contiguous = true;
    for (int l = 1; l < n_layers && contiguous; ++l)
        contiguous = W_old[l] == W_old[l - 1] + (size_t)d[l - 1] * K && W_new[l] == W_new[l - 1] + (size_t)d[l - 1] * K;
    int target_groups = std::min(n_layers, contiguous ? 8 : 24);
    if (const char* e = getenv("UCE_HOST_GROUPS")) { const int t = atoi(e); if (t >= 1) target_groups = std::min(n_layers, t); }
    // Group sizes are graded: the first upload and the last apply + download run alone on the link (pipeline fill and drain), so the
    // first group is ~0.65 and the last two ~0.8 / ~0.4 of the average — the first apply cannot start before the factor ends anyway
    // (~0.12 ms: about 6 MB of upload) — and the groups in between share the rest.  UCE_HOST_EVEN_GROUPS=1 restores equal groups.
    static const bool even_groups = [] { const char* e = getenv("UCE_HOST_EVEN_GROUPS"); return e && atoi(e) != 0; }();
    std::vector<double> share(target_groups, 1.0);
    if (!even_groups && target_groups >= 4) {
        share[0] = 0.65; share[target_groups - 1] = 0.4; share[target_groups - 2] = 0.8;
        const double rest = (target_groups - 1.85) / (target_groups - 3);
        for (int g = 1; g < target_groups - 2; ++g) share[g] = rest;
    }
    std::vector<int> gbeg{0};
    {
        double want = 0.0; size_t acc = 0;
        for (int l = 0, g = 0; l < n_layers; ++l) {
            acc += (size_t)d[l] * K;
            const double bound = (want + share[g]) / target_groups * (double)total;
            // close the group at the layer boundary nearest to its share
            const size_t next = l + 1 < n_layers ? (size_t)d[l + 1] * K : 0;
            if (l + 1 < n_layers && g + 1 < target_groups && (double)acc + 0.5 * (double)next >= bound) { gbeg.push_back(l + 1); want += share[g]; ++g; }
        }
    }
    gbeg.push_back(n_layers);
    const int ng = (int)gbeg.size() - 1;
    while ((int)ws->ev_h2d.size() < ng + 1) {
        cudaEvent_t a, b;
        UCE_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        UCE_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        ws->ev_h2d.push_back(a); ws->ev_done.push_back(b);
    }
    // concept rows first, then the factor on the compute stream while the weights stream in
    UCE_CUDA(cudaMemcpyAsync(ws->hostpath_C, C, (size_t)n_rows * K * sizeof(float), cudaMemcpyHostToDevice, ws->s_h2d));
    if (n_edit > 0) UCE_CUDA(cudaMemcpyAsync(ws->hostpath_G, G, (size_t)n_edit * K * sizeof(float), cudaMemcpyHostToDevice, ws->s_h2d));
    UCE_CUDA(cudaEventRecord(ws->ev_h2d[ng], ws->s_h2d));
    UCE_CUDA(cudaStreamWaitEvent(ws->s_compute, ws->ev_h2d[ng], 0));
    rc = factor_dev(ws, ws->hostpath_C, ws->hostpath_G, scales, n_rows, n_edit, lamb, ws->s_compute);
    if (rc) { cudaDeviceSynchronize(); return rc; }
    std::vector<const float*> po(n_layers); std::vector<float*> pn(n_layers);
    { size_t off = 0; for (int l = 0; l < n_layers; ++l) { po[l] = ws->hostpath_W + off; pn[l] = ws->hostpath_W + off; off += (size_t)d[l] * K; } }
    int launches = 0;
    // (layers whose host buffers follow each other in memory — a caller that keeps its weights in one pinned arena — travel as ONE copy per
    //  group: 32 per-projection copies in each direction cost the link 2.04 ms where two flat copies take 1.57, profiles/r01_e2e_floor.txt)
    auto run_end = [&](const float* const* host, int l, int end) {
        int e = l + 1;
        while (e < end && host[e] == host[e - 1] + (size_t)d[e - 1] * K) ++e;
        return e;
    };
    auto run_bytes = [&](int l, int e) { size_t b = 0; for (int i = l; i < e; ++i) b += (size_t)d[i] * K * sizeof(float); return b; };
    auto upload = [&](int g) -> int {
        for (int l = gbeg[g]; l < gbeg[g + 1];) {
            const int e = run_end(W_old, l, gbeg[g + 1]);
            UCE_CUDA(cudaMemcpyAsync(pn[l], W_old[l], run_bytes(l, e), cudaMemcpyHostToDevice, ws->s_h2d));
            l = e;
        }
        UCE_CUDA(cudaEventRecord(ws->ev_h2d[g], ws->s_h2d));
        return 0;
    };
    // Order of submission matters: the apply of group g uploads its (tiny) block tables with an async copy of its own, and the copy engine
    // serves host-to-device copies in submission order — tables submitted after ALL weight uploads would wait for all of them, and the
    // downloads could not start before the last upload ended (observed: 2.9-3.2 ms instead of 2.1).  So a group's tables are submitted
    // right after that group's weights, one group ahead of the next upload.
    static const bool hoist = [] { const char* e = getenv("UCE_HOST_HOIST_UPLOADS"); return e && atoi(e) != 0; }();
    // UCE_HOST_TRACE=1: device timeline of one call (timing events around every upload, apply and download), printed to stderr
    static const bool trace = [] { const char* e = getenv("UCE_HOST_TRACE"); return e && atoi(e) != 0; }();
    std::vector<cudaEvent_t> tev;
    auto stamp = [&](cudaStream_t s) { if (!trace) return; cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); tev.push_back(e); };
    if (hoist) for (int g = 0; g < ng; ++g) { rc = upload(g); if (rc) return rc; }
    std::vector<double> host_t;
    auto host_now = [&] { if (trace) host_t.push_back(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count()); };
    for (int g = 0; g < ng; ++g) {
        host_now();
        stamp(ws->s_h2d);
        if (!hoist) { rc = upload(g); if (rc) return rc; }
        stamp(ws->s_h2d);
        UCE_CUDA(cudaStreamWaitEvent(ws->s_compute, ws->ev_h2d[g], 0));
        stamp(ws->s_compute);
        host_now();
        rc = apply_dev(ws, po.data() + gbeg[g], pn.data() + gbeg[g], d + gbeg[g], gbeg[g + 1] - gbeg[g], ws->s_compute, true);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        host_now();
        stamp(ws->s_compute);
        launches += ws->launches_apply;
        UCE_CUDA(cudaEventRecord(ws->ev_done[g], ws->s_compute));
        UCE_CUDA(cudaStreamWaitEvent(ws->s_d2h, ws->ev_done[g], 0));
        stamp(ws->s_d2h);
        for (int l = gbeg[g]; l < gbeg[g + 1];) {
            const int e = run_end(W_new, l, gbeg[g + 1]);
            UCE_CUDA(cudaMemcpyAsync(W_new[l], pn[l], run_bytes(l, e), cudaMemcpyDeviceToHost, ws->s_d2h));
            l = e;
        }
        stamp(ws->s_d2h);
    }
    if (trace) {
        cudaDeviceSynchronize();
        for (int g = 0; g < ng; ++g) {
            float t[6];
            for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&t[i], tev[0], tev[6 * g + i]);
            fprintf(stderr, "host-path group %2d: upload %7.3f-%7.3f  apply %7.3f-%7.3f  download %7.3f-%7.3f ms | host: iteration starts at %7.1f us, apply_dev encodes %6.1f-%6.1f us\n",
                    g, t[0], t[1], t[2], t[3], t[4], t[5], host_t[3 * g] - host_t[0], host_t[3 * g + 1] - host_t[0], host_t[3 * g + 2] - host_t[0]);
        }
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    ws->launches_apply = launches;
    UCE_CUDA(cudaStreamSynchronize(ws->s_d2h));
    UCE_CUDA(cudaStreamSynchronize(ws->s_compute));
    return uce_ws_check(ws, ws->s_compute);
}

}  // extern "C"
