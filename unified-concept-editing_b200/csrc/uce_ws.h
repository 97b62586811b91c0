// Internal definition of the opaque workspace behind include/uce_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "uce_common.cuh"

constexpr int UCE_NB = 32;        // Cholesky block size (fp64)
constexpr int UCE_RANK_PAD = 32;  // rank of E/Q is padded to a multiple of this (UMMA N granularity)

struct uce_ws {
    int device = 0, K = 0, max_rows = 0;
    int sm_count = 0;
    int sys_max = 0;      // largest padded linear system (round_up(K, NB))
    // ---- state of the last factor ----
    int mode = 0;         // 0 none, 1 dual (n x n), 2 primal (K x K)
    int n_act = 0, n_edit = 0, n_pres = 0;   // active (scale != 0) rows
    int sys_n = 0;        // padded system size
    int rank = 0, rank_pad = 0;
    int dense = 0;
    float lamb = 0.f;
    int launches_factor = 0, launches_apply = 0;
    int apply_impl = 0;   // 0 auto, 1 fp32 SIMT, 4 two-block tcgen05 (rank pad <= 64), 5 two-GEMM tcgen05 (any rank)
    int debug = 0;
    int force_general = 0;   // 1: always use the general blocked factor (testing)
    int profile = 0;
    cudaEvent_t pev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // factor begin/end, apply begin/mid/end
    int pev_mid = 0;
    // ---- device buffers ----
    float*  Cp = nullptr;     // [max_rows, K] active concept rows, internal order (preserve first, edit last)
    double* Cs64 = nullptr;   // [max_rows, K] s_r * Cp (primal only, lazily allocated)
    float*  E = nullptr;      // [rank_pad, K]  G_e - C_e  (zero padded rows)
    float*  Q = nullptr;      // [rank_pad, K]
    float*  Qt = nullptr;     // [K, rank_pad]
    float*  Dt = nullptr;     // [K, K]   dense factor (lazily allocated)
    float  *E_hi = nullptr, *E_lo = nullptr, *Qt_hi = nullptr, *Qt_lo = nullptr;   // tf32 hi/lo splits for the tcgen05 apply
    double* H = nullptr;      // [sys_max, sys_max]
    double* Hcopy = nullptr;  // debug copy of the assembled system (lazily allocated)
    double* Linv = nullptr;   // [sys_max/NB][NB][NB]
    double* Lsmall = nullptr; // low-latency factor: block triangle of L (15 blocks of 32 x 32) + 1 / L_ii (160), chol_small -> solve_emit
    int     H_dirty = 1;          // 0: the part of H the single-CTA factor uses is known to be zero (that kernel clears what it read)
    // uce_edit_dev_f32: called by the factor right after the kernel that follows the E rows has been launched — launches the apply's
    // first kernel on the side stream; the factor then waits for hook_done before the kernels whose successor is the apply's second kernel
    int   (*hook_after_E)(void*) = nullptr;
    void*   hook_ctx = nullptr;
    cudaEvent_t hook_done = nullptr;
    double* X = nullptr;      // [sys_max, max_rows]  rhs / solution
    int*    src_idx = nullptr;   // [max_rows] API row of internal row r
    double* diag_add = nullptr;  // [max_rows] lamb / s_r  (dual)  or s_r (primal)
    int*    flag = nullptr;      // device: 0 ok, else 1 + failing block
    float*  P = nullptr;         // apply scratch [rows_pad_total, rank_pad]
    size_t  P_cap = 0;
    size_t  P_off = 0;           // floats: the slice of the scratch the current (sliced) apply call uses
    uce::LayerRef* layers_dev = nullptr;
    int layers_cap = 0;
    int ring_pos = 0;
    void* slots_dev = nullptr;   // row-block table of the K-split apply (apply_ab.cu), staged like the layer table
    void* h_slots = nullptr;
    int slots_cap = 0, slots_pos = 0;
    std::vector<int> slots_staged;   // tables uploaded by stage-1 calls, waiting for their stage-2 call
    int stage_pending = 0;
    cudaEvent_t ev_stage = nullptr;   // marks consumption of the factor's pinned staging
    // ---- pinned host staging ----
    int*    h_src_idx = nullptr;
    double* h_diag_add = nullptr;
    uce::LayerRef* h_layers = nullptr;
    int*    h_flag = nullptr;
    // ---- host-buffer path (uce_edit_host_f32) ----
    cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaStream_t s_side = nullptr;                     // kernel A of the K-split apply runs here while the factor computes Q (uce_edit_dev_f32)
    cudaEvent_t ev_fork = nullptr, ev_E = nullptr, ev_A = nullptr;
    int want_ev_E = 0, ev_E_recorded = 0;              // factor_dev records ev_E once E / E_hi / E_lo are complete (and the Gram kernel is enqueued)
    float* hostpath_C = nullptr; float* hostpath_G = nullptr;
    float* hostpath_W = nullptr; size_t hostpath_W_cap = 0;   // device staging for all layers (in place)
    std::vector<cudaEvent_t> ev_h2d, ev_done;
};

namespace uce {
// factor.cu
int factor_dev(uce_ws* ws, const float* C, const float* G, const float* scales_host, int n_rows, int n_edit,
               float lamb, cudaStream_t st);
// apply.cu
// stage 0: the whole apply on `st`.  K-split apply only (apply_stage_split): stage 1 = partial products (needs E), stage 2 = update (needs Q)
int apply_dev(uce_ws* ws, const float* const* W_old, float* const* W_new, const int* d, int n_layers,
              cudaStream_t st, bool no_profile = false, int stage = 0);
// Small host table -> device memory through kernel parameters (no copy engine; see apply.cu).  `bytes` a multiple of 4.
int table_upload(void* dst, const void* src_host, size_t bytes, cudaStream_t st, int* launches);
bool apply_stage_split(const uce_ws* ws, int n_layers);   // true when apply_dev would take the two-kernel K-split path for this edit
// factor_small.cu
bool factor_small_applicable(const uce_ws* ws, int n, int n_edit, bool dual);
int potrf_inv_general(double* H, int ld, int kb, double* Linv, int* flag, cudaStream_t st);     // factor_small.cu: one diagonal block of the general path
int factor_small(uce_ws* ws, const float* C, const float* G, int n, int n_pres, int n_edit, cudaStream_t st, int* launches);
}  // namespace uce
