// Shared helpers for libuce_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include "../../include/uce_b200.h"
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>

namespace uce {

void set_error(const char* fmt, ...);   // defined in uce_api.cu (thread-local message)

#define UCE_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::uce::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                             __FILE__, __LINE__);                                        \
            return (int)_e;                                                              \
        }                                                                                \
    } while (0)

#define UCE_LAUNCH_CHECK()                                                               \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess) {                                                         \
            ::uce::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                        \
            return (int)_e;                                                              \
        }                                                                                \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// One edited projection: a [d,K] weight that is read and a [d,K] weight that is written.
struct LayerRef {
    const float* w_old;
    float*       w_new;
    int          d;          // rows (out_features of attn2.to_k / to_v)
    int          tile_begin; // first row-tile index of this layer in the flattened tile list
    int          tile_rows;  // rows per block of this layer (apply_tc3.cu: chosen per projection by the host planner; 0 elsewhere)
};

}  // namespace uce
