// Fused (flash-style) attention launch descriptor (unet_attn.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstring>

namespace uce {

struct AttnDesc {
    CUtensorMap tmQ, tmK, tmV;     // Q {dhp, L, heads, NB}; K {dhp, Lk, heads, NB}; V^T {Lk, dhp, heads, NB}
    int NB, heads, dhp, L, Lk;
    float scale;                   // 1 / sqrt(true head dim)
    __nv_bfloat16* out; long ldo;  // [NB, L, heads*dhp]
};

bool attn_fused_supported(int dhp);
// ldq / ldk: row strides (elements) of q and k — heads * dhp when 0; a merged [q | k] projection passes 2 * heads * dhp for both
int attn_desc_make(AttnDesc* g, const void* q, const void* k, const void* vt, void* out, int NB, int heads, int dhp, long L, int Lk, int Lkp,
                   float scale, long ldq = 0, long ldk = 0);
int attn_launch(const AttnDesc& g, cudaStream_t st);

}  // namespace uce
