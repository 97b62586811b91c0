// Descriptor of one bf16 tcgen05 GEMM / implicit-GEMM convolution launch (see unet_gemm.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstring>

namespace uce {

struct GemmDesc {
    CUtensorMap tmA, tmB;      // A: linear {K, M, b1, b2} or conv {C, W, H, N};  B: {K, N, b1, b2}
    CUtensorMap tmO, tmR, tmP; // TMA epilogue (pair kernel): output, residual (bf16 [128 x 64] boxes), split-K slabs (fp32 [128 x 32])
    int M, N, Kd;              // per-batch problem (conv: M = NB*Ho*Wo, Kd = taps*Cin)
    int batch, b1cnt;          // grid.z = batch; z -> (b1 = z % b1cnt, b2 = z / b1cnt)
    int a_batched, b_batched;  // whether the operand's tensor map is indexed by (b1, b2)
    int conv, taps, cin, stride, pad;
    int tw, th, tn, Wo, Ho, NBimg;
    int m_tiles;               // grid.y
    int tile_rows;             // valid rows of an A tile (conv rectangles smaller than 128 pixels), else 128
    int a_bytes;               // bytes one A-tile TMA delivers (expect_tx)
    int rows_per_img;          // linear: image index of a row = row / rows_per_img (for rowbias); 0 = none
    int katoms;                // 64-wide k atoms per pipeline stage (1 or 2): 2 halves the barrier round trips of the single MMA thread
    int stages;                // TMA->MMA pipeline depth (3: two CTAs per SM; 6: one CTA per SM, hides the TMA latency of under-filled grids)
    int ksplit;                // > 1: the k loop is split over grid.z; CTA z stores its partial sums into slab z of splitk_ws (fp32 [ksplit,M,N])
    float* splitk_ws;          // shared scratch; the finalize pass sums the slabs in order and applies the epilogue
    int tma_epi;               // 0 register epilogue, 1 bf16 output through TMA (residual prefetched by TMA), 2 fp32 split-K slabs through TMA
    int pair, bn, tmem_cols;   // pair = 1: CTA-pair kernel (cta_group::2) on 256 x bn tiles, each CTA staging bn/2 rows of B
    const void* b_ptr; long b_ld;   // B operand as given to gemm_desc_* (gemm_enable_pair re-encodes its tensor map)
    float alpha;
    void* out; int out_fp32; long ldo, out_b1_stride, out_b2_stride;
    const float* bias;         // [N]
    const float* rowbias;      // [images][N]
    const __nv_bfloat16* residual; long ldr, res_b1_stride, res_b2_stride;
};

// A [b2][b1][M][Kd] via strides (elements), B likewise; batch = b1cnt * b2cnt.
int gemm_desc_linear(GemmDesc* g, const void* A, long lda, long a_b1_stride, long a_b2_stride, const void* B, long ldb,
                     long b_b1_stride, long b_b2_stride, int M, int N, int Kd, int b1cnt, int b2cnt, int a_batched, int b_batched);
int gemm_desc_conv(GemmDesc* g, const void* act_nhwc, int NB, int Hin, int Win, int Cin, const void* w_tapmajor, int Cout,
                   int ksize, int stride);
int gemm_launch(const GemmDesc& g, cudaStream_t st);
// Returns 1 when the descriptor was switched to the CTA-pair kernel, 0 when not applicable, < 0 on error. Call before
// gemm_choose_ksplit / gemm_choose_stages.
int gemm_enable_pair(GemmDesc* g);
// Call after out / residual / ksplit / splitk_ws are final (pair kernel only).
int gemm_enable_tma_epilogue(GemmDesc* g);
// Picks a split factor for under-filled grids (batch == 1 only); returns it (1 = no split).
int gemm_choose_ksplit(const GemmDesc& g, int sm_count);
int gemm_choose_stages(const GemmDesc& g, int sm_count, int* katoms);

}  // namespace uce
