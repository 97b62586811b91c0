// Launchers of the memory-bound U-Net kernels (unet_ops.cu). All return 0 or a cudaError_t.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>

namespace uce {
// one time-embedding projection (ResnetBlock2D.time_emb_proj): out[NB, cout] fp32 = st_emb[NB, K] . w[cout, K]^T + bias
struct TembJob { const __nv_bfloat16* w; const float* bias; float* out; int cout; int first_channel; };
int op_temb_proj_all(const TembJob* jobs_dev, int n_jobs, int total_channels, const __nv_bfloat16* st_emb, int NB, int K, cudaStream_t st);
// GroupNorm(+SiLU): `ws` = this call's own op_groupnorm_ws_floats(NB, G) floats, zeroed once at allocation (statistics are bit-reproducible)
size_t op_groupnorm_ws_floats(int NB, int G);
int op_groupnorm(const __nv_bfloat16* x, __nv_bfloat16* y, int NB, int HW, int C, int G, float* ws, const float* gamma, const float* beta,
                 float eps, int silu, cudaStream_t st);
int op_layernorm(const __nv_bfloat16* x, __nv_bfloat16* y, long rows, int C, const float* gamma, const float* beta, float eps, cudaStream_t st);
int op_softmax(const float* S, long lds, __nv_bfloat16* P, long ldp, long rows, int Lk, cudaStream_t st);
int op_geglu(const __nv_bfloat16* x, __nv_bfloat16* y, long rows, int Hd, cudaStream_t st);
int op_silu(const __nv_bfloat16* x, __nv_bfloat16* y, long n, cudaStream_t st);
int op_upsample2x(const __nv_bfloat16* x, __nv_bfloat16* y, int NB, int H, int W, int C, cudaStream_t st);
int op_concat_c(const __nv_bfloat16* a, const __nv_bfloat16* b, __nv_bfloat16* y, long pixels, int C1, int C2, cudaStream_t st);
int op_conv_in(const float* x, const float* w, const float* bias, __nv_bfloat16* y, int NB, int H, int W, int Cout, cudaStream_t st);
int op_conv_out(const __nv_bfloat16* x, const float* w, const float* bias, float* y, int NB, int H, int W, int Cin, cudaStream_t st);
int op_timestep_embedding(float t, int dim, int NB, __nv_bfloat16* out, cudaStream_t st);
int op_cfg_step(const float* eps2, long n, float gs, float* eps_out, const float* h1, const float* h2, const float* h3, float c0, float c1,
                float c2, float c3, float cx, float ce, const float* x_in, float* x_out, cudaStream_t st);
}  // namespace uce
