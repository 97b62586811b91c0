/*
 * sd_unet_b200.h — C ABI of the B200-native Stable Diffusion v1.x U-Net denoise step (libuce_b200.so).
 *
 * Replaces what the reference reaches through diffusers inside `pipe(...)`:
 *   evalscripts/generate-images-sd.py:37-42 and trainscripts/uce_sd_debias.py:22-26  ->
 *   UNet2DConditionModel.forward (eps prediction), classifier-free guidance and the scheduler update
 *   (SURVEY.md §8 rows a10/a11, Appendix A/B).  Weights use the diffusers state-dict names, so the edited
 *   attn2.to_k/to_v tensors written by the edit solver (trainscripts/uce_sd_erase.py:85-88) load by name, exactly
 *   like `pipe.unet.load_state_dict(..., strict=False)` (generate-images-sd.py:17-19).
 *
 * Conventions: plain C; device pointers unless stated; status 0 = ok, <0 = SD_E_*, >0 = cudaError_t;
 * sd_last_error() describes the last failure; all work is enqueued on the caller's stream.
 */
#ifndef SD_UNET_B200_H
#define SD_UNET_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_E_ARG      (-1)
#define SD_E_STATE    (-2)
#define SD_E_WEIGHT   (-3)   /* unknown name / wrong shape / missing at finalize */
#define SD_E_DEVICE   (-4)

typedef struct sd_unet sd_unet;

typedef struct sd_unet_config {
    int in_channels, out_channels;
    int n_levels;                 /* length of block_out_channels (<= 4) */
    int block_out_channels[4];
    int layers_per_block;
    int down_has_attn[4], up_has_attn[4];
    int cross_attention_dim, context_len;     /* 768, 77 for SD-1.x */
    int heads, norm_groups, temb_dim;
} sd_unet_config;

const char *sd_last_error(void);

/* Engine for `batch` samples per U-Net call (2 x images with classifier-free guidance) at latent size H x W. */
int sd_unet_create(int device, const sd_unet_config *cfg, int batch, int H, int W, sd_unet **out);
int sd_unet_destroy(sd_unet *u);

/* Upload one parameter by its diffusers state-dict name (fp32, HOST pointer, row-major `shape`).  May be called
 * again after finalize to overwrite a parameter in place (edited attn2.to_k / to_v weights). */
int sd_unet_set_weight(sd_unet *u, const char *name, const float *data, const long *shape, int ndim);

/* Allocate activations, build the kernel schedule and the TMA tensor maps.  Fails if a parameter is missing. */
int sd_unet_finalize(sd_unet *u);

/* Text context of the following steps: ctx[batch,context_len,cross_attention_dim] fp32, DEVICE pointer.  Computes the
 * cross-attention K / V^T projections of every transformer block once — the reference's pipe(...) evaluates them at each of
 * its num_inference_steps U-Net calls although the prompt embedding is constant over a row (generate-images-sd.py:37-42).
 * Overwriting a parameter afterwards (sd_unet_set_weight) makes the next forward recompute them from the stored context. */
int sd_unet_set_context(sd_unet *u, const float *ctx, void *stream);

/* eps[batch,4,H,W] (fp32, NCHW) = UNet(x[batch,4,H,W] fp32 NCHW, t, ctx[batch,context_len,cross_attention_dim] fp32).
 * ctx may be NULL after sd_unet_set_context: the step then reuses the cached context projections. */
int sd_unet_forward(sd_unet *u, const float *x, float t, const float *ctx, float *eps, void *stream);

/* Update the timestep a CUDA graph captured around sd_unet_forward will read on its next replay (the captured
 * forward copies t from a pinned host slot; this writes that slot). */
int sd_unet_set_timestep(sd_unet *u, float t);

/* Fused classifier-free guidance + scheduler update on fp32 NCHW latents of n = images*4*H*W elements:
 *   eps = eps_u + gs (eps_t - eps_u)  with eps2 = [uncond | text];  e = c[0] eps + c[1] h1 + c[2] h2 + c[3] h3;
 *   x_out = cx x_in + ce e.   eps_out (optional) receives the guided eps (PLMS history). h1..h3 may be NULL. */
int sd_cfg_step(const float *eps2, long n, float gs, float *eps_out, const float *h1, const float *h2, const float *h3,
                const float c[4], float cx, float ce, const float *x_in, float *x_out, void *stream);

/* Parameter inventory of a configuration (host-only, no device needed): returns the number of parameters; with index >= 0 also
 * writes the index-th (sorted by name) parameter's diffusers name into name[cap] and its shape. */
int sd_unet_inventory(const sd_unet_config *cfg, int index, char *name, size_t cap, long shape[4], int *ndim);

/* Introspection / tests: number of kernels one forward enqueues; copy a named intermediate (conv_in, down.i.j, mid,
 * up.i.j, temb) to HOST as fp32 NCHW (temb: [batch, temb_dim]).  `cap` = floats available in `out`. */
int sd_unet_launch_count(sd_unet *u);           /* per denoise step (context given by sd_unet_set_context) */
int sd_unet_context_launch_count(sd_unet *u);   /* per sd_unet_set_context */
int sd_unet_read_tap(sd_unet *u, const char *name, float *out, size_t cap, int dims[4]);

#ifdef __cplusplus
}
#endif
#endif
