/*
 * sd_vae_b200.h — C ABI of the B200-native Stable Diffusion v1.x VAE DECODER (libuce_b200.so).
 *
 * Replaces what the reference reaches through diffusers at the end of `pipe(...)`:
 *   evalscripts/generate-images-sd.py:37-46 and trainscripts/uce_sd_debias.py:22-26  ->
 *   AutoencoderKL.decode(latents / scaling_factor) -> (x / 2 + 0.5).clamp(0, 1) -> uint8 NHWC
 *   (spelled out in evalscripts/concept_algebra.py:126-135).
 * The decoder is scheduled on the same kernels as the U-Net step (sd_unet_b200.h): tcgen05 implicit-GEMM convolutions,
 * GroupNorm(+SiLU), the GEMM/softmax attention path for the one 512-wide head of the mid block.  Weights use the diffusers
 * state-dict names of `pipe.vae` (`post_quant_conv.*`, `decoder.*`).
 *
 * STATUS (round 1): opt-in.  Built and host-tested; its GPU parity test (tests/test_vae_gpu.py) has not run on hardware yet and is
 * gated by UCE_TEST_VAE=1, and generate_images() uses the engine only with UCE_VAE_ENGINE=1.
 *
 * Conventions as in sd_unet_b200.h: plain C; device pointers unless stated; status 0 = ok, <0 = SD_E_*, >0 = cudaError_t;
 * sd_last_error() describes the last failure; all work is enqueued on the caller's stream.
 */
#ifndef SD_VAE_B200_H
#define SD_VAE_B200_H

#include <stddef.h>
#include "sd_unet_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sd_vae sd_vae;

typedef struct sd_vae_config {
    int latent_channels, out_channels;        /* 4, 3 */
    int n_levels;                             /* length of block_out_channels (<= 4) */
    int block_out_channels[4];                /* encoder order, e.g. 128 256 512 512: the decoder walks it backwards */
    int layers_per_block;                     /* the decoder runs layers_per_block + 1 resnets per level */
    int norm_groups;
    float scaling_factor;                     /* 0.18215: decode() divides the latents by it */
} sd_vae_config;

/* Decoder for `batch` latents of size h x w per call; the image is (h * 2^(n_levels-1)) x (w * 2^(n_levels-1)). */
int sd_vae_create(int device, const sd_vae_config *cfg, int batch, int h, int w, sd_vae **out);
int sd_vae_destroy(sd_vae *v);

/* Upload one parameter by its diffusers state-dict name (fp32, HOST pointer, row-major `shape`); before finalize only. */
int sd_vae_set_weight(sd_vae *v, const char *name, const float *data, const long *shape, int ndim);

/* Allocate activations, build the kernel schedule and the TMA tensor maps.  Fails if a parameter is missing. */
int sd_vae_finalize(sd_vae *v);

/* latents[batch,4,h,w] fp32 NCHW (the denoiser's output, NOT yet divided by scaling_factor)  ->
 *   image[batch,3,H,W] fp32 NCHW = vae.decode(latents / scaling_factor).sample            (optional, may be NULL)
 *   rgb8[batch,H,W,3]  uint8 = round(clamp(image / 2 + 0.5, 0, 1) * 255), ties to even    (optional, may be NULL) */
int sd_vae_decode(sd_vae *v, const float *latents, float *image, unsigned char *rgb8, void *stream);

/* Parameter inventory of a configuration (host-only, no device needed): returns the number of parameters; with index >= 0 also
 * writes the index-th (sorted by name) parameter's diffusers name into name[cap] and its shape. */
int sd_vae_inventory(const sd_vae_config *cfg, int index, char *name, size_t cap, long shape[4], int *ndim);

/* Introspection / tests: kernels one decode enqueues; copy a named intermediate (mid, up.i) to HOST as fp32 NCHW. */
int sd_vae_launch_count(sd_vae *v);
int sd_vae_read_tap(sd_vae *v, const char *name, float *out, size_t cap, int dims[4]);

#ifdef __cplusplus
}
#endif
#endif
