/* C ABI of the CLIP image tower + zero-shot scoring (libuce_b200.so) — SURVEY.md 8(f) rank 3.
 *
 * Replaces the model arithmetic of the classifier the debias edit scores its generated images with:
 * `transformers.pipeline("zero-shot-image-classification", "openai/clip-vit-base-patch32")` built at trainscripts/uce_sd_debias.py:245-250 and
 * called at :27 (`clip(images, candidate_labels=debias_concepts)`, top-1 label per image).  Pieces:
 *   preprocessing  CLIPImageProcessor for square uint8 RGB images: antialiased bicubic resize to the model's input size, rounding to
 *                  uint8 levels, / 255, per-channel normalisation
 *   vision tower   patch embedding (stride = patch, no bias) + class token + position embeddings -> pre-LayerNorm -> L pre-LN layers of
 *                  full multi-head self-attention and a quick-GELU MLP -> post-LayerNorm of the class token -> visual projection
 *   scoring        logit_scale.exp() * cos(image, text) with the text rows (pooled at end-of-text by the text tower,
 *                  include/clip_text_b200.h) put through the text projection
 * fp32 CUDA kernels (the reference runs this model in bf16; fp32 is the more accurate side of its own oracle).  Return codes as in
 * clip_text_b200.h; clipv_last_error() holds the message of the last failure on the calling thread.
 */
#ifndef CLIP_VISION_B200_H
#define CLIP_VISION_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLIPV_E_ARG    (-1)
#define CLIPV_E_STATE  (-2)

typedef struct clipv_enc clipv_enc;

const char *clipv_last_error(void);

/* image_size / patch: 224 / 32 for ViT-B/32; width, heads, layers, ffn of the vision transformer; proj_dim = width of the joint embedding;
 * text_width = hidden width of the text tower whose rows clipv_logits projects; max_batch images per forward. */
int clipv_create(int device, int image_size, int patch, int width, int heads, int layers, int ffn, int proj_dim, int text_width, int max_batch,
                 clipv_enc **out);
int clipv_destroy(clipv_enc *e);

/* Parameters by their transformers CLIPModel state-dict names: `vision_model.*`, `visual_projection.weight`, `text_projection.weight`,
 * `logit_scale` (1 element); HOST fp32. */
int clipv_set_weight(clipv_enc *e, const char *name, const float *data, size_t n);
int clipv_finalize(clipv_enc *e);

/* pixel_values [batch, 3, S, S] (DEVICE fp32) <- images [batch, H, H, 3] (DEVICE uint8, square).  mean / std: HOST, 3 floats each. */
int clipv_preprocess_u8(clipv_enc *e, const unsigned char *images, int batch, int H, const float *mean, const float *std,
                        float *pixel_values, void *stream);

/* image features [batch, proj_dim] (DEVICE fp32, NOT normalised) for pixel_values [batch, 3, S, S] (DEVICE fp32). */
int clipv_image_features(clipv_enc *e, const float *pixel_values, int batch, float *features, void *stream);

/* logits [n_images, n_texts] (DEVICE fp32) = exp(logit_scale) * cos(image_features[i], text_projection . text_rows[j]);
 * text_rows [n_texts, text_width] are the text tower's rows at the end-of-text token (clipt_concept_rows). */
int clipv_logits(clipv_enc *e, const float *image_features, int n_images, const float *text_rows, int n_texts, float *logits, void *stream);

int clipv_launch_count(clipv_enc *e);

#ifdef __cplusplus
}
#endif
#endif
