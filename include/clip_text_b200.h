/* C ABI of the batched CLIP text encoder (libuce_b200.so) — SURVEY.md 8(f) rank 2.
 *
 * Replaces the text-encoder forward the reference runs ONCE PER CONCEPT before its arithmetic starts
 * (`pipe.encode_prompt(...)` at trainscripts/uce_sd_erase.py:29-33 and uce_sd_debias.py:52-56,72-76 -> transformers
 * `CLIPTextModel.forward`, models/clip/modeling_clip.py) by one batched forward over all distinct prompts, with the row select
 * `t_emb[0][:, attention_mask.sum() - 2, :]` (uce_sd_erase.py:34-42) fused behind it.
 *
 * fp32 throughout: the reference loads its pipeline with torch_dtype=float32 for the edits (uce_sd_erase.py:117,197-200) and the edited
 * weights are a function of these rows, so the encoder is held to fp32 parity with the library, not to bf16.  Algorithm: token +
 * position embeddings; L pre-LayerNorm layers of causal multi-head self-attention (q scaled by head_dim^-0.5, only the causal mask —
 * encode_prompt passes no attention mask for SD-1.x) and a quick-GELU MLP; final LayerNorm.
 *
 * Hand-written CUDA (sm_100a), no torch / transformers on this path; return 0 = ok, < 0 = CLIPT_E_*, > 0 = cudaError_t;
 * clipt_last_error() holds the message of the last failure on the calling thread.
 */
#ifndef CLIP_TEXT_B200_H
#define CLIP_TEXT_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLIPT_E_ARG    (-1)
#define CLIPT_E_STATE  (-2)

typedef struct clipt_enc clipt_enc;   /* opaque encoder */

const char *clipt_last_error(void);

/* An encoder on CUDA device `device`: vocabulary, hidden width, heads, layers, MLP width, positions (77 for every SD text encoder) and
 * the largest batch (prompts per forward) it will be asked for. */
int clipt_create(int device, int vocab, int width, int heads, int layers, int ffn, int max_pos, int max_batch, clipt_enc **out);
int clipt_destroy(clipt_enc *e);

/* Upload one parameter by its transformers state-dict name (`text_model.embeddings.token_embedding.weight`,
 * `text_model.encoder.layers.<i>.self_attn.q_proj.weight`, ... `text_model.final_layer_norm.bias`); HOST fp32 data of `n` elements.
 * Unknown names return CLIPT_E_ARG (so a caller can skip `position_ids` and friends); a wrong element count too. */
int clipt_set_weight(clipt_enc *e, const char *name, const float *data, size_t n);

/* All parameters present?  Must be called once before the first forward. */
int clipt_finalize(clipt_enc *e);

/* last_hidden_state [batch, T, width] (DEVICE, fp32) for HOST token ids [batch, T] (int32), T <= max_pos, batch <= max_batch. */
int clipt_encode(clipt_enc *e, const int *input_ids, int batch, int T, float *hidden_out, void *stream);

/* The rows the edit keeps: rows_out[b, :] = last_hidden_state[b, row_index[b], :] (DEVICE [batch, width], fp32); row_index is a HOST
 * array (attention_mask.sum() - 2 per prompt, uce_sd_erase.py:34-42). */
int clipt_concept_rows(clipt_enc *e, const int *input_ids, const int *row_index, int batch, int T, float *rows_out, void *stream);

/* Kernels one forward enqueues (introspection for tests / bench). */
int clipt_launch_count(clipt_enc *e);

#ifdef __cplusplus
}
#endif
#endif
