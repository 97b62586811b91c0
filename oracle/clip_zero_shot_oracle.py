"""TEST INFRASTRUCTURE — CPU restatement of the zero-shot image classifier the debias edit scores its generated images with
(``transformers.pipeline(task="zero-shot-image-classification", model="openai/clip-vit-base-patch32")`` built at
trainscripts/uce_sd_debias.py:245-250 and called at :27: top-1 label over ``debias_concepts`` per image).

Parity status: **pinned** — ``transformers`` is installed in the build container, so the model arithmetic below is checked against the
library's own ``CLIPModel`` on random weights of a reduced configuration (tests/test_clip_zero_shot_oracle.py), and the label logic
against what ``get_ratios`` consumes.  Groundwork for SURVEY.md 8(f) rank 3 (the classifier loop on the B200 kernels): nothing in the
product imports it.

Algorithm (transformers/models/clip/modeling_clip.py, pipelines/zero_shot_image_classification.py):
  vision tower  patch embedding (conv, stride = patch, no bias) + class token + position embeddings -> pre-LayerNorm -> L pre-LN layers
                of full (unmasked) multi-head self attention and a quick-GELU MLP -> post-LayerNorm of the class token -> projection;
  text tower    oracle/clip_text_oracle.encode (causal), pooled at the end-of-text token, -> projection;
  score         logit_scale.exp() * cos(image, text) -> softmax over the candidate labels ("This is a photo of {label}.") -> top-1.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from oracle import clip_text_oracle as T

HYPOTHESIS_TEMPLATE = "This is a photo of {}."          # the pipeline's default


def _layers(P, pre, x, heads, eps, n_layers):
    B, L, D = x.shape
    dh = D // heads
    for i in range(n_layers):
        l = f"{pre}encoder.layers.{i}."
        h = F.layer_norm(x, (D,), P[l + "layer_norm1.weight"], P[l + "layer_norm1.bias"], eps)
        q = F.linear(h, P[l + "self_attn.q_proj.weight"], P[l + "self_attn.q_proj.bias"]) * dh ** -0.5
        k = F.linear(h, P[l + "self_attn.k_proj.weight"], P[l + "self_attn.k_proj.bias"])
        v = F.linear(h, P[l + "self_attn.v_proj.weight"], P[l + "self_attn.v_proj.bias"])
        q, k, v = (t.reshape(B, L, heads, dh).transpose(1, 2) for t in (q, k, v))
        a = torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v
        x = x + F.linear(a.transpose(1, 2).reshape(B, L, D), P[l + "self_attn.out_proj.weight"], P[l + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (D,), P[l + "layer_norm2.weight"], P[l + "layer_norm2.bias"], eps)
        h = F.linear(h, P[l + "mlp.fc1.weight"], P[l + "mlp.fc1.bias"])
        x = x + F.linear(h * torch.sigmoid(1.702 * h), P[l + "mlp.fc2.weight"], P[l + "mlp.fc2.bias"])
    return x


def image_features(P, pixel_values, heads, eps=1e-5):
    """pixel_values [B, 3, S, S] (already resized / normalised) -> projected image embeddings [B, E]."""
    pre = "vision_model."
    w = P[pre + "embeddings.patch_embedding.weight"]
    patch = w.shape[-1]
    x = F.conv2d(pixel_values, w, stride=patch).flatten(2).transpose(1, 2)               # [B, n_patches, D]
    cls = P[pre + "embeddings.class_embedding"].expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1) + P[pre + "embeddings.position_embedding.weight"][None]
    D = x.shape[-1]
    x = F.layer_norm(x, (D,), P[pre + "pre_layrnorm.weight"], P[pre + "pre_layrnorm.bias"], eps)      # (sic: the checkpoint's spelling)
    n_layers = 1 + max(int(k.split(".")[3]) for k in P if k.startswith(pre + "encoder.layers."))
    x = _layers(P, pre, x, heads, eps, n_layers)
    pooled = F.layer_norm(x[:, 0], (D,), P[pre + "post_layernorm.weight"], P[pre + "post_layernorm.bias"], eps)
    return F.linear(pooled, P["visual_projection.weight"])


def text_features(P, input_ids, heads, eos_token_id):
    """input_ids [N, T] -> projected text embeddings [N, E]; pooled at the first end-of-text token (for the original checkpoints, whose
    eos id is the largest id, that is ``argmax(input_ids)`` — modeling_clip.py keeps both spellings)."""
    h = T.encode({k: v for k, v in P.items() if k.startswith("text_model.")}, input_ids, heads)
    eos = (input_ids == eos_token_id).int().argmax(dim=-1)
    return F.linear(h[torch.arange(h.shape[0]), eos], P["text_projection.weight"])


def logits_per_image(P, pixel_values, input_ids, vision_heads, text_heads, eos_token_id):
    """[B, N] = exp(logit_scale) * cos(image_b, text_n)."""
    im = image_features(P, pixel_values, vision_heads)
    tx = text_features(P, input_ids, text_heads, eos_token_id)
    im = im / im.norm(dim=-1, keepdim=True)
    tx = tx / tx.norm(dim=-1, keepdim=True)
    return P["logit_scale"].exp() * im @ tx.t()


def classify(logits, candidate_labels):
    """What the pipeline returns per image: labels sorted by softmax score, best first (uce_sd_debias.py:27 keeps result[0]['label'])."""
    probs = torch.softmax(logits, dim=-1)
    out = []
    for row in probs:
        order = sorted(range(len(candidate_labels)), key=lambda j: -float(row[j]))
        out.append([{"score": float(row[j]), "label": candidate_labels[j]} for j in order])
    return out


def preprocess(images_uint8, size=224, mean=(0.48145466, 0.4578275, 0.40821073), std=(0.26862954, 0.26130258, 0.27577711)):
    """CLIPImageProcessor for SQUARE uint8 RGB images [B, H, H, 3] (the generator's 512 x 512 output): bicubic resize to ``size`` (PIL
    semantics are antialiased; this uses torch's antialiased bicubic, equal to ~1 uint8 level), rescale by 1/255, normalise."""
    x = images_uint8.permute(0, 3, 1, 2).float()
    if x.shape[-1] != size:
        x = F.interpolate(x, size=(size, size), mode="bicubic", antialias=True, align_corners=False).round().clamp(0, 255)
    x = x / 255.0
    m = torch.tensor(mean).view(1, 3, 1, 1); s = torch.tensor(std).view(1, 3, 1, 1)
    return (x - m) / s
