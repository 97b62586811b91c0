"""TEST INFRASTRUCTURE — CPU restatement of the CLIP text encoder the reference runs once per concept
(``pipe.encode_prompt`` at trainscripts/uce_sd_erase.py:29-33 -> transformers CLIPTextModel; the kept row is the last real token,
``attention_mask.sum() - 2``, :34-42), BATCHED over concepts — the reference encodes them one by one, which dwarfs the solve for
large concept lists.

Parity status: **pinned** — ``transformers`` is installed in the build container, so this restatement is checked against the library's
own CLIPTextModel (random weights from the same state dict; tests/test_clip_text_oracle.py).  Groundwork for the B200 text encoder:
nothing in the product imports it.

Algorithm (transformers/models/clip/modeling_clip.py): token + position embeddings; L pre-LayerNorm layers of causal multi-head self
attention (q scaled by head_dim^-0.5) and a quick-GELU MLP (x * sigmoid(1.702 x)); final LayerNorm.  Padding is NOT masked by the
pipeline's encode_prompt for SD-1.x (no attention_mask is passed): only the causal mask applies.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def encode(P, input_ids, heads, eps=1e-5):
    """P: CLIPTextModel state dict (names ``text_model.*``).  input_ids [B, T] -> last_hidden_state [B, T, D]."""
    pre = "text_model."
    B, T = input_ids.shape
    x = P[pre + "embeddings.token_embedding.weight"][input_ids] + P[pre + "embeddings.position_embedding.weight"][:T][None]
    D = x.shape[-1]
    dh = D // heads
    causal = torch.full((T, T), float("-inf"), dtype=x.dtype).triu(1)
    n_layers = 1 + max(int(k.split(".")[3]) for k in P if k.startswith(pre + "encoder.layers."))
    for i in range(n_layers):
        l = f"{pre}encoder.layers.{i}."
        h = F.layer_norm(x, (D,), P[l + "layer_norm1.weight"], P[l + "layer_norm1.bias"], eps)
        q = F.linear(h, P[l + "self_attn.q_proj.weight"], P[l + "self_attn.q_proj.bias"]) * dh ** -0.5
        k = F.linear(h, P[l + "self_attn.k_proj.weight"], P[l + "self_attn.k_proj.bias"])
        v = F.linear(h, P[l + "self_attn.v_proj.weight"], P[l + "self_attn.v_proj.bias"])
        q, k, v = (t.reshape(B, T, heads, dh).transpose(1, 2) for t in (q, k, v))
        a = torch.softmax(q @ k.transpose(-1, -2) + causal, dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, T, D)
        x = x + F.linear(a, P[l + "self_attn.out_proj.weight"], P[l + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (D,), P[l + "layer_norm2.weight"], P[l + "layer_norm2.bias"], eps)
        h = F.linear(h, P[l + "mlp.fc1.weight"], P[l + "mlp.fc1.bias"])
        h = h * torch.sigmoid(1.702 * h)
        x = x + F.linear(h, P[l + "mlp.fc2.weight"], P[l + "mlp.fc2.bias"])
    return F.layer_norm(x, (D,), P[pre + "final_layer_norm.weight"], P[pre + "final_layer_norm.bias"], eps)


def concept_rows(P, input_ids, attention_mask, heads):
    """The one row per concept the edit keeps: index ``attention_mask.sum() - 2`` (uce_sd_erase.py:34-42), for a whole batch."""
    h = encode(P, input_ids, heads)
    idx = attention_mask.sum(dim=1) - 2
    return h[torch.arange(h.shape[0]), idx]
