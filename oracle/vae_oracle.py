"""TEST INFRASTRUCTURE — CPU torch restatement of the VAE decode that ends the reference's generation call
(``pipe(...)`` at evalscripts/generate-images-sd.py:37-42 -> ``vae.decode(latents / scaling_factor)`` -> ``(x / 2 + 0.5).clamp(0, 1)``
-> uint8; spelled out in evalscripts/concept_algebra.py:126-135).  The arithmetic lives in ``diffusers==0.33.0``
(models/autoencoders/autoencoder_kl.py, vae.py: Decoder, UNetMidBlock2D, UpDecoderBlock2D, ResnetBlock2D with temb=None, Attention with one
head), which is not vendored and not installable here: this restates its published algorithm for the SD-1.x configuration.

Parity status: **unpinned** (same situation as oracle/unet_oracle.py) — anchored on the exact parameter inventory of the published
checkpoint layout (decoder: 49 490 179 parameters, diffusers state-dict names) and fp64-vs-fp32 self-consistency.  It is groundwork for
SURVEY.md 8(f) rank 1 (the decoder on the B200 kernels); nothing in the product imports it.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from uce_b200.vae_spec import SD14_VAE, decoder_param_shapes


def random_weights(cfg=SD14_VAE, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, shp in decoder_param_shapes(cfg).items():
        if name.endswith(".weight") and len(shp) == 1:
            w = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif name.endswith(".bias"):
            w = 0.02 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            w = torch.randn(shp, generator=g) / math.sqrt(fan_in)
        P[name] = w.to(dtype)
    return P


def _resnet(P, p, x, groups):
    """ResnetBlock2D without time embedding, eps 1e-6, output_scale_factor 1."""
    h = F.silu(F.group_norm(x, groups, P[p + ".norm1.weight"], P[p + ".norm1.bias"], eps=1e-6))
    h = F.conv2d(h, P[p + ".conv1.weight"], P[p + ".conv1.bias"], padding=1)
    h = F.silu(F.group_norm(h, groups, P[p + ".norm2.weight"], P[p + ".norm2.bias"], eps=1e-6))
    h = F.conv2d(h, P[p + ".conv2.weight"], P[p + ".conv2.bias"], padding=1)
    if (p + ".conv_shortcut.weight") in P:
        x = F.conv2d(x, P[p + ".conv_shortcut.weight"], P[p + ".conv_shortcut.bias"])
    return x + h


def _attention(P, p, x, groups):
    """One head over the H*W tokens, head dim = channels; GroupNorm(eps 1e-6) first, residual, rescale_output_factor 1."""
    B, C, H, W = x.shape
    h = F.group_norm(x, groups, P[p + ".group_norm.weight"], P[p + ".group_norm.bias"], eps=1e-6)
    t = h.reshape(B, C, H * W).transpose(1, 2)
    q = F.linear(t, P[p + ".to_q.weight"], P[p + ".to_q.bias"])
    k = F.linear(t, P[p + ".to_k.weight"], P[p + ".to_k.bias"])
    v = F.linear(t, P[p + ".to_v.weight"], P[p + ".to_v.bias"])
    a = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(C), dim=-1) @ v
    o = F.linear(a, P[p + ".to_out.0.weight"], P[p + ".to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(B, C, H, W)


def decode(P, latents, cfg=SD14_VAE, taps=None):
    """``vae.decode(latents / scaling_factor).sample``: [B, 4, h, w] latents -> [B, 3, 8h, 8w] (for the 4-level configuration)."""
    g = cfg["norm_groups"]
    z = latents / cfg["scaling_factor"]
    z = F.conv2d(z, P["post_quant_conv.weight"], P["post_quant_conv.bias"])
    h = F.conv2d(z, P["decoder.conv_in.weight"], P["decoder.conv_in.bias"], padding=1)
    h = _resnet(P, "decoder.mid_block.resnets.0", h, g)
    h = _attention(P, "decoder.mid_block.attentions.0", h, g)
    h = _resnet(P, "decoder.mid_block.resnets.1", h, g)
    if taps is not None:
        taps["mid"] = h
    n_up = len(cfg["block_out_channels"])
    for i in range(n_up):
        for j in range(cfg["layers_per_block"] + 1):
            h = _resnet(P, f"decoder.up_blocks.{i}.resnets.{j}", h, g)
        if i != n_up - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, P[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"], P[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
        if taps is not None:
            taps[f"up.{i}"] = h
    h = F.silu(F.group_norm(h, g, P["decoder.conv_norm_out.weight"], P["decoder.conv_norm_out.bias"], eps=1e-6))
    return F.conv2d(h, P["decoder.conv_out.weight"], P["decoder.conv_out.bias"], padding=1)


def to_uint8(images):
    """(x / 2 + 0.5).clamp(0, 1) -> NHWC -> round to uint8 (concept_algebra.py:130-133)."""
    x = (images.float() / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1)
    return (x * 255).round().to(torch.uint8)
