"""TEST INFRASTRUCTURE — CPU torch restatement of the edited-model generation step.

The reference's generation path (evalscripts/generate-images-sd.py:37-42, and
trainscripts/uce_sd_debias.py:22-26) is a call into ``diffusers==0.33.0`` (requirements.txt:1), which is
NOT vendored under /root/reference and not installable here (no network).  Per SURVEY.md §8c this file
restates the published algorithm of that dependency for the one configuration the reference uses
(UNet2DConditionModel as configured by CompVis/stable-diffusion-v1-4, PNDM/PLMS default scheduler, DDIM
optional, classifier-free guidance) — operationally specified in SURVEY.md Appendix A/B.

Parity status: **unpinned** — the reference ships no golden vectors for this path and diffusers cannot be
executed here; the restatement is anchored on (i) the exact parameter inventory of the published
checkpoint layout (859 520 964 parameters, diffusers state-dict names — real checkpoints load with
``strict`` key equality), (ii) fp64-vs-fp32 self-consistency, (iii) the in-repo hand-rolled loop
evalscripts/concept_algebra.py:56-135 for the call order (tokenize → encode → unet → guidance →
scheduler.step).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from oracle.unet_params import unet_named_parameters

# The one configuration the reference uses (CompVis/stable-diffusion-v1-4/unet/config.json), restated here: the oracle does not import
# the product's architecture table (tests/test_unet_params.py holds the two to each other).
SD14 = dict(in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
            down_has_attn=(True, True, True, False), up_has_attn=(False, True, True, True),
            cross_attention_dim=768, heads=8, norm_groups=32, temb_dim=1280)


def param_shapes(cfg=SD14):
    """named_parameters() -> shape for a configuration dict of the engine's form, from the oracle's own walk of the diffusers tree."""
    return unet_named_parameters(block_out_channels=tuple(cfg["block_out_channels"]), layers_per_block=cfg["layers_per_block"],
                                 cross_attention_dim=cfg["cross_attention_dim"], in_channels=cfg["in_channels"], out_channels=cfg["out_channels"],
                                 down_block_types=tuple("CrossAttnDownBlock2D" if a else "DownBlock2D" for a in cfg["down_has_attn"]),
                                 up_block_types=tuple("CrossAttnUpBlock2D" if a else "UpBlock2D" for a in cfg["up_has_attn"]))


# ------------------------------------------------------------------------------------------ weights
def random_weights(cfg=SD14, seed=0, dtype=torch.float32):
    """Seeded synthetic weights with fan-in scaling (activations stay O(1)); norm gains 1, biases small."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, shp in param_shapes(cfg).items():
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


# ------------------------------------------------------------------------------------------ blocks
def timestep_embedding(t, dim, dtype):
    """flip_sin_to_cos=True, freq_shift=0: [cos | sin], freqs exp(-ln(1e4) i / half) — computed in fp32."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    args = torch.as_tensor(t, dtype=torch.float32).reshape(-1, 1) * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1).to(dtype)


def _resnet(P, p, x, temb, groups):
    h = F.silu(F.group_norm(x, groups, P[p + ".norm1.weight"], P[p + ".norm1.bias"], eps=1e-5))
    h = F.conv2d(h, P[p + ".conv1.weight"], P[p + ".conv1.bias"], padding=1)
    h = h + F.linear(F.silu(temb), P[p + ".time_emb_proj.weight"], P[p + ".time_emb_proj.bias"])[:, :, None, None]
    h = F.silu(F.group_norm(h, groups, P[p + ".norm2.weight"], P[p + ".norm2.bias"], eps=1e-5))
    h = F.conv2d(h, P[p + ".conv2.weight"], P[p + ".conv2.bias"], padding=1)
    if (p + ".conv_shortcut.weight") in P:
        x = F.conv2d(x, P[p + ".conv_shortcut.weight"], P[p + ".conv_shortcut.bias"])
    return x + h


def _attention(P, p, x, ctx, heads):
    q = F.linear(x, P[p + ".to_q.weight"])
    k = F.linear(ctx, P[p + ".to_k.weight"])
    v = F.linear(ctx, P[p + ".to_v.weight"])
    B, L, C = q.shape
    d = C // heads
    q = q.view(B, L, heads, d).transpose(1, 2)
    k = k.view(B, -1, heads, d).transpose(1, 2)
    v = v.view(B, -1, heads, d).transpose(1, 2)
    a = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(d), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, L, C)
    return F.linear(o, P[p + ".to_out.0.weight"], P[p + ".to_out.0.bias"])


def _transformer(P, p, x, ctx, groups, heads):
    B, C, H, W = x.shape
    res = x
    h = F.group_norm(x, groups, P[p + ".norm.weight"], P[p + ".norm.bias"], eps=1e-6)
    h = F.conv2d(h, P[p + ".proj_in.weight"], P[p + ".proj_in.bias"])
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    b = p + ".transformer_blocks.0"
    n = F.layer_norm(h, (C,), P[b + ".norm1.weight"], P[b + ".norm1.bias"], eps=1e-5)
    h = h + _attention(P, b + ".attn1", n, n, heads)
    n = F.layer_norm(h, (C,), P[b + ".norm2.weight"], P[b + ".norm2.bias"], eps=1e-5)
    h = h + _attention(P, b + ".attn2", n, ctx, heads)
    n = F.layer_norm(h, (C,), P[b + ".norm3.weight"], P[b + ".norm3.bias"], eps=1e-5)
    ff = F.linear(n, P[b + ".ff.net.0.proj.weight"], P[b + ".ff.net.0.proj.bias"])
    hid, gate = ff.chunk(2, dim=-1)
    h = h + F.linear(hid * F.gelu(gate), P[b + ".ff.net.2.weight"], P[b + ".ff.net.2.bias"])
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    h = F.conv2d(h, P[p + ".proj_out.weight"], P[p + ".proj_out.bias"])
    return h + res


def unet_forward(P, x, t, ctx, cfg=SD14, taps=None):
    """eps = UNet(x [NB,4,H,W], t, ctx [NB,77,ctx_dim]).  ``taps`` (dict) collects named intermediates."""
    ch = cfg["block_out_channels"]; lpb = cfg["layers_per_block"]; G = cfg["norm_groups"]; heads = cfg["heads"]
    dtype = x.dtype
    temb = timestep_embedding(t, ch[0], dtype).expand(x.shape[0], -1)
    temb = F.linear(temb, P["time_embedding.linear_1.weight"], P["time_embedding.linear_1.bias"])
    temb = F.linear(F.silu(temb), P["time_embedding.linear_2.weight"], P["time_embedding.linear_2.bias"])
    h = F.conv2d(x, P["conv_in.weight"], P["conv_in.bias"], padding=1)
    if taps is not None:
        taps["temb"] = temb; taps["conv_in"] = h
    skips = [h]
    for i in range(len(ch)):
        for j in range(lpb):
            h = _resnet(P, f"down_blocks.{i}.resnets.{j}", h, temb, G)
            if cfg["down_has_attn"][i]:
                h = _transformer(P, f"down_blocks.{i}.attentions.{j}", h, ctx, G, heads)
            skips.append(h)
            if taps is not None:
                taps[f"down.{i}.{j}"] = h
        if i < len(ch) - 1:
            h = F.conv2d(h, P[f"down_blocks.{i}.downsamplers.0.conv.weight"], P[f"down_blocks.{i}.downsamplers.0.conv.bias"], stride=2, padding=1)
            skips.append(h)
    h = _resnet(P, "mid_block.resnets.0", h, temb, G)
    h = _transformer(P, "mid_block.attentions.0", h, ctx, G, heads)
    h = _resnet(P, "mid_block.resnets.1", h, temb, G)
    if taps is not None:
        taps["mid"] = h
    for i in range(len(ch)):
        for j in range(lpb + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = _resnet(P, f"up_blocks.{i}.resnets.{j}", h, temb, G)
            if cfg["up_has_attn"][i]:
                h = _transformer(P, f"up_blocks.{i}.attentions.{j}", h, ctx, G, heads)
            if taps is not None:
                taps[f"up.{i}.{j}"] = h
        if i < len(ch) - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, P[f"up_blocks.{i}.upsamplers.0.conv.weight"], P[f"up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
    h = F.silu(F.group_norm(h, G, P["conv_norm_out.weight"], P["conv_norm_out.bias"], eps=1e-5))
    return F.conv2d(h, P["conv_out.weight"], P["conv_out.bias"], padding=1)


# ------------------------------------------------------------------------------------------ schedulers
def alphas_cumprod():
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2     # scaled_linear
    return torch.cumprod(1.0 - betas, dim=0)


class PNDMOracle:
    """PNDM with skip_prk_steps=True (PLMS), steps_offset=1, set_alpha_to_one=False — the SD-1.4 default the
    reference's pipeline uses: S+1 U-Net calls for S steps."""

    def __init__(self, steps):
        self.ac = alphas_cumprod()
        self.r = 1000 // steps
        ts = (np.arange(steps) * self.r).round().astype(np.int64) + 1
        self.timesteps = np.concatenate([ts[:-1], ts[-2:-1], ts[-1:]])[::-1].copy()
        self.ets, self.counter, self.cur_sample = [], 0, None

    def step(self, eps, t, x):
        t = int(t)
        prev = t - self.r
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(eps)
        else:
            prev = t
            t = t + self.r
        if len(self.ets) == 1 and self.counter == 0:
            e = eps
            self.cur_sample = x
        elif len(self.ets) == 1 and self.counter == 1:
            e = (eps + self.ets[-1]) / 2
            x = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            e = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            e = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            e = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2] + 37 * self.ets[-3] - 9 * self.ets[-4])
        self.counter += 1
        return self._prev(x, t, prev, e)

    def _prev(self, x, t, prev, e):
        a = self.ac[t].item()
        ap = self.ac[prev].item() if prev >= 0 else self.ac[0].item()
        b, bp = 1 - a, 1 - ap
        coeff = (ap / a) ** 0.5
        denom = a * bp ** 0.5 + (a * b * ap) ** 0.5
        return coeff * x - (ap - a) * e / denom

    def coefficients(self, t, prev):
        """(coeff_x, coeff_e) of x_prev = coeff_x x + coeff_e e — what the fused CUDA step consumes."""
        a = self.ac[t].item()
        ap = self.ac[prev].item() if prev >= 0 else self.ac[0].item()
        b, bp = 1 - a, 1 - ap
        return (ap / a) ** 0.5, -(ap - a) / (a * bp ** 0.5 + (a * b * ap) ** 0.5)


class DDIMOracle:
    """DDIM, eta=0, steps_offset=1, set_alpha_to_one=False, no clipping."""

    def __init__(self, steps):
        self.ac = alphas_cumprod()
        self.r = 1000 // steps
        self.timesteps = ((np.arange(steps) * self.r).round()[::-1].copy().astype(np.int64)) + 1

    def step(self, eps, t, x):
        t = int(t)
        prev = t - self.r
        a = self.ac[t].item()
        ap = self.ac[prev].item() if prev >= 0 else self.ac[0].item()
        x0 = (x - (1 - a) ** 0.5 * eps) / a ** 0.5
        return ap ** 0.5 * x0 + (1 - ap) ** 0.5 * eps


def cfg_combine(eps2, guidance_scale):
    """[uncond, text] batch order; eps = eps_u + gs (eps_t − eps_u)."""
    eu, et = eps2.chunk(2)
    return eu + guidance_scale * (et - eu)


def denoise_loop(P, latents, ctx_uncond_text, steps=50, guidance_scale=7.5, scheduler="pndm", cfg=SD14):
    """The hot loop of StableDiffusionPipeline.__call__ on given initial latents and [2B,77,D] context."""
    sch = PNDMOracle(steps) if scheduler == "pndm" else DDIMOracle(steps)
    x = latents
    for t in sch.timesteps:
        eps2 = unet_forward(P, torch.cat([x, x]), int(t), ctx_uncond_text, cfg)
        x = sch.step(cfg_combine(eps2, guidance_scale), int(t), x)
    return x
