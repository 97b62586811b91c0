"""Architecture table of the Stable Diffusion v1.x VAE DECODER (AutoencoderKL as configured by
CompVis/stable-diffusion-v1-4/vae/config.json: block_out_channels (128, 256, 512, 512), layers_per_block 2, latent_channels 4,
norm_num_groups 32, scaling_factor 0.18215) — what ``vae.decode`` needs (post_quant_conv + decoder), parameter names as in the
diffusers state dict so real checkpoints load unchanged.  Pure data: no torch ops.  The decoder is §8(f) rank 1 of SURVEY.md (the
reference reaches it inside ``pipe(...)``: evalscripts/generate-images-sd.py:37-46; explicit form in evalscripts/concept_algebra.py:126-135)."""
from __future__ import annotations

from collections import OrderedDict

SD14_VAE = dict(latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2, norm_groups=32,
                scaling_factor=0.18215)
SD14_VAE_DECODER_PARAMS = 49_490_179          # decoder only; + 20 for post_quant_conv


def tiny_vae_config(ch=(16, 32), groups=4):
    return dict(latent_channels=4, out_channels=3, block_out_channels=tuple(ch), layers_per_block=1, norm_groups=groups, scaling_factor=0.18215)


def _resnet(s, p, cin, cout):
    s[p + ".norm1.weight"] = (cin,); s[p + ".norm1.bias"] = (cin,)
    s[p + ".conv1.weight"] = (cout, cin, 3, 3); s[p + ".conv1.bias"] = (cout,)
    s[p + ".norm2.weight"] = (cout,); s[p + ".norm2.bias"] = (cout,)
    s[p + ".conv2.weight"] = (cout, cout, 3, 3); s[p + ".conv2.bias"] = (cout,)
    if cin != cout:
        s[p + ".conv_shortcut.weight"] = (cout, cin, 1, 1); s[p + ".conv_shortcut.bias"] = (cout,)


def decoder_param_shapes(cfg=SD14_VAE) -> "OrderedDict[str, tuple]":
    ch = cfg["block_out_channels"]; top = ch[-1]; lat = cfg["latent_channels"]
    s = OrderedDict()
    s["post_quant_conv.weight"] = (lat, lat, 1, 1); s["post_quant_conv.bias"] = (lat,)
    s["decoder.conv_in.weight"] = (top, lat, 3, 3); s["decoder.conv_in.bias"] = (top,)
    _resnet(s, "decoder.mid_block.resnets.0", top, top)
    a = "decoder.mid_block.attentions.0"
    s[a + ".group_norm.weight"] = (top,); s[a + ".group_norm.bias"] = (top,)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        s[f"{a}.{n}.weight"] = (top, top); s[f"{a}.{n}.bias"] = (top,)
    _resnet(s, "decoder.mid_block.resnets.1", top, top)
    cur = top
    rev = list(reversed(ch))
    for i, cout in enumerate(rev):
        for j in range(cfg["layers_per_block"] + 1):
            _resnet(s, f"decoder.up_blocks.{i}.resnets.{j}", cur, cout)
            cur = cout
        if i != len(rev) - 1:
            s[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"] = (cout, cout, 3, 3); s[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"] = (cout,)
    s["decoder.conv_norm_out.weight"] = (cur,); s["decoder.conv_norm_out.bias"] = (cur,)
    s["decoder.conv_out.weight"] = (cfg["out_channels"], cur, 3, 3); s["decoder.conv_out.bias"] = (cfg["out_channels"],)
    return s


def decoder_param_count(cfg=SD14_VAE, with_post_quant=False) -> int:
    n = 0
    for name, shp in decoder_param_shapes(cfg).items():
        if not with_post_quant and name.startswith("post_quant_conv"):
            continue
        k = 1
        for d in shp:
            k *= d
        n += k
    return n
