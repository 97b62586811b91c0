"""Architecture table of the Stable Diffusion v1.x U-Net (UNet2DConditionModel as configured by
CompVis/stable-diffusion-v1-4/unet/config.json) — parameter names follow the
diffusers state-dict so real checkpoints load unchanged.  Pure data: no torch ops."""
from __future__ import annotations

from collections import OrderedDict

SD14 = dict(
    in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
    down_has_attn=(True, True, True, False), up_has_attn=(False, True, True, True),
    cross_attention_dim=768, heads=8, norm_groups=32, temb_dim=1280,
)


def tiny_config(ch=(32, 64), ctx_dim=48, heads=4, groups=8):
    """Reduced copy of the same topology for fast CPU/GPU parity tests."""
    n = len(ch)
    return dict(in_channels=4, out_channels=4, block_out_channels=tuple(ch), layers_per_block=2,
                down_has_attn=tuple([True] * (n - 1) + [False]), up_has_attn=tuple([False] + [True] * (n - 1)),
                cross_attention_dim=ctx_dim, heads=heads, norm_groups=groups, temb_dim=4 * ch[0])


def _resnet(shapes, p, cin, cout, temb):
    shapes[p + ".norm1.weight"] = (cin,); shapes[p + ".norm1.bias"] = (cin,)
    shapes[p + ".conv1.weight"] = (cout, cin, 3, 3); shapes[p + ".conv1.bias"] = (cout,)
    shapes[p + ".time_emb_proj.weight"] = (cout, temb); shapes[p + ".time_emb_proj.bias"] = (cout,)
    shapes[p + ".norm2.weight"] = (cout,); shapes[p + ".norm2.bias"] = (cout,)
    shapes[p + ".conv2.weight"] = (cout, cout, 3, 3); shapes[p + ".conv2.bias"] = (cout,)
    if cin != cout:
        shapes[p + ".conv_shortcut.weight"] = (cout, cin, 1, 1); shapes[p + ".conv_shortcut.bias"] = (cout,)


def _transformer(shapes, p, c, ctx):
    shapes[p + ".norm.weight"] = (c,); shapes[p + ".norm.bias"] = (c,)
    shapes[p + ".proj_in.weight"] = (c, c, 1, 1); shapes[p + ".proj_in.bias"] = (c,)
    b = p + ".transformer_blocks.0"
    for i, kdim in ((1, c), (2, ctx)):
        shapes[f"{b}.norm{i}.weight"] = (c,); shapes[f"{b}.norm{i}.bias"] = (c,)
        shapes[f"{b}.attn{i}.to_q.weight"] = (c, c)
        shapes[f"{b}.attn{i}.to_k.weight"] = (c, kdim)
        shapes[f"{b}.attn{i}.to_v.weight"] = (c, kdim)
        shapes[f"{b}.attn{i}.to_out.0.weight"] = (c, c); shapes[f"{b}.attn{i}.to_out.0.bias"] = (c,)
    shapes[b + ".norm3.weight"] = (c,); shapes[b + ".norm3.bias"] = (c,)
    shapes[b + ".ff.net.0.proj.weight"] = (8 * c, c); shapes[b + ".ff.net.0.proj.bias"] = (8 * c,)
    shapes[b + ".ff.net.2.weight"] = (c, 4 * c); shapes[b + ".ff.net.2.bias"] = (c,)
    shapes[p + ".proj_out.weight"] = (c, c, 1, 1); shapes[p + ".proj_out.bias"] = (c,)


def param_shapes(cfg=SD14) -> "OrderedDict[str, tuple]":
    ch = cfg["block_out_channels"]; temb = cfg["temb_dim"]; ctx = cfg["cross_attention_dim"]; lpb = cfg["layers_per_block"]
    s = OrderedDict()
    s["conv_in.weight"] = (ch[0], cfg["in_channels"], 3, 3); s["conv_in.bias"] = (ch[0],)
    s["time_embedding.linear_1.weight"] = (temb, ch[0]); s["time_embedding.linear_1.bias"] = (temb,)
    s["time_embedding.linear_2.weight"] = (temb, temb); s["time_embedding.linear_2.bias"] = (temb,)
    skip_ch = [ch[0]]
    cur = ch[0]
    for i, cout in enumerate(ch):
        for j in range(lpb):
            _resnet(s, f"down_blocks.{i}.resnets.{j}", cur, cout, temb)
            cur = cout
            if cfg["down_has_attn"][i]:
                _transformer(s, f"down_blocks.{i}.attentions.{j}", cout, ctx)
            skip_ch.append(cur)
        if i < len(ch) - 1:
            s[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (cout, cout, 3, 3); s[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (cout,)
            skip_ch.append(cur)
    _resnet(s, "mid_block.resnets.0", cur, cur, temb)
    _transformer(s, "mid_block.attentions.0", cur, ctx)
    _resnet(s, "mid_block.resnets.1", cur, cur, temb)
    rev = list(reversed(ch))
    for i, cout in enumerate(rev):
        for j in range(lpb + 1):
            cin = cur + skip_ch.pop()
            _resnet(s, f"up_blocks.{i}.resnets.{j}", cin, cout, temb)
            cur = cout
            if cfg["up_has_attn"][i]:
                _transformer(s, f"up_blocks.{i}.attentions.{j}", cout, ctx)
        if i < len(ch) - 1:
            s[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (cout, cout, 3, 3); s[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (cout,)
    s["conv_norm_out.weight"] = (ch[0],); s["conv_norm_out.bias"] = (ch[0],)
    s["conv_out.weight"] = (cfg["out_channels"], ch[0], 3, 3); s["conv_out.bias"] = (cfg["out_channels"],)
    return s


def param_count(cfg=SD14) -> int:
    n = 0
    for shp in param_shapes(cfg).values():
        k = 1
        for d in shp:
            k *= d
        n += k
    return n
