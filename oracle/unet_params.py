"""TEST INFRASTRUCTURE — parameter inventory of diffusers' UNet2DConditionModel (0.33.0) for the SD-1.x configuration, written
INDEPENDENTLY of the product's table (unified-concept-editing_b200/unet_spec.py): it walks the module tree the way diffusers builds
it (``UNet2DConditionModel.__init__`` -> get_down_block / UNetMidBlock2DCrossAttn / get_up_block -> ResnetBlock2D,
Transformer2DModel -> BasicTransformerBlock -> Attention / FeedForward(GEGLU), Downsample2D, Upsample2D) and emits
``named_parameters()`` names with their shapes.  tests/test_unet_params.py holds the product's table to this one, name by name,
and both to the published parameter count of the SD-1.x U-Net (859 520 964) — so a mistake in one table is not silently shared by
the oracle and the engine.  Parity status: unpinned against the real dependency (diffusers is not installable here).
"""
from __future__ import annotations

from collections import OrderedDict


def _linear(out, name, fin, fout, bias=True):
    out[name + ".weight"] = (fout, fin)
    if bias:
        out[name + ".bias"] = (fout,)


def _conv(out, name, cin, cout, k):
    out[name + ".weight"] = (cout, cin, k, k)
    out[name + ".bias"] = (cout,)


def _norm(out, name, c):
    out[name + ".weight"] = (c,)
    out[name + ".bias"] = (c,)


def _resnet_block_2d(out, name, cin, cout, temb):
    _norm(out, name + ".norm1", cin)
    _conv(out, name + ".conv1", cin, cout, 3)
    _linear(out, name + ".time_emb_proj", temb, cout)
    _norm(out, name + ".norm2", cout)
    _conv(out, name + ".conv2", cout, cout, 3)
    if cin != cout:                                   # use_in_shortcut
        _conv(out, name + ".conv_shortcut", cin, cout, 1)


def _attention(out, name, query_dim, cross_dim):
    _linear(out, name + ".to_q", query_dim, query_dim, bias=False)
    _linear(out, name + ".to_k", cross_dim, query_dim, bias=False)
    _linear(out, name + ".to_v", cross_dim, query_dim, bias=False)
    _linear(out, name + ".to_out.0", query_dim, query_dim)     # to_out = [Linear, Dropout]


def _basic_transformer_block(out, name, dim, cross_dim):
    _norm(out, name + ".norm1", dim)
    _attention(out, name + ".attn1", dim, dim)
    _norm(out, name + ".norm2", dim)
    _attention(out, name + ".attn2", dim, cross_dim)
    _norm(out, name + ".norm3", dim)
    _linear(out, name + ".ff.net.0.proj", dim, 2 * 4 * dim)     # GEGLU: Linear(dim, 2 * inner), inner = 4 * dim
    _linear(out, name + ".ff.net.2", 4 * dim, dim)


def _transformer_2d(out, name, channels, cross_dim):
    _norm(out, name + ".norm", channels)
    _conv(out, name + ".proj_in", channels, channels, 1)        # use_linear_projection = False
    _basic_transformer_block(out, name + ".transformer_blocks.0", channels, cross_dim)
    _conv(out, name + ".proj_out", channels, channels, 1)


def unet_named_parameters(block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, cross_attention_dim=768, in_channels=4,
                          out_channels=4, down_block_types=("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",),
                          up_block_types=("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3) -> "OrderedDict[str, tuple]":
    ch = list(block_out_channels)
    temb = 4 * ch[0]
    out = OrderedDict()
    _conv(out, "conv_in", in_channels, ch[0], 3)
    _linear(out, "time_embedding.linear_1", ch[0], temb)
    _linear(out, "time_embedding.linear_2", temb, temb)
    # down blocks
    output_channel = ch[0]
    for i, kind in enumerate(down_block_types):
        input_channel, output_channel = output_channel, ch[i]
        is_final = i == len(ch) - 1
        for j in range(layers_per_block):
            _resnet_block_2d(out, f"down_blocks.{i}.resnets.{j}", input_channel if j == 0 else output_channel, output_channel, temb)
            if kind == "CrossAttnDownBlock2D":
                _transformer_2d(out, f"down_blocks.{i}.attentions.{j}", output_channel, cross_attention_dim)
        if not is_final:
            _conv(out, f"down_blocks.{i}.downsamplers.0.conv", output_channel, output_channel, 3)
    # mid block: resnet, attention, resnet
    _resnet_block_2d(out, "mid_block.resnets.0", ch[-1], ch[-1], temb)
    _transformer_2d(out, "mid_block.attentions.0", ch[-1], cross_attention_dim)
    _resnet_block_2d(out, "mid_block.resnets.1", ch[-1], ch[-1], temb)
    # up blocks
    rev = list(reversed(ch))
    output_channel = rev[0]
    for i, kind in enumerate(up_block_types):
        prev_output_channel, output_channel = output_channel, rev[i]
        input_channel = rev[min(i + 1, len(ch) - 1)]
        is_final = i == len(ch) - 1
        for j in range(layers_per_block + 1):
            res_skip = input_channel if j == layers_per_block else output_channel
            resnet_in = prev_output_channel if j == 0 else output_channel
            _resnet_block_2d(out, f"up_blocks.{i}.resnets.{j}", resnet_in + res_skip, output_channel, temb)
            if kind == "CrossAttnUpBlock2D":
                _transformer_2d(out, f"up_blocks.{i}.attentions.{j}", output_channel, cross_attention_dim)
        if not is_final:
            _conv(out, f"up_blocks.{i}.upsamplers.0.conv", output_channel, output_channel, 3)
    _norm(out, "conv_norm_out", ch[0])
    _conv(out, "conv_out", ch[0], out_channels, 3)
    return out


def count(params) -> int:
    n = 0
    for shp in params.values():
        k = 1
        for d in shp:
            k *= d
        n += k
    return n
