"""Oracle of the VAE decode (oracle/vae_oracle.py; groundwork for SURVEY.md 8(f) rank 1): parameter inventory of the published
SD-1.x layout, shapes, fp64-vs-fp32 self-consistency.  CPU only."""
import torch

from oracle import vae_oracle as V
from uce_b200.vae_spec import SD14_VAE, SD14_VAE_DECODER_PARAMS, decoder_param_count, decoder_param_shapes, tiny_vae_config


def test_sd14_decoder_inventory():
    assert decoder_param_count(SD14_VAE) == SD14_VAE_DECODER_PARAMS == 49_490_179
    assert decoder_param_count(SD14_VAE, with_post_quant=True) == 49_490_179 + 20
    s = decoder_param_shapes(SD14_VAE)
    assert s["decoder.conv_in.weight"] == (512, 4, 3, 3) and s["decoder.conv_out.weight"] == (3, 128, 3, 3)
    assert s["decoder.mid_block.attentions.0.to_q.weight"] == (512, 512)
    assert s["decoder.up_blocks.2.resnets.0.conv_shortcut.weight"] == (256, 512, 1, 1)
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in s and "decoder.up_blocks.2.upsamplers.0.conv.weight" in s
    assert all(not k.endswith("time_emb_proj.weight") for k in s)


def test_tiny_decode_shapes_and_precision():
    cfg = tiny_vae_config()
    P = V.random_weights(cfg, seed=1)
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(2, 4, 6, 5, generator=g) * cfg["scaling_factor"]
    taps = {}
    y32 = V.decode(P, lat, cfg, taps=taps)
    assert y32.shape == (2, 3, 12, 10) and taps["mid"].shape == (2, 32, 6, 5) and torch.isfinite(y32).all()
    P64 = {k: v.double() for k, v in P.items()}
    y64 = V.decode(P64, lat.double(), cfg)
    assert float((y32.double() - y64).norm() / y64.norm()) < 1e-5
    u8 = V.to_uint8(y32)
    assert u8.dtype == torch.uint8 and u8.shape == (2, 12, 10, 3)


def test_bf16_storage_error_budget_behind_the_gpu_tolerance():
    """Where the bars of tests/test_vae_gpu.py come from: the decoder run with bf16 weights and activations (what the engine stores;
    torch accumulates in fp32 like the kernels) against the fp32 run on bf16-rounded weights — SD-1.4 configuration, 16 x 16 latents.
    Measured: rel-RMS 1.1e-2, 0.56 uint8 levels mean, none of the channels more than 5 levels off.  The GPU bars (5e-2, 2 % above 3 levels)
    leave a factor of ~5 for the different accumulation orders of the tensor-core kernels."""
    P = V.random_weights(SD14_VAE, seed=5)
    g = torch.Generator().manual_seed(6)
    lat = torch.randn(1, 4, 16, 16, generator=g) * SD14_VAE["scaling_factor"] * 3.0
    keep = ("norm", "bias", "post_quant_conv", "conv_in", "conv_out")
    ref = V.decode({k: (v if any(t in k for t in keep) else v.to(torch.bfloat16).float()) for k, v in P.items()}, lat, SD14_VAE)
    low = V.decode({k: v.to(torch.bfloat16) for k, v in P.items()}, lat.to(torch.bfloat16), SD14_VAE).float()
    rel = float((low - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    d = (V.to_uint8(low).int() - V.to_uint8(ref).int()).abs()
    assert rel < 2e-2, rel
    assert float((d > 3).float().mean()) < 0.005 and float(d.float().mean()) < 1.0, (float((d > 3).float().mean()), float(d.float().mean()))


def test_value_bias_fold_used_by_the_engine_is_exact():
    """csrc/vae_engine.cu produces V^T by a GEMM without the value bias and adds  bo + Wo bv  in the output projection instead (rows of
    the softmax sum to 1).  Same result as the literal attention block, in fp64."""
    import math
    import torch.nn.functional as F
    cfg = tiny_vae_config(ch=(16, 32), groups=4)
    P = {k: v.double() for k, v in V.random_weights(cfg, seed=2).items()}
    p = "decoder.mid_block.attentions.0"
    for k in (p + ".to_v.bias", p + ".to_out.0.bias", p + ".to_q.bias", p + ".to_k.bias"):
        P[k] = P[k] * 25.0                                   # make the biases matter
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 32, 5, 7, generator=g, dtype=torch.float64)
    want = V._attention(P, p, x, cfg["norm_groups"])
    B, C, H, W = x.shape
    h = F.group_norm(x, cfg["norm_groups"], P[p + ".group_norm.weight"], P[p + ".group_norm.bias"], eps=1e-6)
    t = h.reshape(B, C, H * W).transpose(1, 2)
    q = F.linear(t, P[p + ".to_q.weight"], P[p + ".to_q.bias"])
    k = F.linear(t, P[p + ".to_k.weight"], P[p + ".to_k.bias"])
    vt = P[p + ".to_v.weight"] @ t.transpose(1, 2)           # V^T [B, C, L], no bias
    a = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(C), dim=-1) @ vt.transpose(1, 2)
    folded = P[p + ".to_out.0.bias"] + P[p + ".to_out.0.weight"] @ P[p + ".to_v.bias"]
    got = x + F.linear(a, P[p + ".to_out.0.weight"], folded).transpose(1, 2).reshape(B, C, H, W)
    assert float((got - want).abs().max()) < 1e-10
