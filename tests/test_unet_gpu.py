"""GPU parity of the U-Net denoise engine (C ABI include/sd_unet_b200.h) against the CPU torch restatement
oracle/unet_oracle.py.  bf16 storage with fp32 accumulation: tolerance is relative to each tensor's RMS."""
import pytest
import torch

from oracle import unet_oracle as U

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.mark.parametrize("hw", [16, 32])
def test_tiny_unet_forward_matches_oracle(hw):
    from uce_b200.unet import UNetEngine
    from uce_b200.unet_spec import tiny_config
    cfg = tiny_config(ch=(64, 128), ctx_dim=64, heads=4, groups=8)
    P = U.random_weights(cfg, seed=3)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, hw, hw, generator=g)
    ctx = torch.randn(2, 77, cfg["cross_attention_dim"], generator=g)
    taps = {}
    ref = U.unet_forward(P, x, 481.0, ctx, cfg, taps=taps)
    eng = UNetEngine(cfg, batch=2, H=hw, W=hw)
    eng.load_state_dict(P)
    eng.finalize()
    out = eng.forward(x.cuda(), 481.0, ctx.cuda()).cpu()
    torch.cuda.synchronize()
    report = {}
    for name in ["temb", "conv_in", "down.0.0", "down.0.1", "down.1.0", "down.1.1", "mid", "up.0.0", "up.0.2", "up.1.0", "up.1.2"]:
        report[name] = _rel(eng.read_tap(name), taps[name])
    report["eps"] = _rel(out, ref)
    print("per-tap relative error:", {k: round(v, 4) for k, v in report.items()})
    assert report["temb"] < 2e-2 and report["conv_in"] < 1e-2, report
    assert report["eps"] < 5e-2, report
    assert torch.isfinite(out).all()
    eng.close()
