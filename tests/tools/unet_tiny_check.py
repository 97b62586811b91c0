"""Per-tap comparison of the CUDA U-Net engine against the CPU oracle on a tiny configuration (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import unet_oracle as U
from uce_b200.unet import UNetEngine
from uce_b200.unet_spec import tiny_config

hw = int(sys.argv[1]) if len(sys.argv) > 1 else 16
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
print("launches per forward:", eng.launch_count(), flush=True)
out = eng.forward(x.cuda(), 481.0, ctx.cuda())
torch.cuda.synchronize()
out = out.cpu()
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
for name in ["temb", "conv_in", "down.0.0", "down.0.1", "down.1.0", "down.1.1", "mid", "up.0.0", "up.0.1", "up.0.2", "up.1.0", "up.1.1", "up.1.2"]:
    t = eng.read_tap(name)
    print(f"{name:10s} shape {tuple(t.shape)} rel err {rel(t, taps[name]):.4e}  (rms ref {taps[name].pow(2).mean().sqrt():.3f})", flush=True)
print(f"eps        rel err {rel(out, ref):.4e} finite={bool(torch.isfinite(out).all())}")
