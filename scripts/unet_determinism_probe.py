"""Run the same SD-1.4 U-Net call several times and report how far the outputs are apart (0 = bit-identical).  GroupNorm statistics
use fp32 atomics, so tiny differences are expected; anything at the 1e-3 level or above points at a race."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uce_b200.synthetic import unet_random_state
from uce_b200.unet import UNetEngine
from uce_b200.unet_spec import SD14
eng = UNetEngine(SD14, batch=2, H=64, W=64)
eng.load_state_dict(unet_random_state(SD14, seed=0)); eng.finalize()
g = torch.Generator().manual_seed(0)
x = torch.randn(2, 4, 64, 64, generator=g).cuda(); ctx = torch.randn(2, 77, 768, generator=g).cuda()
eng.set_context(ctx)
ref = eng.forward(x, 481.0, None).clone()
for i in range(5):
    out = eng.forward(x, 481.0, None)
    d = (out - ref)
    print(f"run {i}: rel-RMS difference to run 0 = {float(d.norm() / ref.norm()):.3e}, max abs = {float(d.abs().max()):.3e}, bit-identical = {bool(torch.equal(out, ref))}")
