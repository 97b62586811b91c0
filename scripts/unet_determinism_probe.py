"""Run the same SD-1.4 U-Net call several times and report how far the outputs are apart (0 = bit-identical), tap by tap: the first
tap that differs between two runs names the block where a race (or a floating-point atomic) sits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uce_b200.synthetic import unet_random_state
from uce_b200.unet import UNetEngine
from uce_b200.unet_spec import SD14
H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = UNetEngine(SD14, batch=2, H=H, W=H)
eng.load_state_dict(unet_random_state(SD14, seed=0)); eng.finalize()
g = torch.Generator().manual_seed(0)
x = torch.randn(2, 4, H, H, generator=g).cuda(); ctx = torch.randn(2, 77, 768, generator=g).cuda()
eng.set_context(ctx)
names = ["temb", "conv_in"] + [f"down.{i}.{j}" for i in range(4) for j in range(2)] + ["mid"] + [f"up.{i}.{j}" for i in range(4) for j in range(3)]
ref = eng.forward(x, 481.0, None).clone()
ref_taps = {n: eng.read_tap(n).clone() for n in names}
for i in range(4):
    out = eng.forward(x, 481.0, None)
    d = (out - ref)
    first = None
    for n in names:
        t = eng.read_tap(n)
        if not torch.equal(t, ref_taps[n]):
            first = (n, float((t.float() - ref_taps[n].float()).norm() / ref_taps[n].float().norm()), float((t != ref_taps[n]).float().mean()))
            break
    print(f"run {i}: eps rel-RMS difference to run 0 = {float(d.norm() / ref.norm()):.3e}, bit-identical = {bool(torch.equal(out, ref))}, first differing tap = {first}")
