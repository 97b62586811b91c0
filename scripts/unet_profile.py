"""Two eager SD-1.4-shaped U-Net calls; the second one is bracketed by cudaProfilerStart/Stop (for an ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uce_b200.synthetic import unet_random_state
from uce_b200.unet import UNetEngine
from uce_b200.unet_spec import SD14
state = unet_random_state(SD14, seed=0)
eng = UNetEngine(SD14, batch=2, H=64, W=64)
eng.load_state_dict(state); eng.finalize()
x = torch.randn(2, 4, 64, 64).cuda(); ctx = torch.randn(2, 77, 768).cuda()
eng.set_context(ctx)                  # per-prompt work (cross-attention K / V^T of the text context): outside the step
out = eng.forward(x, 481.0, None)
torch.cuda.synchronize()
torch.cuda.profiler.start()          # ncu --profile-from-start off: exactly one denoise-step U-Net call is listed
out = eng.forward(x, 481.0, None)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(out.abs().mean()))
