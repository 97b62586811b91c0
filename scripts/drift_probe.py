"""Where does the 50-step drift come from?  Tiny U-Net configuration, PNDM, guidance 7.5 (tests/test_unet_gpu.py::test_fifty_step...).
Per step: (a) error of the engine's eps for the ORACLE's latents of that step (per-call error along the trajectory, no accumulation),
(b) the same for torch-eager bf16, (c) accumulated error of both free-running trajectories."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import unet_oracle as U
from uce_b200.generate import Denoiser
from uce_b200.schedulers import make_plan
from uce_b200.unet import UNetEngine, cfg_step
from uce_b200.unet_spec import tiny_config

cfg = tiny_config(ch=(64, 128), ctx_dim=64, heads=4, groups=8)
P = U.random_weights(cfg, seed=5)
Pb = {k: v.to(torch.bfloat16) for k, v in P.items()}
g = torch.Generator().manual_seed(31)
lat = torch.randn(2, 4, 16, 16, generator=g); ctx = torch.randn(4, 77, cfg["cross_attention_dim"], generator=g)
gs, steps = 7.5, 50
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
eng = UNetEngine(cfg, batch=4, H=16, W=16); eng.load_state_dict(P); eng.finalize(); eng.set_context(ctx.cuda())
sch, schb = U.PNDMOracle(steps), U.PNDMOracle(steps)
x, xb = lat.clone(), lat.to(torch.bfloat16)
den = Denoiser(eng, 2)
# engine free-running trajectory, step by step (same code as Denoiser.run, unrolled to read the latents after every step)
den.x.copy_(lat.cuda()); hist = []; free = list(den.hist)
plans = list(make_plan("pndm", steps))
for i, t in enumerate(sch.timesteps):
    x2 = torch.cat([x, x])
    e_ref = U.unet_forward(P, x2, int(t), ctx, cfg)
    e_eng = eng.forward(x2.cuda(), float(t), None).cpu()
    e_bf = U.unet_forward(Pb, x2.to(torch.bfloat16), int(t), ctx.to(torch.bfloat16), cfg).float()
    ge = lambda e: U.cfg_combine(e, gs)
    # free-running engine
    plan = plans[i]
    den.x2[:2].copy_(den.x); den.x2[2:].copy_(den.x)
    eng.forward(den.x2, float(plan.t), None, out=den.eps2)
    if plan.save_sample: den.saved.copy_(den.x)
    x_in = den.saved if plan.use_saved_sample else den.x
    if plan.append_eps:
        buf = free.pop() if free else hist.pop(); eps_out = buf
    else:
        buf, eps_out = None, den.scratch
    cfg_step(den.eps2, gs, x_in, den.x, plan.coeffs, plan.cx, plan.ce, hist=hist[:3], eps_out=eps_out)
    if buf is not None:
        hist.insert(0, buf)
        if len(hist) > 3: free.append(hist.pop())
    # free-running eager bf16
    eb = U.unet_forward(Pb, torch.cat([xb, xb]), int(t), ctx.to(torch.bfloat16), cfg)
    xb = schb.step(U.cfg_combine(eb, gs), int(t), xb)
    x = sch.step(ge(e_ref), int(t), x)
    if i % 5 == 0 or i >= 48:
        print(f"step {i:2d} t {int(t):4d} |x| {float(x.std()):7.3f}  per-call eps err: engine {rel(e_eng, e_ref):.4f} eager {rel(e_bf, e_ref):.4f}  guided: engine {rel(ge(e_eng), ge(e_ref)):.4f} eager {rel(ge(e_bf), ge(e_ref)):.4f}"
              f"  trajectory err: engine {rel(den.x.cpu(), x):.4f} eager {rel(xb.float(), x):.4f}", flush=True)
eng.close()
