"""GPU parity of the U-Net denoise engine (C ABI include/sd_unet_b200.h) against the CPU torch restatement
oracle/unet_oracle.py.  bf16 storage with fp32 accumulation: tolerance is relative to each tensor's RMS."""
import os

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


def _tiny():
    from uce_b200.unet_spec import tiny_config
    cfg = tiny_config(ch=(64, 128), ctx_dim=64, heads=4, groups=8)
    return cfg, U.random_weights(cfg, seed=5)


@pytest.mark.parametrize("sched,steps", [("pndm", 6), ("ddim", 5)])
def test_denoise_loop_matches_oracle(sched, steps):
    """CFG + scheduler + U-Net over several steps (PLMS history, saved-sample special case) vs the fp32 CPU loop."""
    from uce_b200.generate import Denoiser
    from uce_b200.unet import UNetEngine
    cfg, P = _tiny()
    g = torch.Generator().manual_seed(1)
    lat = torch.randn(2, 4, 16, 16, generator=g)
    ctx = torch.randn(4, 77, cfg["cross_attention_dim"], generator=g)
    gs = 2.0     # classifier-free guidance multiplies the bf16 rounding noise of (eps_t - eps_u) by the scale: keep it moderate here
    ref = U.denoise_loop(P, lat, ctx, steps=steps, guidance_scale=gs, scheduler=sched, cfg=cfg)
    eng = UNetEngine(cfg, batch=4, H=16, W=16)
    eng.load_state_dict(P)
    eng.finalize()
    out = Denoiser(eng, 2).run(lat, ctx, steps=steps, guidance_scale=gs, scheduler=sched).cpu()
    err = _rel(out, ref)
    print(sched, "latents rel err after", steps, "steps:", err)
    # one U-Net call is ~1% off the fp32 oracle (bf16 activations, see the tap test); 5-6 scheduler steps with guidance
    # compound that, and the exact figure moves with the summation order of the kernels (0.03-0.065 observed)
    assert err < 0.1, err
    eng.close()


def test_cfg_step_kernel_exact():
    from uce_b200.unet import cfg_step
    g = torch.Generator().manual_seed(2)
    n = 2 * 4 * 16 * 16
    eps2 = torch.randn(2 * n, generator=g).cuda()
    x = torch.randn(n, generator=g).cuda()
    h = [torch.randn(n, generator=g).cuda() for _ in range(3)]
    out, eo = torch.empty_like(x), torch.empty_like(x)
    c = (55 / 24, -59 / 24, 37 / 24, -9 / 24)
    cfg_step(eps2, 7.5, x, out, c, 1.01, -0.02, hist=h, eps_out=eo)
    eps = eps2[:n] + 7.5 * (eps2[n:] - eps2[:n])
    ref = 1.01 * x + (-0.02) * (c[0] * eps + c[1] * h[0] + c[2] * h[1] + c[3] * h[2])
    assert torch.allclose(eo, eps, rtol=1e-6, atol=1e-6) and torch.allclose(out, ref, rtol=1e-5, atol=1e-5)


def test_generate_images_drop_in(tmp_path):
    """generate_images(): CSV rows, seeds, case filter, file naming, strict=False weight overlay — against the oracle loop."""
    import numpy as np
    import pandas as pd
    from PIL import Image
    from safetensors.torch import save_file
    from oracle.fake_pipe import FakeGenPipe
    from uce_b200.generate import generate_images
    cfg, P = _tiny()
    pipe = FakeGenPipe(cfg, P, latent_size=16)
    csv = tmp_path / "p.csv"
    pd.DataFrame({"case_number": [0, 1, 2], "prompt": ["a cat", "Starry Night by Van Gogh", "a dog"], "evaluation_seed": [11, 2219, 5]}).to_csv(csv)
    # an "edited" attn2 weight overlay, as written by the edit solver
    key = "down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight"
    edited = {key: P[key] * 0.5}
    save_file(edited, str(tmp_path / "uce.safetensors"))
    generate_images("unused", str(tmp_path / "uce.safetensors"), str(csv), str(tmp_path), exp_name="out", device="cuda:0",
                    torch_dtype=torch.bfloat16, guidance_scale=2.0, num_inference_steps=4, num_images_per_prompt=2,
                    from_case=1, till_case=2, pipe=pipe, unet_config=cfg)
    files = sorted(os.listdir(tmp_path / "out"))
    assert files == ["1_0.png", "1_1.png", "2_0.png", "2_1.png"]
    P2 = dict(P); P2.update(edited)
    for case, prompt, seed in [(1, "Starry Night by Van Gogh", 2219), (2, "a dog", 5)]:
        text, uncond = pipe.encode_prompt(prompt, num_images_per_prompt=2)
        lat = torch.randn((2, 4, 16, 16), generator=torch.Generator().manual_seed(seed), dtype=torch.bfloat16).float()
        ref = FakeGenPipe.latents_to_uint8(U.denoise_loop(P2, lat, torch.cat([uncond, text]), steps=4, guidance_scale=2.0, cfg=cfg))
        for i in range(2):
            got = np.asarray(Image.open(tmp_path / "out" / f"{case}_{i}.png")).astype(np.int32)
            assert got.shape == ref[i].shape
            assert np.abs(got - ref[i].astype(np.int32)).mean() < 8.0, (case, i, np.abs(got - ref[i]).mean())


def test_sd14_shapes_forward_matches_oracle():
    """The real SD-1.4 U-Net configuration (859.5 M parameters, 8 heads of 40/80/160, 77x768 context) at 32x32 latents."""
    from uce_b200.unet import UNetEngine
    from uce_b200.unet_spec import SD14
    P = U.random_weights(SD14, seed=0)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 32, 32, generator=g)
    ctx = torch.randn(2, 77, 768, generator=g)
    taps = {}
    torch.set_num_threads(8)
    ref = U.unet_forward(P, x, 801.0, ctx, SD14, taps=taps)
    eng = UNetEngine(SD14, batch=2, H=32, W=32)
    eng.load_state_dict(P)
    eng.finalize()
    out = eng.forward(x.cuda(), 801.0, ctx.cuda()).cpu()
    rep = {k: _rel(eng.read_tap(k), taps[k]) for k in ["conv_in", "down.0.1", "down.2.1", "mid", "up.1.2", "up.3.2"]}
    rep["eps"] = _rel(out, ref)
    print("SD-1.4 per-tap relative error:", {k: round(v, 4) for k, v in rep.items()})
    assert rep["eps"] < 5e-2 and torch.isfinite(out).all(), rep
    eng.close()


def test_sd14_full_latent_size_forward_matches_oracle_per_tap():
    """The shape bench.py times and every real row uses: SD-1.4 at 64 x 64 latents (4096-token self-attention, the split-K / tile
    decisions of the 64 x 64 level), NB = 2.  Every tap is held to the bar, not printed."""
    from uce_b200.unet import UNetEngine
    from uce_b200.unet_spec import SD14
    P = U.random_weights(SD14, seed=1)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 4, 64, 64, generator=g)
    ctx = torch.randn(2, 77, 768, generator=g)
    taps = {}
    torch.set_num_threads(max(8, os.cpu_count() or 8))
    ref = U.unet_forward(P, x, 621.0, ctx, SD14, taps=taps)
    eng = UNetEngine(SD14, batch=2, H=64, W=64)
    eng.load_state_dict(P)
    eng.finalize()
    out = eng.forward(x.cuda(), 621.0, ctx.cuda()).cpu()
    names = ["temb", "conv_in", "down.0.0", "down.0.1", "down.1.0", "down.1.1", "down.2.0", "down.2.1", "down.3.1", "mid",
             "up.0.2", "up.1.0", "up.1.2", "up.2.0", "up.2.2", "up.3.0", "up.3.1", "up.3.2"]
    rep = {k: _rel(eng.read_tap(k), taps[k]) for k in names}
    rep["eps"] = _rel(out, ref)
    print("SD-1.4 @64x64 per-tap relative error:", {k: round(v, 4) for k, v in rep.items()})
    assert torch.isfinite(out).all()
    for k, v in rep.items():
        assert v < 5e-2, (k, v, rep)
    again = eng.forward(x.cuda(), 621.0, ctx.cuda()).cpu()
    assert torch.equal(out, again), ("the engine is bit-reproducible run to run", _rel(again, out))
    eng.close()


def test_engine_is_bit_reproducible():
    """Same inputs, same engine: identical bits (GroupNorm statistics are reduced in a fixed order, split-K slabs are summed in
    order, no floating-point atomics) — and identical again from a second engine built from the same weights."""
    from uce_b200.unet import UNetEngine
    cfg, P = _tiny()
    g = torch.Generator().manual_seed(21)
    x = torch.randn(4, 4, 32, 32, generator=g).cuda()
    ctx = torch.randn(4, 77, cfg["cross_attention_dim"], generator=g).cuda()
    outs = []
    for _ in range(2):
        eng = UNetEngine(cfg, batch=4, H=32, W=32)
        eng.load_state_dict(P)
        eng.finalize()
        outs.append([eng.forward(x, 500.0, ctx).clone() for _ in range(3)])
        eng.close()
    for o in outs[0] + outs[1]:
        assert torch.equal(o, outs[0][0]), _rel(o, outs[0][0])


def test_fifty_step_guided_loop_no_worse_than_eager_bf16():
    """The reference's default call (generate-images-sd.py:37-42,58,62): 50 PNDM steps (51 U-Net calls) at guidance scale 7.5.
    north_star asks for latents within 1e-3 max-abs at 16-bit precision; one 16-bit ulp at |x| in [1, 4] is already 4e-3..1.6e-2
    (bf16, the reference's dtype, :76), so after 51 guided calls the honest statement is relative: the engine's drift
    from the fp32 oracle must not exceed what torch-eager bf16 — the same oracle code run on bf16 weights and activations, i.e. what
    the reference's pipeline computes — shows against the same fp32 run.  Both drifts are reported as rel-RMS and max-abs."""
    from uce_b200.generate import Denoiser
    from uce_b200.unet import UNetEngine
    cfg, P = _tiny()
    g = torch.Generator().manual_seed(31)
    lat = torch.randn(2, 4, 16, 16, generator=g)
    ctx = torch.randn(4, 77, cfg["cross_attention_dim"], generator=g)
    gs, steps = 7.5, 50
    ref = U.denoise_loop(P, lat, ctx, steps=steps, guidance_scale=gs, scheduler="pndm", cfg=cfg)
    Pb = {k: v.to(torch.bfloat16) for k, v in P.items()}
    eager = U.denoise_loop(Pb, lat.to(torch.bfloat16), ctx.to(torch.bfloat16), steps=steps, guidance_scale=gs, scheduler="pndm", cfg=cfg).float()
    eng = UNetEngine(cfg, batch=4, H=16, W=16)
    eng.load_state_dict(P)
    eng.finalize()
    out = Denoiser(eng, 2).run(lat, ctx, steps=steps, guidance_scale=gs, scheduler="pndm").cpu()
    eng.close()
    e_rel, e_max = _rel(out, ref), float((out - ref).abs().max())
    b_rel, b_max = _rel(eager, ref), float((eager - ref).abs().max())
    print(f"50-step gs 7.5 drift vs fp32 oracle: engine rel-RMS {e_rel:.4f} max-abs {e_max:.4f} | torch-eager bf16 rel-RMS {b_rel:.4f} max-abs {b_max:.4f}")
    assert torch.isfinite(out).all()
    # "no worse than eager", with the slack two different 16-bit roundings of one chaotic 51-call trajectory need (eager: 0.031 / 3.2)
    assert e_rel <= 1.5 * b_rel, (e_rel, b_rel)
    assert e_max <= 1.5 * b_max, (e_max, b_max)


def test_batch8_images_per_prompt():
    """BASELINE config 5 uses --num_images_per_prompt 8 (16 samples per U-Net call): batch handling of every kernel
    (image index of the time-embedding bias, attention batching, conv rectangles spanning images)."""
    from uce_b200.unet import UNetEngine
    cfg, P = _tiny()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(16, 4, 16, 16, generator=g)
    ctx = torch.randn(16, 77, cfg["cross_attention_dim"], generator=g)
    ref = U.unet_forward(P, x, 321.0, ctx, cfg)
    eng = UNetEngine(cfg, batch=16, H=16, W=16)
    eng.load_state_dict(P)
    eng.finalize()
    out = eng.forward(x.cuda(), 321.0, ctx.cuda()).cpu()
    per_sample = [_rel(out[i], ref[i]) for i in range(16)]
    assert max(per_sample) < 5e-2, per_sample
    eng.close()


def test_edited_weights_overlay_changes_only_attn2():
    """set_weight after finalize (load_state_dict(strict=False) of the UCE artifact) takes effect in place."""
    from uce_b200.unet import UNetEngine
    cfg, P = _tiny()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 4, 16, 16, generator=g)
    ctx = torch.randn(2, 77, cfg["cross_attention_dim"], generator=g)
    eng = UNetEngine(cfg, batch=2, H=16, W=16)
    eng.load_state_dict(P)
    eng.finalize()
    base = eng.forward(x.cuda(), 500.0, ctx.cuda()).cpu()
    key = "mid_block.attentions.0.transformer_blocks.0.attn2.to_v.weight"
    eng.load_state_dict({key: P[key] * -1.0, "not.a.unet.key": torch.zeros(1)}, strict=False)
    P2 = dict(P); P2[key] = P[key] * -1.0
    ref = U.unet_forward(P2, x, 500.0, ctx, cfg)
    out = eng.forward(x.cuda(), 500.0, ctx.cuda()).cpu()
    assert _rel(out, ref) < 5e-2 and _rel(out, base) > 1e-3
    eng.close()


def test_context_cache_matches_per_call_context():
    """sd_unet_set_context + forward(ctx=None) — the per-prompt cache of the cross-attention K / V^T projections
    (generate-images-sd.py:37-42 keeps one prompt embedding over all steps of a row) — gives the same results (to the engine's run-to-run noise) as passing
    the context with every call, follows a new context, and is recomputed after an attn2 weight is overwritten."""
    from uce_b200.unet import UNetEngine
    cfg, P = _tiny()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 4, 16, 16, generator=g).cuda()
    ctx_a = torch.randn(2, 77, cfg["cross_attention_dim"], generator=g).cuda()
    ctx_b = torch.randn(2, 77, cfg["cross_attention_dim"], generator=g).cuda()
    eng = UNetEngine(cfg, batch=2, H=16, W=16)
    eng.load_state_dict(P)
    eng.finalize()
    with pytest.raises(Exception):
        eng.forward(x, 300.0, None)                      # no context yet
    assert eng.context_launch_count() > 0 and eng.launch_count() > 0
    per_call_a = eng.forward(x, 300.0, ctx_a).clone()
    per_call_b = eng.forward(x, 300.0, ctx_b).clone()
    assert torch.equal(eng.forward(x, 300.0, ctx_a), per_call_a)      # the engine is bit-reproducible
    tol = 1e-6                                                         # the cached K / V^T are the same GEMMs on the same inputs
    assert _rel(per_call_a, per_call_b) > 5 * tol
    eng.set_context(ctx_a)
    assert _rel(eng.forward(x, 300.0, None), per_call_a) <= tol
    assert _rel(eng.forward(x, 300.0, None), per_call_a) <= tol          # and again: nothing of the cache is consumed
    eng.set_context(ctx_b)
    assert _rel(eng.forward(x, 300.0, None), per_call_b) <= tol
    key = "mid_block.attentions.0.transformer_blocks.0.attn2.to_k.weight"
    eng.load_state_dict({key: P[key] * 0.5}, strict=False)            # edited weight: the cached K must not survive
    P2 = dict(P); P2[key] = P[key] * 0.5
    ref = U.unet_forward(P2, x.cpu(), 300.0, ctx_b.cpu(), cfg)
    out = eng.forward(x, 300.0, None).cpu()
    assert _rel(out, ref) < 5e-2 and _rel(out, per_call_b.cpu()) > 1e-3
    eng.close()


def test_engine_generator_matches_oracle_loop_and_follows_weight_overlays():
    """EngineGenerator (the generation rounds of the debias edit's get_ratios(), trainscripts/uce_sd_debias.py:14-26, on the U-Net engine):
    same oracle loop and tolerance as generate_images above; loading edited attn2 weights changes what the next call generates, loading
    the originals back restores it (every get_ratios call starts with such an overlay, :17-20)."""
    import numpy as np
    from oracle.fake_pipe import FakeGenPipe
    from uce_b200.generate import EngineGenerator
    cfg, P = _tiny()
    pipe = FakeGenPipe(cfg, P, latent_size=16)
    gen = EngineGenerator(pipe, 2, device="cuda:0", unet_config=cfg)
    key = "down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight"

    def run(seed):
        out = gen("a doctor", num_inference_steps=4, num_images_per_prompt=2, guidance_scale=2.0, generator=torch.Generator().manual_seed(seed))
        return np.stack([np.asarray(im).astype(np.int32) for im in out.images])

    def oracle(weights, seed):
        text, uncond = pipe.encode_prompt("a doctor", num_images_per_prompt=2)
        lat = torch.randn((2, 4, 16, 16), generator=torch.Generator().manual_seed(seed), dtype=torch.bfloat16).float()
        return FakeGenPipe.latents_to_uint8(U.denoise_loop(weights, lat, torch.cat([uncond, text]), steps=4, guidance_scale=2.0, cfg=cfg)).astype(np.int32)

    base = run(3)
    assert np.abs(base - oracle(P, 3)).mean() < 8.0
    gen.unet.load_state_dict({key: P[key] * -1.0}, strict=False)
    P2 = dict(P); P2[key] = P[key] * -1.0
    edited = run(3)
    assert np.abs(edited - oracle(P2, 3)).mean() < 8.0
    assert np.abs(edited - base).mean() > 0.0                      # the overlay reached the kernels (cached context K / V^T recomputed)
    gen.unet.load_state_dict({key: P[key]}, strict=False)
    again = run(3)
    assert np.array_equal(again, base)                             # back to the original weights: the same bits
    gen.close()
