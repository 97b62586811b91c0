"""Host logic of EngineGenerator (uce_b200/generate.py): the pipeline face get_ratios() of the debias edit needs
(trainscripts/uce_sd_debias.py:14-26 — load the current weights into the U-Net, generate num_images_per_prompt images per edit concept)
with the denoise loop handed to the engine.  CPU: the engine and the denoiser are injected stand-ins, so this covers the wiring —
weight overlay, prompt encoding order [uncond | text], latent shape / dtype / seeding, image hand-over, the debias driver option."""
import numpy as np
import pytest
import torch

from oracle.fake_pipe import FakeGenPipe
from uce_b200 import unet_spec as U


class _StubEngine:
    def __init__(self):
        self.loaded = []

    def load_state_dict(self, state, strict=True):
        self.loaded.append((sorted(state), strict))

    def close(self):
        self.closed = True


class _StubDenoiser:
    """Records what the generator hands to the denoise loop and returns the latents shifted by the context mean."""

    def __init__(self):
        self.calls = []

    def run(self, latents, ctx, steps=50, guidance_scale=7.5, scheduler="pndm"):
        self.calls.append(dict(lat=latents.clone(), ctx=ctx.clone(), steps=steps, gs=guidance_scale, scheduler=scheduler))
        return latents.float() + ctx.mean()


def _gen(images=3):
    from uce_b200.generate import EngineGenerator
    cfg = U.tiny_config(ch=(64, 128), ctx_dim=64, heads=4, groups=8)
    pipe = FakeGenPipe(cfg, {}, latent_size=16)
    eng, den = _StubEngine(), _StubDenoiser()
    return EngineGenerator(pipe, images, device="cpu", engine=eng, denoiser=den), pipe, eng, den


def test_call_hands_the_reference_arguments_to_the_denoise_loop():
    gen, pipe, eng, den = _gen(images=3)
    out = gen("a doctor", num_inference_steps=20, num_images_per_prompt=3, guidance_scale=7.5, generator=torch.Generator().manual_seed(11))
    assert len(out.images) == 3 and gen.calls == 1
    c = den.calls[0]
    assert c["steps"] == 20 and c["gs"] == 7.5 and c["scheduler"] == "pndm"           # pipeline-default scheduler (SURVEY 8a checklist 8)
    assert c["lat"].shape == (3, 4, 16, 16) and c["lat"].dtype == torch.bfloat16      # generation dtype of uce_sd_debias.py:90
    want = torch.randn((3, 4, 16, 16), generator=torch.Generator().manual_seed(11), dtype=torch.bfloat16)
    assert torch.equal(c["lat"], want)                                                 # CPU generator, like generate-images-sd.py:41
    text, uncond = pipe.encode_prompt("a doctor", num_images_per_prompt=3)
    assert c["ctx"].shape == (6, 77, 64) and torch.equal(c["ctx"], torch.cat([uncond, text]))    # [uncond | text]
    ref = FakeGenPipe.latents_to_uint8(want.float() + c["ctx"].mean())
    assert all(np.array_equal(np.asarray(a), b) for a, b in zip(out.images, ref))
    with pytest.raises(ValueError):
        gen("a doctor", num_images_per_prompt=4)


def test_weight_overlay_goes_to_the_engine_with_strict_false():
    gen, _, eng, _ = _gen()
    state = {"down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight": torch.zeros(2, 2)}
    gen.unet.load_state_dict(state, strict=False)          # what get_ratios does first (uce_sd_debias.py:17-20)
    assert eng.loaded == [(sorted(state), False)]
    assert gen.to(torch.bfloat16) is gen


def test_get_ratios_through_the_generator():
    """get_ratios() (ratio / dead-band logic of uce_sd_debias.py:28-35) fed by the generator instead of a pipeline."""
    from uce_b200.debias import get_ratios
    gen, _, eng, den = _gen(images=4)
    edit, deb = ["doctor", "nurse"], ["male", "female"]
    script = iter([["male", "male", "male", "female"], ["female", "female", "male", "male"]])       # top-1 labels per edit concept, in call order

    def clip(images, candidate_labels):
        labels = next(script)
        assert len(images) == len(labels) and candidate_labels == deb
        return [[{"label": lab, "score": 0.9}] + [{"label": c, "score": 0.1} for c in candidate_labels if c != lab] for lab in labels]

    r = get_ratios(gen, clip, ["m.to_k"], [torch.zeros(2, 2)], edit, deb, [0.5, 0.5], 0.05, num_images_per_prompt=4, num_inference_steps=5)
    assert eng.loaded == [(["m.to_k.weight"], False)]
    assert len(den.calls) == 2 and den.calls[0]["steps"] == 5
    assert np.allclose(r, [[-0.25, 0.25], [0.0, 0.0]])


def test_debias_cli_generator_flag():
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("cli_debias_gen", os.path.join(root, "trainscripts", "uce_sd_debias.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    base = ["--edit_concepts", "doctor", "--debias_concepts", "male; female"]
    assert m.build_parser().parse_args(base).generator == "pipe"          # the reference's behaviour stays the default
    assert m.build_parser().parse_args(base + ["--generator", "engine"]).generator == "engine"
