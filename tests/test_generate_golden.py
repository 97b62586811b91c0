"""Row loop of the generation driver against the REAL reference: tests/golden/generate_calls.json records every
``pipe(...)`` call and every file the reference's own ``generate_images()`` (evalscripts/generate-images-sd.py:10-46) made for a small
CSV under several from_case / till_case windows (oracle/make_generate_golden.py runs it unmodified with a recording pipeline).  Our
``generate_images`` must hand the same prompts, seeds, step counts, guidance and image counts to the denoise loop, in the same order,
overlay the same UCE weights with strict=False, and write the same files — also when the rows are dealt over ranks.  CPU: the U-Net
engine and the denoise loop are stand-ins (the GPU twin is tests/test_unet_gpu.py::test_generate_images_drop_in)."""
import json
import os

import pytest
import torch

from oracle.fake_pipe import FakeGenPipe
from oracle.make_generate_golden import write_csv
from uce_b200 import unet_spec as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "generate_calls.json")))
KEY = "down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight"


class StubEngine:
    def __init__(self):
        self.loads = []

    def load_state_dict(self, state, strict=True):
        self.loads.append((sorted(state), strict))


class StubDenoiser:
    def __init__(self):
        self.calls = []

    def run(self, latents, ctx, steps=50, guidance_scale=7.5, scheduler="pndm"):
        self.calls.append(dict(lat=latents.clone(), ctx=ctx.clone(), steps=steps, gs=guidance_scale, scheduler=scheduler))
        return latents.float()


class RecordingPipe(FakeGenPipe):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.prompts = []

    def encode_prompt(self, prompt, **kw):
        self.prompts.append(prompt)
        return super().encode_prompt(prompt, **kw)


def _run(tmp_path, window, with_weights, rank=0, world=1, monkeypatch=None):
    from safetensors.torch import save_file
    from uce_b200.generate import generate_images
    cfg = U.tiny_config(ch=(64, 128), ctx_dim=64, heads=4, groups=8)
    pipe = RecordingPipe(cfg, {}, latent_size=16)
    csv = tmp_path / "p.csv"
    write_csv(str(csv))
    uce = None
    if with_weights:
        uce = str(tmp_path / "uce.safetensors")
        save_file({KEY: torch.zeros(2, 2)}, uce)
    if monkeypatch is not None:
        monkeypatch.setenv("RANK", str(rank)); monkeypatch.setenv("WORLD_SIZE", str(world))
    eng, den = StubEngine(), StubDenoiser()
    generate_images("some/model", uce, str(csv), str(tmp_path / f"out{rank}"), exp_name="exp", device="cpu", guidance_scale=6.5,
                    num_inference_steps=9, num_images_per_prompt=window["n"], from_case=window["from_case"], till_case=window["till_case"],
                    pipe=pipe, unet_config=cfg, engine=eng, denoiser=den)
    folder = tmp_path / f"out{rank}" / "exp"
    return pipe, eng, den, sorted(os.listdir(folder))


@pytest.mark.parametrize("i", range(len(GOLD["windows"])))
def test_row_loop_matches_the_reference(i, tmp_path):
    g = GOLD["windows"][i]
    pipe, eng, den, files = _run(tmp_path, g["window"], with_weights=g["unet_loaded"] is not None)
    ref = g["calls"]
    assert files == g["files"]
    assert pipe.prompts == [c["prompt"] for c in ref] and all(isinstance(p, str) for p in pipe.prompts)      # str(row.prompt), file order
    assert len(den.calls) == len(ref)
    n = g["window"]["n"]
    for ours, c in zip(den.calls, ref):
        assert (ours["steps"], ours["gs"]) == (c["steps"], c["guidance_scale"]) and c["n"] == n
        want = torch.randn((n, 4, 16, 16), generator=torch.Generator().manual_seed(c["seed"]), dtype=torch.bfloat16)
        assert torch.equal(ours["lat"], want)                     # CPU generator seeded with evaluation_seed (:41), pipeline dtype
        assert ours["ctx"].shape[0] == 2 * n and c["generator_device"] == "cpu"
    if g["unet_loaded"] is not None:                              # load_state_dict(uce_weights, strict=False) (:17-19): a subset dict
        assert g["unet_loaded"] == [[KEY], False]
        assert len(eng.loads) == 1 and KEY in eng.loads[0][0] and eng.loads[0][1] is False
    assert g["from_pretrained"]["torch_dtype"] == "torch.bfloat16" and g["from_pretrained"]["safety_checker"] is None


@pytest.mark.parametrize("world", [2, 3])
def test_rows_dealt_over_ranks_cover_the_reference_calls(world, tmp_path, monkeypatch):
    g = GOLD["windows"][0]
    ref = [(c["prompt"], c["seed"]) for c in g["calls"]]
    got, files = [], set()
    for rank in range(world):
        pipe, _, den, f = _run(tmp_path, g["window"], with_weights=False, rank=rank, world=world, monkeypatch=monkeypatch)
        mine = ref[rank::world]
        assert pipe.prompts == [p for p, _ in mine]               # rank r takes every world-th surviving row, in order
        got += pipe.prompts
        files |= set(f)
    assert sorted(got) == sorted(p for p, _ in ref)
    assert sorted(files) == g["files"]


def test_engine_config_is_read_off_the_pipeline():
    """generate_images() derives the engine configuration and the latent size from pipe.unet.config (the reference accepts any
    --model_id, evalscripts/generate-images-sd.py:13-15): the SD-1.4 config maps to the engine's SD14 table, an SDXL-style U-Net is
    refused with a message instead of failing on a weight shape."""
    import types
    import pytest
    from uce_b200.generate import unet_config_of
    from uce_b200.unet_spec import SD14
    sd14 = dict(sample_size=64, in_channels=4, out_channels=4, layers_per_block=2, block_out_channels=[320, 640, 1280, 1280],
                down_block_types=["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"], up_block_types=["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3,
                cross_attention_dim=768, attention_head_dim=8, norm_num_groups=32, use_linear_projection=False)
    pipe = types.SimpleNamespace(unet=types.SimpleNamespace(config=sd14))
    cfg, latent = unet_config_of(pipe)
    assert latent == 64 and all(cfg[k] == (tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in SD14.items())
    pipe.unet.config = types.SimpleNamespace(**{**sd14, "sample_size": 96})          # attribute-style config (diffusers FrozenDict)
    assert unet_config_of(pipe)[1] == 96
    sdxl = {**sd14, "block_out_channels": [320, 640, 1280], "down_block_types": ["DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"],
            "up_block_types": ["CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"], "transformer_layers_per_block": [1, 2, 10],
            "use_linear_projection": True, "addition_embed_type": "text_time", "cross_attention_dim": 2048}
    pipe.unet.config = sdxl
    with pytest.raises(NotImplementedError):
        unet_config_of(pipe)
    assert unet_config_of(types.SimpleNamespace(unet=types.SimpleNamespace(), latent_size=16))[1] == 16      # synthetic pipes: defaults
