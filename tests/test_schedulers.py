"""CPU: the host scheduler plans (what the fused CUDA update consumes) reproduce the oracle's PNDM/PLMS and DDIM
stepping on arbitrary eps sequences; generation CLI keeps the reference's flags."""
import importlib.util
import os

import pytest
import torch

from oracle import unet_oracle as U
from uce_b200.schedulers import make_plan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,steps", [("pndm", 50), ("pndm", 20), ("pndm", 5), ("ddim", 50), ("ddim", 7)])
def test_plan_matches_oracle_stepping(name, steps):
    orc = U.PNDMOracle(steps) if name == "pndm" else U.DDIMOracle(steps)
    plans = make_plan(name, steps)
    assert [p.t for p in plans] == [int(t) for t in orc.timesteps]
    g = torch.Generator().manual_seed(steps)
    x_ref = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
    x, saved, hist = x_ref.clone(), None, []
    for p in plans:
        eps = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
        x_ref = orc.step(eps, p.t, x_ref)
        if p.save_sample:
            saved = x.clone()
        x_in = saved if p.use_saved_sample else x
        e = p.coeffs[0] * eps
        for c, h in zip(p.coeffs[1:], hist[:3]):
            e = e + c * h
        x = p.cx * x_in + p.ce * e
        if p.append_eps:
            hist.insert(0, eps)
            hist = hist[:4]
        assert torch.allclose(x, x_ref, rtol=1e-9, atol=1e-9), (name, p.t)
    n_calls = steps + 1 if name == "pndm" else steps
    assert len(plans) == n_calls


def test_generate_cli_flags_and_defaults():
    spec = importlib.util.spec_from_file_location("cli_gen", os.path.join(ROOT, "evalscripts", "generate-images-sd.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    a = m.build_parser().parse_args(["--prompts_path", "x.csv"])
    assert (a.model_id, a.uce_model_path, a.save_path, a.device, a.exp_name) == ("CompVis/stable-diffusion-v1-4", None, "../uce_results/", "cuda:0", "test_images")
    assert (a.guidance_scale, a.till_case, a.from_case, a.num_images_per_prompt, a.num_inference_steps) == (7.5, 1000000, 0, 1, 50)
    with pytest.raises(SystemExit):
        m.build_parser().parse_args([])


def test_unet_param_inventory_matches_sd14():
    from uce_b200.unet_spec import SD14, param_count, param_shapes
    assert param_count(SD14) == 859520964
    ks = [k for k in param_shapes(SD14) if "attn2" in k and k.endswith(("to_k.weight", "to_v.weight"))]
    assert len(ks) == 32
