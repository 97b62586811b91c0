"""The reference-facing Python drivers (UCE() for erase and debias) on the CUDA solver,
checked against the golden outputs of the real reference run on the same FakePipe."""
import os

import pytest
import torch

from oracle import uce_oracle as O
from oracle.fake_pipe import ScriptedClip
from tests import golden_util as GU

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GU.names("erase"))
def test_erase_driver(name, tmp_path):
    from safetensors.torch import load_file
    from uce_b200.erase import UCE
    meta, pipe, ref = GU.load(name)
    UCE(pipe, meta["edit"], meta["guide"], meta["preserve"], meta["erase_scale"], meta["preserve_scale"], meta["lamb"],
        str(tmp_path), "out", device="cuda:0", verbose=False)
    got = load_file(os.path.join(tmp_path, "out.safetensors"))
    assert list(got.keys()) == list(ref.keys()) or set(got) == set(ref)
    ws = dict(pipe.weights())
    ce, cg, cp = GU.rows(pipe, meta["edit"]), GU.rows(pipe, meta["guide"]), GU.rows(pipe, meta["preserve"])
    for k, r in ref.items():
        assert got[k].dtype == torch.float32 and got[k].shape == r.shape
        w = ws[k[: -len(".weight")]]
        e = O.erase_exact_f64([w], ce, cg, cp, meta["erase_scale"], meta["preserve_scale"], meta["lamb"])[0]
        assert O.rel_fro(got[k], e) <= 2e-5
        assert O.rel_fro(got[k], r) <= O.rel_fro(r, e) + 2e-5


@pytest.mark.parametrize("name", GU.names("debias"))
def test_debias_driver(name, tmp_path):
    from safetensors.torch import load_file
    from uce_b200.debias import UCE
    meta, pipe, ref = GU.load(name)
    clip = ScriptedClip(meta["script"], meta["edit"])
    UCE(pipe, clip, meta["edit"], meta["debias"], meta["preserve"], meta["edit_scale"], meta["preserve_scale"], meta["lamb"],
        str(tmp_path), "out", meta["max_diff"], 0.1, meta["n_img"], 20, 7.5,
        max_iterations=meta["max_iterations"], desired_ratios=meta["desired"], device="cuda:0", verbose=False)
    got = load_file(os.path.join(tmp_path, "out.safetensors"))
    assert set(got) == set(ref)
    assert len(pipe.calls) == meta["n_pipe_calls"]          # same number of generation rounds as the reference
    scales = []
    for it in range(meta["max_iterations"]):
        step = meta["script"][min(it, len(meta["script"]) - 1)]
        scales.append(O.ratios_port([step[c] for c in meta["edit"]], meta["debias"], meta["desired"], meta["max_diff"]))
    _, pipe0, _ = GU.load(name)             # fresh weights: the debias loop loads edited weights into pipe.unet
    ws = dict(pipe0.weights())
    ce, cd, cp = GU.rows(pipe0, meta["edit"]), GU.rows(pipe0, meta["debias"]), GU.rows(pipe0, meta["preserve"])
    for k, r in ref.items():
        w = ws[k[: -len(".weight")]]
        e = O.debias_exact_f64([w], ce, cd, cp, scales, meta["edit_scale"], meta["preserve_scale"], meta["lamb"])[0]
        assert O.rel_fro(got[k], e) <= 2e-5, (k, O.rel_fro(got[k], e))
        assert O.rel_fro(got[k], r) <= O.rel_fro(r, e) + 2e-5, (k, O.rel_fro(got[k], r), O.rel_fro(r, e))
