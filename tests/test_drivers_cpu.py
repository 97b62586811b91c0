"""Host logic of the reference-facing drivers (uce_b200.erase.UCE, uce_b200.debias.UCE) WITHOUT a GPU: the CUDA solver is replaced
by a stand-in that evaluates the same closed form in fp64 on the CPU, so what is checked here is everything around the solve —
module discovery, concept rows, guide broadcast, the cumulative debias targets (uce_sd_debias.py:122-127), the stop rule, the number
of generation rounds, the artifact keys — against the golden outputs of the REAL reference (tests/golden/, oracle/make_golden.py).
The GPU twin (tests/test_drivers_gpu.py) runs the same drivers on the CUDA solver."""
import os

import pytest
import torch

from oracle import uce_oracle as O
from oracle.fake_pipe import ScriptedClip
from tests import golden_util as GU


class ClosedFormSolver:
    """Same call surface as uce_b200.solver.EditSolver.edit; W_new = W_old (lam I + G^T S C)(lam I + C^T S C)^-1 in fp64
    (uce_sd_erase.py:58-82 with the shared-M identity of DESIGN.md 1)."""

    def __init__(self):
        self.calls = 0

    def edit(self, C, G, scales, n_edit, lamb, w_old):
        self.calls += 1
        Cd = C.double().cpu()
        Gf = torch.cat([G.double().cpu(), Cd[n_edit:]])
        S = torch.tensor(scales, dtype=torch.float64)
        eye = torch.eye(Cd.shape[1], dtype=torch.float64)
        M = (lamb * eye + Gf.T @ (S[:, None] * Cd)) @ torch.linalg.inv(lamb * eye + Cd.T @ (S[:, None] * Cd))
        return [(w.double().cpu() @ M).float() for w in w_old]

    def close(self):
        pass


@pytest.mark.parametrize("name", GU.names("debias"))
@pytest.mark.parametrize("through_generator", [False, True])
def test_debias_driver_host_logic(name, through_generator, tmp_path):
    from uce_b200.artifact import load_artifact
    from uce_b200.debias import UCE
    meta, pipe, ref = GU.load(name)
    clip = ScriptedClip(meta["script"], meta["edit"])
    generator = None
    if through_generator:          # any object with the pipeline's face may generate (EngineGenerator on a GPU box): here a pass-through

        class PassThrough:
            def __init__(self, p):
                self.p, self.unet, self.rounds = p, p.unet, 0

            def __call__(self, *a, **k):
                self.rounds += 1
                return self.p(*a, **k)

        generator = PassThrough(pipe)
    solver = ClosedFormSolver()
    UCE(pipe, clip, meta["edit"], meta["debias"], meta["preserve"], meta["edit_scale"], meta["preserve_scale"], meta["lamb"],
        str(tmp_path), "out", meta["max_diff"], 0.1, meta["n_img"], 20, 7.5,
        max_iterations=meta["max_iterations"], desired_ratios=meta["desired"], device="cpu", solver=solver, verbose=False, generator=generator)
    got = load_artifact(os.path.join(tmp_path, "out.safetensors"))
    assert set(got) == set(ref)
    assert len(pipe.calls) == meta["n_pipe_calls"]          # same number of generation rounds as the reference
    if through_generator:
        assert generator.rounds == meta["n_pipe_calls"]
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
        assert got[k].dtype == torch.float32 and got[k].shape == r.shape
        assert O.rel_fro(got[k], e) <= 2e-6, (k, O.rel_fro(got[k], e))                       # fp64 stand-in: only the final fp32 rounding
        assert O.rel_fro(got[k], r) <= O.rel_fro(r, e) + 2e-6, (k, O.rel_fro(got[k], r), O.rel_fro(r, e))


@pytest.mark.parametrize("name", GU.names("erase"))
def test_erase_driver_host_logic(name, tmp_path):
    from uce_b200.artifact import load_artifact
    from uce_b200.erase import UCE
    meta, pipe, ref = GU.load(name)
    solver = ClosedFormSolver()
    UCE(pipe, meta["edit"], meta["guide"], meta["preserve"], meta["erase_scale"], meta["preserve_scale"], meta["lamb"],
        str(tmp_path), "out", device="cpu", verbose=False, solver=solver)
    assert solver.calls == 1                                  # one factor + apply for all projections
    got = load_artifact(os.path.join(tmp_path, "out.safetensors"))
    assert set(got) == set(ref)
    ws = dict(pipe.weights())
    ce, cg, cp = GU.rows(pipe, meta["edit"]), GU.rows(pipe, meta["guide"]), GU.rows(pipe, meta["preserve"])
    for k, r in ref.items():
        w = ws[k[: -len(".weight")]]
        e = O.erase_exact_f64([w], ce, cg, cp, meta["erase_scale"], meta["preserve_scale"], meta["lamb"])[0]
        assert O.rel_fro(got[k], e) <= 2e-6
        assert O.rel_fro(got[k], r) <= O.rel_fro(r, e) + 2e-6
