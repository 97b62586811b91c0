"""Pin the oracle (oracle/uce_oracle.py) against outputs of the reference's own UCE()."""
import numpy as np
import pytest
import torch

from oracle import uce_oracle as O
from oracle.fake_pipe import ScriptedClip
from tests import golden_util as GU


@pytest.mark.parametrize("name", GU.names("erase"))
def test_erase_port_matches_reference(name):
    meta, pipe, ref = GU.load(name)
    ws = pipe.weights()
    ce, cg, cp = GU.rows(pipe, meta["edit"]), GU.rows(pipe, meta["guide"]), GU.rows(pipe, meta["preserve"])
    port = O.erase_port_f32([w for _, w in ws], ce, cg, cp, meta["erase_scale"], meta["preserve_scale"], meta["lamb"])
    exact = O.erase_exact_f64([w for _, w in ws], ce, cg, cp, meta["erase_scale"], meta["preserve_scale"], meta["lamb"])
    assert set(ref) == {n + ".weight" for n, _ in ws}
    for (n, _), p, e in zip(ws, port, exact):
        r = ref[n + ".weight"]
        assert r.dtype == torch.float32 and r.shape == p.shape
        # same arithmetic order in fp32: agreement to rounding noise
        assert O.rel_fro(p, r) < 2e-6, (n, O.rel_fro(p, r))
        # the reference itself sits within fp32 conditioning error of the exact answer
        assert O.rel_fro(r, e) < 5e-3, (n, O.rel_fro(r, e))


def _scales_from_script(meta):
    out = []
    for it in range(meta["max_iterations"]):
        step = meta["script"][min(it, len(meta["script"]) - 1)]
        out.append(O.ratios_port([step[c] for c in meta["edit"]], meta["debias"], meta["desired"], meta["max_diff"]))
    return out


@pytest.mark.parametrize("name", GU.names("debias"))
def test_debias_port_matches_reference(name):
    meta, pipe, ref = GU.load(name)
    ws = pipe.weights()
    ce, cd, cp = GU.rows(pipe, meta["edit"]), GU.rows(pipe, meta["debias"]), GU.rows(pipe, meta["preserve"])
    scales = _scales_from_script(meta)
    port = O.debias_port_f32([w for _, w in ws], ce, cd, cp, scales, meta["edit_scale"], meta["preserve_scale"], meta["lamb"])
    exact = O.debias_exact_f64([w for _, w in ws], ce, cd, cp, scales, meta["edit_scale"], meta["preserve_scale"], meta["lamb"])
    for (n, _), p, e in zip(ws, port, exact):
        r = ref[n + ".weight"]
        assert O.rel_fro(p, r) < 2e-6, (n, O.rel_fro(p, r))
        assert O.rel_fro(r, e) < 5e-3, (n, O.rel_fro(r, e))


def test_ratio_deadband():
    r = O.ratios_port([["a"] * 5 + ["b"] * 5, ["a"] * 10, ["a"] * 6 + ["b"] * 4], ["a", "b"], [0.5, 0.5], 0.05)
    assert np.all(r[0] == 0) and np.allclose(r[1], [-0.5, 0.5]) and np.allclose(r[2], [-0.1, 0.1])
    # 0.52/0.48 style inside the dead-band
    r = O.ratios_port([["a"] * 13 + ["b"] * 12], ["a", "b"], [0.5, 0.5], 0.05)
    assert np.all(r == 0)
