"""TEST INFRASTRUCTURE — regenerate tests/golden/*.npz from the REAL reference.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden

Every fixture holds the outputs of the reference's own ``UCE()`` (erase:
trainscripts/uce_sd_erase.py:12, debias: trainscripts/uce_sd_debias.py:37) executed
through oracle/ref_harness.py on a seeded FakePipe, plus everything needed to
rebuild the inputs on a box without the reference (pipe spec + concept lists; a
sha256 of the regenerated inputs guards against RNG drift).
"""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np
import torch

from .fake_pipe import FakePipe, ScriptedClip, layer_table
from .ref_harness import reference_available, run_reference_debias, run_reference_erase

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

ERASE_CASES = {
    # BASELINE.json configs[0]: 2 concepts, preserve 3, one [320,768] attn2 (to_k + to_v)
    "erase_cfg1": dict(pipe=dict(kind="tiny:320", k_dim=768, seed=11, correlated=False),
                       edit=["Van Gogh", "Picasso"], guide=["art", "art"],
                       preserve=["Monet", "Rembrandt", "Andy Warhol"],
                       erase_scale=1.0, preserve_scale=1.0, lamb=0.5),
    "erase_small_guided": dict(pipe=dict(kind="tiny:24,40", k_dim=64, seed=3, correlated=True),
                               edit=["french horn", "golf ball", "church"], guide=["trumpet", "tennis ball", ""],
                               preserve=["parachute", "gas pump"],
                               erase_scale=2.0, preserve_scale=0.75, lamb=0.1),
    # duplicates are summed twice (uce_sd_erase.py:66-79) although encoded once (:27-28)
    "erase_small_dups": dict(pipe=dict(kind="tiny:16", k_dim=32, seed=5, correlated=True),
                             edit=["cat", "cat", "a cat"], guide=["", "", ""],
                             preserve=["dog", "dog"],
                             erase_scale=1.0, preserve_scale=1.0, lamb=0.5),
    "erase_small_nopreserve": dict(pipe=dict(kind="tiny:8,8,8", k_dim=48, seed=9, correlated=False),
                                   edit=["Kelly McKernan"], guide=["art"], preserve=[],
                                   erase_scale=1.0, preserve_scale=1.0, lamb=0.5),
    # more concepts than K: the primal (K x K) branch of the solver
    "erase_small_wide": dict(pipe=dict(kind="tiny:12", k_dim=16, seed=21, correlated=False),
                             edit=[f"artist number {i}" for i in range(20)], guide=["art"] * 20,
                             preserve=[f"thing {i}" for i in range(15)],
                             erase_scale=1.0, preserve_scale=1.0, lamb=0.5),
}

_G = ["male", "female"]
DEBIAS_CASES = {
    "debias_small": dict(pipe=dict(kind="tiny:24,16", k_dim=48, seed=17, correlated=True),
                         edit=["doctor", "nurse", "ceo"], debias=_G, preserve=["tree", "car"],
                         desired=[0.5, 0.5], max_iterations=6, max_diff=0.05, n_img=10,
                         edit_scale=1.0, preserve_scale=1.0, lamb=0.5,
                         script=[
                             {"doctor": ["male"] * 9 + ["female"], "nurse": ["female"] * 10, "ceo": ["male"] * 5 + ["female"] * 5},
                             {"doctor": ["male"] * 7 + ["female"] * 3, "nurse": ["female"] * 8 + ["male"] * 2, "ceo": ["male"] * 6 + ["female"] * 4},
                             {"doctor": ["male"] * 6 + ["female"] * 4, "nurse": ["female"] * 5 + ["male"] * 5, "ceo": ["male"] * 5 + ["female"] * 5},
                             {"doctor": ["male"] * 5 + ["female"] * 5, "nurse": ["female"] * 5 + ["male"] * 5, "ceo": ["male"] * 5 + ["female"] * 5},
                         ]),
    "debias_stop_at_0": dict(pipe=dict(kind="tiny:8", k_dim=16, seed=2, correlated=False),
                             edit=["teacher"], debias=_G, preserve=[],
                             desired=[0.5, 0.5], max_iterations=3, max_diff=0.05, n_img=10,
                             edit_scale=1.0, preserve_scale=1.0, lamb=0.5,
                             script=[{"teacher": ["male"] * 5 + ["female"] * 5}]),
    "debias_hits_max_iter": dict(pipe=dict(kind="tiny:8", k_dim=16, seed=4, correlated=False),
                                 edit=["pilot", "chef"], debias=_G, preserve=["river"],
                                 desired=[0.5, 0.5], max_iterations=2, max_diff=0.05, n_img=10,
                                 edit_scale=1.5, preserve_scale=1.0, lamb=0.5,
                                 script=[{"pilot": ["male"] * 10, "chef": ["male"] * 8 + ["female"] * 2}]),
}


def make_pipe(spec) -> FakePipe:
    return FakePipe(layer_table(spec["kind"], spec["k_dim"]), seed=spec["seed"], correlated=spec["correlated"])


def inputs_digest(pipe: FakePipe, prompts) -> str:
    h = hashlib.sha256()
    for _, w in pipe.weights():
        h.update(w.numpy().tobytes())
    for p in prompts:
        h.update(pipe.token_row(p).numpy().tobytes())
    return h.hexdigest()


def _save(name, meta, tensors):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    arrays = {k: v.numpy() for k, v in tensors.items()}
    np.savez(os.path.join(GOLDEN_DIR, name + ".npz"), __meta__=np.array(json.dumps(meta)), **arrays)
    print(f"{name}: {len(arrays)} tensors, {sum(a.nbytes for a in arrays.values())/1e6:.2f} MB")


def main():
    if not reference_available():
        raise SystemExit("reference not mounted; golden fixtures can only be regenerated in the build container")
    torch.manual_seed(0)
    for name, case in ERASE_CASES.items():
        pipe = make_pipe(case["pipe"])
        out = run_reference_erase(pipe, case["edit"], case["guide"], case["preserve"],
                                  case["erase_scale"], case["preserve_scale"], case["lamb"])
        meta = dict(case, type="erase", digest=inputs_digest(pipe, case["edit"] + case["guide"] + case["preserve"]),
                    reference_commit="28c81ed5", torch=torch.__version__)
        _save(name, meta, out)
    for name, case in DEBIAS_CASES.items():
        pipe = make_pipe(case["pipe"])
        clip = ScriptedClip(case["script"], case["edit"])
        # digest BEFORE the run: get_ratios loads the edited weights into pipe.unet (uce_sd_debias.py:19)
        digest = inputs_digest(pipe, case["edit"] + case["debias"] + case["preserve"])
        out = run_reference_debias(pipe, clip, case["edit"], case["debias"], case["preserve"], case["desired"],
                                   case["max_iterations"], case["edit_scale"], case["preserve_scale"], case["lamb"],
                                   case["max_diff"], num_images_per_prompt=case["n_img"])
        meta = dict(case, type="debias", digest=digest,
                    n_pipe_calls=len(pipe.calls), reference_commit="28c81ed5", torch=torch.__version__)
        _save(name, meta, out)


if __name__ == "__main__":
    main()
