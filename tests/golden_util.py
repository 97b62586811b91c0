"""Load tests/golden/*.npz (outputs of the real reference, see oracle/make_golden.py)
and rebuild the matching inputs from the recorded FakePipe spec."""
import glob
import json
import os

import numpy as np
import torch

from oracle.fake_pipe import FakePipe, layer_table
from oracle.make_golden import inputs_digest

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(kind):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, f"{kind}_*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["__meta__"]))
    ref = {k: torch.from_numpy(z[k]) for k in z.files if k != "__meta__"}
    spec = meta["pipe"]
    pipe = FakePipe(layer_table(spec["kind"], spec["k_dim"]), seed=spec["seed"], correlated=spec["correlated"])
    prompts = meta["edit"] + meta.get("guide", meta.get("debias")) + meta["preserve"]
    assert inputs_digest(pipe, prompts) == meta["digest"], "synthetic inputs drifted from the ones the fixture was made with"
    return meta, pipe, ref


def rows(pipe, prompts):
    K = pipe.K
    if not prompts:
        return torch.zeros(0, K)
    return torch.stack([pipe.token_row(p) for p in prompts])
