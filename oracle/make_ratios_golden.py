"""TEST INFRASTRUCTURE — golden vectors for ``get_ratios`` of the debias edit.

Calls the reference's own function (trainscripts/uce_sd_debias.py:14-35), unmodified, with the keyword form its caller uses (:96-107), on
a recording pipeline and a scripted classifier: what it loads into ``pipe.unet`` (keys, strict flag), every ``pipe(...)`` call with its
keyword arguments, and the returned direction scales are written to tests/golden/get_ratios.json; tests/test_debias_ratios.py holds
``uce_b200.debias.get_ratios`` to them through the same call.  Only usable where /root/reference is mounted.

    python -m oracle.make_ratios_golden
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from oracle.ref_harness import _load, reference_available

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "get_ratios.json")
EDIT = ["doctor", "nurse", "teacher", "ceo", "chef"]
CASES = [
    dict(debias=["male", "female"], desired=[0.5, 0.5], max_diff=0.05, n_img=10,
         labels={"doctor": ["male"] * 8 + ["female"] * 2, "nurse": ["female"] * 9 + ["male"], "teacher": ["male"] * 5 + ["female"] * 5,
                 "ceo": ["male"] * 10, "chef": ["female"] * 6 + ["male"] * 4}),
    dict(debias=["male", "female"], desired=[0.3, 0.7], max_diff=0.15, n_img=10,
         labels={"doctor": ["male"] * 8 + ["female"] * 2, "nurse": ["female"] * 9 + ["male"], "teacher": ["male"] * 4 + ["female"] * 6,
                 "ceo": ["male"] * 3 + ["female"] * 7, "chef": ["female"] * 6 + ["male"] * 4}),
    dict(debias=["white", "black", "asian"], desired=[0.4, 0.3, 0.3], max_diff=0.1, n_img=6,
         labels={"doctor": ["white"] * 6, "nurse": ["white", "black", "asian", "white", "black", "asian"], "teacher": ["asian"] * 3 + ["white"] * 3,
                 "ceo": ["white", "white", "white", "black", "black", "asian"], "chef": ["black"] * 5 + ["white"]}),
]


class Images:
    def __init__(self, images):
        self.images = images


class RecordingPipe:
    def __init__(self):
        self.calls, self.loaded, self.unet = [], [], self

    def load_state_dict(self, state, strict=True):
        self.loaded.append(dict(keys=list(state), strict=strict, shapes=[list(v.shape) for v in state.values()]))

    def __call__(self, prompt, **kw):
        self.calls.append(dict(prompt=prompt, kwargs={k: kw[k] for k in sorted(kw)}))
        return Images([(prompt, i) for i in range(kw["num_images_per_prompt"])])


def scripted_clip(labels):
    def clip(images, candidate_labels):
        return [[{"label": lab, "score": 0.9}] for lab in labels[images[0][0]]]
    return clip


def modules():
    torch.manual_seed(0)
    names = ["down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k", "mid_block.attentions.0.transformer_blocks.0.attn2.to_v"]
    return names, [torch.nn.Linear(8, 4, bias=False), torch.nn.Linear(8, 6, bias=False)]


def main():
    if not reference_available():
        raise SystemExit("the reference tree is not mounted: fixtures can only be regenerated in the build container")
    mod = _load("uce_sd_debias.py", "_ref_uce_sd_debias_ratios")
    out = []
    for case in CASES:
        pipe = RecordingPipe()
        names, mods = modules()
        r = mod.get_ratios(pipe=pipe, clip=scripted_clip(case["labels"]), uce_module_names=names, uce_modules=mods, edit_concepts=EDIT,
                           debias_concepts=case["debias"], desired_ratios=case["desired"], max_diff=case["max_diff"], step_size=0.1,
                           num_images_per_prompt=case["n_img"], num_inference_steps=7, guidance_scale=6.5)
        out.append(dict(case=case, edit=EDIT, loaded=pipe.loaded, calls=pipe.calls, result=np.asarray(r).tolist(),
                        result_type=type(r).__name__, result_dtype=str(np.asarray(r).dtype)))
        print(case["debias"], np.asarray(r).shape, type(r).__name__)
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
