"""TEST INFRASTRUCTURE — golden vectors for the generation driver's row loop.

Runs the reference's own ``generate_images()`` (evalscripts/generate-images-sd.py:10-46), unmodified, in the build container: a stub
``diffusers`` module provides ``DiffusionPipeline.from_pretrained`` returning a recording pipeline, so every ``pipe(...)`` call the
reference makes — prompt, generator seed, steps, guidance, images per prompt — and every file it writes is captured for a small CSV
(with a non-string prompt, unsorted and duplicated case numbers) under several from_case / till_case windows.  The records go to
tests/golden/generate_calls.json; tests/test_generate_golden.py holds our ``generate_images`` row loop to them (also split over ranks).
Only usable where /root/reference is mounted.

    python -m oracle.make_generate_golden
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import tempfile
import types

import torch

from oracle.ref_harness import REFERENCE_ROOT, reference_available

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "generate_calls.json")
CSV_ROWS = [  # case_number, prompt, evaluation_seed
    (0, "A Wheatfield, with Cypresses by Vincent van Gogh", 2219),
    (1, "Almond Blossoms by Vincent van Gogh", 4965),
    (3, 1889, 32),                      # a numeric cell: the reference passes str(row.prompt) (:30)
    (2, "a dog", 7),                    # out of order
    (7, "a cat", 11),
    (7, "a cat, again", 12),            # duplicated case number: the second row overwrites the files of the first (:46)
    (40, "The Starry Night", 4000),
]
WINDOWS = [dict(from_case=0, till_case=1000000, n=2), dict(from_case=1, till_case=7, n=1), dict(from_case=3, till_case=3, n=3),
           dict(from_case=50, till_case=60, n=1)]


def write_csv(path):
    import pandas as pd
    pd.DataFrame({"case_number": [r[0] for r in CSV_ROWS], "prompt": [r[1] for r in CSV_ROWS], "evaluation_seed": [r[2] for r in CSV_ROWS]}).to_csv(path)


class _Unet:
    def __init__(self):
        self.loaded = None

    def load_state_dict(self, state, strict=True):
        self.loaded = (sorted(state), strict)


class RecordingPipe:
    def __init__(self):
        self.calls, self.unet = [], _Unet()

    def to(self, device):
        self.device = device
        return self

    def __call__(self, prompt=None, num_inference_steps=None, guidance_scale=None, num_images_per_prompt=None, generator=None):
        from PIL import Image
        self.calls.append(dict(prompt=prompt, prompt_type=type(prompt).__name__, seed=int(generator.initial_seed()),
                               generator_device=str(generator.device), steps=num_inference_steps, guidance_scale=guidance_scale,
                               n=num_images_per_prompt))
        return types.SimpleNamespace(images=[Image.new("RGB", (8, 8), (len(self.calls) % 256, i, 0)) for i in range(num_images_per_prompt)])


def main():
    if not reference_available():
        raise SystemExit("the reference tree is not mounted: fixtures can only be regenerated in the build container")
    pipe_box = {}
    stub = types.ModuleType("diffusers")

    class DiffusionPipeline:
        @staticmethod
        def from_pretrained(model_id, torch_dtype=None, safety_checker="unset"):
            pipe_box["pipe"] = RecordingPipe()
            pipe_box["from_pretrained"] = dict(model_id=model_id, torch_dtype=str(torch_dtype), safety_checker=safety_checker)
            return pipe_box["pipe"]

    stub.DiffusionPipeline = DiffusionPipeline
    old = sys.modules.get("diffusers")
    sys.modules["diffusers"] = stub
    try:
        spec = importlib.util.spec_from_file_location("_ref_generate", os.path.join(REFERENCE_ROOT, "evalscripts", "generate-images-sd.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        out = {"csv_rows": CSV_ROWS, "windows": []}
        from safetensors.torch import save_file
        with tempfile.TemporaryDirectory() as td:
            csv = os.path.join(td, "p.csv"); write_csv(csv)
            key = "down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight"
            save_file({key: torch.zeros(2, 2)}, os.path.join(td, "uce.safetensors"))
            for i, w in enumerate(WINDOWS):
                save = os.path.join(td, f"out{i}")
                mod.generate_images("some/model", os.path.join(td, "uce.safetensors") if i % 2 == 0 else None, csv, save, exp_name="exp",
                                    device="cpu", guidance_scale=6.5, num_inference_steps=9, num_images_per_prompt=w["n"],
                                    from_case=w["from_case"], till_case=w["till_case"])
                pipe = pipe_box["pipe"]
                out["windows"].append(dict(window=w, calls=pipe.calls, files=sorted(os.listdir(os.path.join(save, "exp"))),
                                           unet_loaded=pipe.unet.loaded, from_pretrained=pipe_box["from_pretrained"]))
    finally:
        if old is not None:
            sys.modules["diffusers"] = old
        else:
            del sys.modules["diffusers"]
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    for w in out["windows"]:
        print(w["window"], len(w["calls"]), "calls,", len(w["files"]), "files", w["unet_loaded"])


if __name__ == "__main__":
    main()
