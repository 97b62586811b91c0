"""TEST INFRASTRUCTURE — run the reference's own ``UCE()`` code objects on CPU.

Only usable where /root/reference is mounted (the build container).  Nothing on
the GPU box may call this; the outputs travel as fixtures in tests/golden/
(see make_golden.py).

Recipe: ``diffusers`` is not installed, so a dummy module with a
``DiffusionPipeline`` attribute is placed in ``sys.modules``; the reference script
is loaded by path (its ``__main__`` guard keeps argparse from running); the module
globals its ``UCE()`` reads (``device``, ``torch_dtype`` — uce_sd_erase.py:116-117,
plus ``max_iterations``/``desired_ratios`` — uce_sd_debias.py:213-214) are set
explicitly; the function is then called unmodified with a FakePipe.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import types

import torch

REFERENCE_ROOT = os.environ.get("UCE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "trainscripts", "uce_sd_erase.py"))


def _load(script: str, alias: str):
    if "diffusers" not in sys.modules:
        stub = types.ModuleType("diffusers")
        stub.DiffusionPipeline = object
        sys.modules["diffusers"] = stub
    path = os.path.join(REFERENCE_ROOT, "trainscripts", script)
    spec = importlib.util.spec_from_file_location(alias, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_reference_erase(pipe, edit, guide, preserve, erase_scale=1.0, preserve_scale=1.0, lamb=0.5):
    """Execute the reference uce_sd_erase.UCE() (uce_sd_erase.py:12) → {key: W_new}."""
    from safetensors.torch import load_file

    mod = _load("uce_sd_erase.py", "_ref_uce_sd_erase")
    mod.device = "cpu"
    mod.torch_dtype = torch.float32
    with tempfile.TemporaryDirectory() as td:
        mod.UCE(pipe, list(edit), list(guide), list(preserve), erase_scale, preserve_scale, lamb, td, "ref")
        return load_file(os.path.join(td, "ref.safetensors"))


def run_reference_debias(pipe, clip, edit, debias, preserve, desired_ratios, max_iterations,
                         edit_scale=1.0, preserve_scale=1.0, lamb=0.5, max_diff=0.05,
                         num_images_per_prompt=10, num_inference_steps=20, guidance_scale=7.5):
    """Execute the reference uce_sd_debias.UCE() (uce_sd_debias.py:37) → {key: W_new}."""
    from safetensors.torch import load_file

    mod = _load("uce_sd_debias.py", "_ref_uce_sd_debias")
    mod.device = "cpu"
    mod.torch_dtype = torch.float32
    mod.max_iterations = max_iterations
    mod.desired_ratios = list(desired_ratios)
    with tempfile.TemporaryDirectory() as td:
        mod.UCE(pipe, clip, list(edit), list(debias), list(preserve), edit_scale, preserve_scale, lamb, td, "ref",
                max_diff, 0.1, num_images_per_prompt, num_inference_steps, guidance_scale)
        return load_file(os.path.join(td, "ref.safetensors"))
