"""Drop-in for the reference's iterative debias edit, ``get_ratios()`` / ``UCE()`` of
trainscripts/uce_sd_debias.py:14-149, with the solve on the B200 solver.

Per iteration the reference regenerates images with the current weights, classifies them,
turns the label fractions into ``direction_scale`` (ratio = desired − observed, dead-band
``max_diff``; :28-35), nudges the edit targets v*_e += sum_j ratio_ej (W_old c_dj) IN PLACE
(cumulative over iterations, :124-126) and re-solves from W_old (:114-140).  In row form the
cumulative target is the guide row  g_e = c_e + sum_j A_ej c_dj  with A the running sum of the
direction scales, which is what is handed to the solver.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from .concepts import embed_concepts, select_projections
from .solver import EditSolver


def _rank_world():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return 0, 1


def get_ratios(pipe, clip, uce_module_names, uce_modules=None, edit_concepts=(), debias_concepts=(), desired_ratios=(), max_diff=0.05,
               step_size=0.1, num_images_per_prompt=10, num_inference_steps=20, guidance_scale=7.5, uce_weights=None):
    """Mirror of uce_sd_debias.py:14-35, same parameter names and order (``step_size`` is accepted and unused, as there).
    ``uce_modules``: the edited projections as the reference passes them (modules with a ``.weight``, :16-17) or their weight tensors;
    ``uce_weights`` is a keyword alias for the latter."""
    if uce_modules is None:
        uce_modules = uce_weights
    if uce_modules is None:
        raise TypeError("get_ratios needs uce_modules (or uce_weights)")
    state = {name + ".weight": (m.weight if hasattr(m, "weight") else m) for name, m in zip(uce_module_names, uce_modules)}
    pipe.unet.load_state_dict(state, strict=False)
    # Multi-GPU: the edit concepts are dealt round-robin to the ranks — each rank generates and classifies only its
    # own — and ONE all-reduce of the [n_edit, n_debias] label-count matrix (plus the image counts) gives every rank the same ratios.
    rank, world = _rank_world()
    counts = np.zeros((len(edit_concepts), len(debias_concepts)), dtype=np.float64)
    totals = np.zeros(len(edit_concepts), dtype=np.float64)
    for i, concept in enumerate(edit_concepts):
        if i % world != rank:
            continue
        out = pipe(concept, num_inference_steps=num_inference_steps, num_images_per_prompt=num_images_per_prompt, guidance_scale=guidance_scale)
        images = out.images
        if getattr(clip, "accepts_device_images", False) and getattr(out, "images_u8", None) is not None:
            images = out.images_u8                 # VAE engine -> classifier engine: the pixels never leave the device
        results = clip(images, candidate_labels=debias_concepts)
        top1 = np.array([r[0]["label"] for r in results])
        counts[i] = [np.sum(top1 == c) for c in debias_concepts]
        totals[i] = len(top1)
    if world > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.distributed.get_backend() == "nccl" else torch.device("cpu")
        t = torch.from_numpy(np.concatenate([counts.ravel(), totals])).to(dev)
        torch.distributed.all_reduce(t)
        t = t.cpu().numpy()
        counts, totals = t[: counts.size].reshape(counts.shape), t[counts.size:]
    direction_scale = []
    for i in range(len(edit_concepts)):
        ratios = np.array([want - (counts[i, j] / totals[i]) for j, want in zip(range(len(debias_concepts)), desired_ratios)])      # uce_sd_debias.py:29
        if max(ratios) < max_diff and abs(min(ratios)) < max_diff:
            ratios = 0 * ratios
        direction_scale.append(ratios)
    return np.array(direction_scale)


def UCE(pipe, clip, edit_concepts, debias_concepts, preserve_concepts, edit_scale, preserve_scale, lamb, save_dir, exp_name,
        max_diff, step_size, num_images_per_prompt, num_inference_steps, guidance_scale,
        max_iterations=30, desired_ratios=(0.5, 0.5), device="cuda:0", solver: EditSolver | None = None, verbose=True,
        generator=None):
    """``generator``: what ``get_ratios`` generates with.  None = the caller's ``pipe`` (the reference's behaviour, uce_sd_debias.py:22-26);
    "engine" (or UCE_DEBIAS_ENGINE=1) = an ``EngineGenerator`` built from ``pipe`` — the denoise loop of every generation round on the B200
    U-Net engine; or any object with the same face (``.unet.load_state_dict``, ``__call__(...).images``)."""
    if generator is None and os.environ.get("UCE_DEBIAS_ENGINE") == "1":
        generator = "engine"
    projections = select_projections(pipe.unet)
    names = [n for n, _ in projections]
    dev = torch.device(device)
    rows = embed_concepts(pipe, list(edit_concepts) + list(debias_concepts) + list(preserve_concepts), device)
    c_edit = torch.stack([rows[e] for e in edit_concepts]).to(dev)
    c_deb = torch.stack([rows[c] for c in debias_concepts]).to(dev)
    c_pres = [rows[p].to(dev) for p in preserve_concepts]
    C = torch.cat([c_edit] + ([torch.stack(c_pres)] if c_pres else []), 0)
    scales = [float(edit_scale)] * len(edit_concepts) + [float(preserve_scale)] * len(c_pres)
    w_old = [m.weight.detach().to(dev, torch.float32).contiguous().clone() for _, m in projections]
    K = w_old[0].shape[1]
    own = solver is None
    if own:
        solver = EditSolver(K, max(16, C.shape[0]), dev)

    pipe = pipe.to(torch.bfloat16)          # generation dtype (uce_sd_debias.py:90); the solve stays fp32
    own_gen = False
    if isinstance(generator, str):
        if generator != "engine":
            raise ValueError(f"generator must be None, 'engine' or a generator object, got {generator!r}")
        from .generate import EngineGenerator
        generator, own_gen = EngineGenerator(pipe, num_images_per_prompt, device=device), True
    gen = generator if generator is not None else pipe
    current = [w.clone() for w in w_old]    # weights the next generation round uses (:45-46)
    A = np.zeros((len(edit_concepts), len(debias_concepts)), dtype=np.float64)
    start = time.time()
    iterations = 0
    for iteration in range(max_iterations):
        direction_scale = get_ratios(pipe=gen, clip=clip, uce_module_names=names, uce_modules=current,
                                     edit_concepts=edit_concepts, debias_concepts=debias_concepts,
                                     desired_ratios=desired_ratios, max_diff=max_diff, step_size=step_size,
                                     num_images_per_prompt=num_images_per_prompt,
                                     num_inference_steps=num_inference_steps, guidance_scale=guidance_scale)
        if np.abs(direction_scale).max() == 0:
            if verbose:
                print("All concepts are debiased")
            break
        A += direction_scale
        G = c_edit + torch.from_numpy(A).to(dev, torch.float64).matmul(c_deb.to(torch.float64)).to(torch.float32)
        current = solver.edit(C, G, scales, len(edit_concepts), lamb, w_old)
        iterations += 1
    elapsed = time.time() - start
    state = {name + ".weight": w for name, w in zip(names, current)}
    if save_dir is not None and _rank_world()[0] == 0:          # every rank holds the same weights: one writer
        from .artifact import save_artifact       # byte-identical to safetensors.torch.save_file (tests/test_artifact.py)
        os.makedirs(save_dir, exist_ok=True)
        save_artifact(state, os.path.join(save_dir, exp_name + ".safetensors"))
    if own:
        solver.close()
    if own_gen:
        generator.close()
    if verbose:
        print(f"\n\nDebiased concepts using UCE\nModel edited in {elapsed} seconds\n")
    return state
