"""Drop-in for the reference's erase/moderate/replace edit, ``UCE()`` of
trainscripts/uce_sd_erase.py:12-91, with the arithmetic on the B200 solver.

Same inputs (a duck-typed diffusers pipeline and concept lists), same side effect
(``<save_dir>/<exp_name>.safetensors`` holding only the edited attn2.to_k/to_v weights, fp32,
keys ``<module path>.weight`` — :85-88).  The edited weights are also returned.
"""
from __future__ import annotations

import os
import time

import torch

from .concepts import embed_concepts, select_projections
from .solver import EditSolver


def build_rows(rows: dict, edit_concepts, guide_concepts, preserve_concepts, erase_scale, preserve_scale):
    """Concept lists -> (C [n,K], G [n_edit,K], scales, n_edit) in the solver's row order.
    Every LISTED concept is a row: a concept listed twice is summed twice, exactly like the
    reference's loops (uce_sd_erase.py:66-79)."""
    c_edit = [rows[e] for e in edit_concepts]
    c_guide = [rows[g] for g in guide_concepts]
    c_pres = [rows[p] for p in preserve_concepts]
    C = torch.stack(c_edit + c_pres)
    G = torch.stack(c_guide) if c_guide else None
    scales = [float(erase_scale)] * len(c_edit) + [float(preserve_scale)] * len(c_pres)
    return C, G, scales, len(c_edit)


def UCE(pipe, edit_concepts, guide_concepts, preserve_concepts, erase_scale, preserve_scale, lamb, save_dir, exp_name,
        device="cuda:0", solver: EditSolver | None = None, verbose=True):
    start = time.time()
    projections = select_projections(pipe.unet)
    names = [n for n, _ in projections]
    dev = torch.device(device)

    # text rows (one per distinct prompt; uce_sd_erase.py:25-42)
    rows = embed_concepts(pipe, list(edit_concepts) + list(guide_concepts) + list(preserve_concepts), device)
    if len(guide_concepts) != len(edit_concepts):
        raise Exception("edit and guide concept lists must pair up")
    C, G, scales, n_edit = build_rows(rows, edit_concepts, guide_concepts, preserve_concepts, erase_scale, preserve_scale)

    # original projection weights, fp32 on the solve device (uce_sd_erase.py:21,58,117)
    w_old = [m.weight.detach().to(dev, torch.float32).contiguous() for _, m in projections]
    K = w_old[0].shape[1]
    own = solver is None
    if own:
        solver = EditSolver(K, max(16, C.shape[0]), dev)

    dist_on = torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
    if dist_on:
        from .sharding import GatherPlan, default_chunks, sharded_edit
        rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
        dims = [w.shape[0] for w in w_old]
        # this rank's projections are written by the apply kernels straight into the gather buffer; large edits go in chunks whose
        # all-gathers overlap the next chunk's kernels; a rank without projections (more ranks than projections) still takes part
        plan = GatherPlan(dims, K, world, rank, dev, chunks=default_chunks(dims, K, world))
        w_new = sharded_edit(solver, plan, C.to(dev), G.to(dev) if G is not None else None, scales, n_edit, lamb, dict(enumerate(w_old)))
        is_writer = rank == 0
    else:
        w_new = solver.edit(C.to(dev), G.to(dev) if G is not None else None, scales, n_edit, lamb, w_old)
        is_writer = True

    state = {name + ".weight": w for name, w in zip(names, w_new)}
    if is_writer and save_dir is not None:
        from .artifact import save_artifact       # byte-identical to safetensors.torch.save_file (tests/test_artifact.py)
        os.makedirs(save_dir, exist_ok=True)
        save_artifact(state, os.path.join(save_dir, exp_name + ".safetensors"))
    if own:
        solver.close()
    if verbose:
        print(f"\n\nErased concepts using UCE\nModel edited in {time.time() - start} seconds\n")
    return state
