"""TEST INFRASTRUCTURE — CPU restatement of the reference's UCE edit solver.

Two flavours, both tensor-level (text encoding is outside the hot path):

* ``*_port_f32``   — follows the reference's arithmetic order in fp32: per
  projection, start from ``lamb*W_old`` / ``lamb*I`` (uce_sd_erase.py:58-63), add one
  rank-1 pair per (edit, guide) concept (:66-71) and per preserve concept
  (:74-79), then ``mat1 @ inverse(mat2)`` (:82).  This is what the reference
  computes, rounding behaviour included; it is also the timed CPU baseline.
* ``*_exact_f64``  — the same algebra in fp64 through the shared-factor closed
  form ``W_new = W_old (lamb I + G^T S C)(lamb I + C^T S C)^-1`` (SURVEY.md §0
  finding 2) — the accuracy yard-stick.

Pinned against the real reference by tests/golden/*.npz (make_golden.py).
"""
from __future__ import annotations

import numpy as np
import torch


# --------------------------------------------------------------------------- erase
def erase_port_f32(weights, c_edit, c_guide, c_pres, erase_scale=1.0, preserve_scale=1.0, lamb=0.5):
    """fp32 port of uce_sd_erase.py:45-82 on tensors.

    weights: list of [d_l, K] fp32; c_edit/c_guide: [Ne, K]; c_pres: [Np, K] (may be empty).
    Returns list of W_new [d_l, K] fp32, on the device of the inputs (CPU in the tests; bench.py also times it with the tensors on
    cuda:0 — the reference's own execution path, torch library kernels — as context next to the CPU baseline).
    """
    out = []
    K = weights[0].shape[1]
    with torch.no_grad():
        for w_old in weights:
            w_old = w_old.to(torch.float32)
            # guide / preserve targets through the ORIGINAL weight (uce_sd_erase.py:45-53)
            v_guide = [w_old @ c_guide[i] for i in range(c_guide.shape[0])]
            v_pres = [w_old @ c_pres[i] for i in range(c_pres.shape[0])]
            m1 = lamb * w_old
            m2 = lamb * torch.eye(K, dtype=torch.float32, device=w_old.device)
            for i in range(c_edit.shape[0]):
                c = c_edit[i].reshape(K, 1)
                m1 += erase_scale * (v_guide[i].reshape(-1, 1) @ c.T)
                m2 += erase_scale * (c @ c.T)
            for i in range(c_pres.shape[0]):
                c = c_pres[i].reshape(K, 1)
                m1 += preserve_scale * (v_pres[i].reshape(-1, 1) @ c.T)
                m2 += preserve_scale * (c @ c.T)
            out.append(m1 @ torch.inverse(m2))
    return out


def _as64(x):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float64)


def shared_factor_exact_f64(c_rows, g_rows, scales, lamb):
    """M = (lamb I + G^T S C)(lamb I + C^T S C)^-1 in fp64 ([K,K])."""
    C, G, s = _as64(c_rows), _as64(g_rows), _as64(scales)
    K = C.shape[1]
    A = lamb * np.eye(K) + G.T @ (s[:, None] * C)
    B = lamb * np.eye(K) + C.T @ (s[:, None] * C)
    # M = A B^-1  <=>  B M^T = A^T  (B symmetric)
    return np.linalg.solve(B, A.T).T


def erase_exact_f64(weights, c_edit, c_guide, c_pres, erase_scale=1.0, preserve_scale=1.0, lamb=0.5):
    """Exact (fp64) result of the same edit; returns list of float64 ndarrays."""
    ce, cg, cp = _as64(c_edit), _as64(c_guide), _as64(c_pres)
    if cp.size == 0:
        cp = np.zeros((0, ce.shape[1]))
    C = np.concatenate([ce, cp], 0)
    G = np.concatenate([cg, cp], 0)
    s = np.concatenate([np.full(ce.shape[0], float(erase_scale)), np.full(cp.shape[0], float(preserve_scale))])
    M = shared_factor_exact_f64(C, G, s, lamb)
    return [_as64(w) @ M for w in weights]


# --------------------------------------------------------------------------- debias
def ratios_port(labels_per_concept, debias_concepts, desired_ratios, max_diff):
    """Port of get_ratios' arithmetic (uce_sd_debias.py:28-35).

    labels_per_concept: list (one per edit concept) of top-1 label lists.
    ratio = desired − observed fraction; the whole row is zeroed when
    max(row) < max_diff and |min(row)| < max_diff (dead-band, :31-32).
    """
    rows = []
    for labels in labels_per_concept:
        res = np.array(labels)
        r = np.array([d - (np.sum(res == c) / len(res)) for c, d in zip(debias_concepts, desired_ratios)])
        if max(r) < max_diff and abs(min(r)) < max_diff:
            r = 0 * r
        rows.append(r)
    return np.array(rows)


def debias_port_f32(weights, c_edit, c_debias, c_pres, direction_scales, edit_scale=1.0, preserve_scale=1.0, lamb=0.5):
    """fp32 port of the solve inside uce_sd_debias.py:95-141 for a scripted
    sequence of ``direction_scale`` matrices ([Ne,Nd] each, float64 like numpy).

    Faithful to the reference's quirks: targets start at W_old·c_edit (:87-88),
    are updated IN PLACE and therefore accumulate across iterations (:124-126),
    every solve restarts from W_old (:116), the loop stops at the first all-zero
    ``direction_scale`` (:110-112) leaving the previous iteration's weights.
    Returns list of W_new (or W_old if the first matrix is already zero).
    """
    K = weights[0].shape[1]
    result = [w.clone().to(torch.float32) for w in weights]
    with torch.no_grad():
        v_edit = [[(w @ c_edit[i]).to(torch.float32) for i in range(c_edit.shape[0])] for w in weights]
        v_deb = [[(w @ c_debias[j]).to(torch.float32) for j in range(c_debias.shape[0])] for w in weights]
        v_pres = [[(w @ c_pres[i]).to(torch.float32) for i in range(c_pres.shape[0])] for w in weights]
        for ds in direction_scales:
            ds = np.asarray(ds, dtype=np.float64)
            if np.abs(ds).max() == 0:
                break
            for l, w_old in enumerate(weights):
                m1 = lamb * w_old.to(torch.float32)
                m2 = lamb * torch.eye(K, dtype=torch.float32, device=w_old.device)
                for i in range(c_edit.shape[0]):
                    v = v_edit[l][i]
                    for j in range(c_debias.shape[0]):
                        v += ds[i][j] * v_deb[l][j]          # in place, cumulative
                    c = c_edit[i].reshape(K, 1)
                    m1 += edit_scale * (v.reshape(-1, 1) @ c.T)
                    m2 += edit_scale * (c @ c.T)
                for i in range(c_pres.shape[0]):
                    c = c_pres[i].reshape(K, 1)
                    m1 += preserve_scale * (v_pres[l][i].reshape(-1, 1) @ c.T)
                    m2 += preserve_scale * (c @ c.T)
                result[l] = m1 @ torch.inverse(m2)
    return result


def debias_exact_f64(weights, c_edit, c_debias, c_pres, direction_scales, edit_scale=1.0, preserve_scale=1.0, lamb=0.5):
    """fp64 closed form: G_e = C_e + A C_d with A = cumulative sum of the non-zero-prefix
    of direction_scales."""
    ce, cd, cp = _as64(c_edit), _as64(c_debias), _as64(c_pres)
    if cp.size == 0:
        cp = np.zeros((0, ce.shape[1]))
    A = np.zeros((ce.shape[0], cd.shape[0]))
    any_step = False
    for ds in direction_scales:
        ds = np.asarray(ds, dtype=np.float64)
        if np.abs(ds).max() == 0:
            break
        A += ds
        any_step = True
    if not any_step:
        return [_as64(w) for w in weights]
    C = np.concatenate([ce, cp], 0)
    G = np.concatenate([ce + A @ cd, cp], 0)
    s = np.concatenate([np.full(ce.shape[0], float(edit_scale)), np.full(cp.shape[0], float(preserve_scale))])
    M = shared_factor_exact_f64(C, G, s, lamb)
    return [_as64(w) @ M for w in weights]


# --------------------------------------------------------------------------- metrics
def rel_fro(a, b) -> float:
    """‖a−b‖_F / ‖b‖_F in fp64."""
    a, b = _as64(a), _as64(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
