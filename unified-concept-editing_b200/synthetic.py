"""Seeded synthetic workloads of the edit solve: SD-1.4 / SDXL-shaped
attn2.to_k/to_v weights and CLIP-like concept rows (norm ≈ 28; optional shared component so
that prompts correlate, cos ≈ 0.64).  Used by bench.py and the full-size property tests."""
from __future__ import annotations

import torch

SD14_DIMS = [320] * 4 + [640] * 4 + [1280] * 4 + [1280] * 6 + [640] * 6 + [320] * 6 + [1280] * 2     # 32 projections, K=768
SDXL_DIMS = [640] * 8 + [1280] * 40 + [1280] * 60 + [640] * 12 + [1280] * 20                            # 140 projections, K=2048

WORKLOADS = {
    # BASELINE.json configs[0..3] (config 3 = debias: 10 professions, 2 debias concepts)
    "cfg1": dict(dims=[320, 320], K=768, n_edit=2, n_pres=3),
    "cfg2": dict(dims=SD14_DIMS, K=768, n_edit=50, n_pres=100),
    "cfg3": dict(dims=SD14_DIMS, K=768, n_edit=10, n_pres=0),
    "cfg4": dict(dims=SDXL_DIMS, K=2048, n_edit=1000, n_pres=0),
}


def concept_rows(n: int, K: int, seed: int, correlated: bool = True, norm: float = 28.0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, K, generator=g)
    if correlated:
        u = torch.randn(K, generator=torch.Generator().manual_seed(seed + 7919))
        x = 0.8 * u + 0.6 * x
    return (x * (norm / x.norm(dim=-1, keepdim=True))).to(torch.float32)


def weights(dims, K: int, seed: int, scale: float = 0.03, pin_memory: bool = False):
    g = torch.Generator().manual_seed(seed + 104729)
    out = []
    for d in dims:
        w = torch.randn(d, K, generator=g) * scale
        out.append(w.pin_memory() if pin_memory else w)
    return out


def problem(name: str, seed: int = 0, correlated: bool = True, pin_memory: bool = False):
    """dict(C [n,K] edit rows first, G [n_edit,K], scales, n_edit, lamb, W list) — all CPU fp32."""
    w = WORKLOADS[name]
    K, ne, npz = w["K"], w["n_edit"], w["n_pres"]
    rows = concept_rows(ne + npz + 1, K, seed, correlated)
    C = rows[: ne + npz].contiguous()
    G = rows[ne + npz:].expand(ne, K).contiguous()          # one shared guide row ("art")
    return dict(name=name, K=K, C=C, G=G, scales=[1.0] * (ne + npz), n_edit=ne, lamb=0.5,
                W=weights(w["dims"], K, seed, pin_memory=pin_memory), dims=list(w["dims"]))


def unet_random_state(cfg, seed: int = 0):
    """Seeded synthetic U-Net parameters (diffusers names) with fan-in scaling so activations stay O(1)
    (norm gains ~1, small biases) — there is no network access for real checkpoints."""
    import math
    from .unet_spec import param_shapes
    g = torch.Generator().manual_seed(seed)
    state = {}
    for name, shp in param_shapes(cfg).items():
        if name.endswith(".weight") and len(shp) == 1:
            w = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif name.endswith(".bias"):
            w = 0.02 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            w = torch.randn(shp, generator=g) / math.sqrt(fan_in)
        state[name] = w
    return state
