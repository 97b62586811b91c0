"""Host-side scheduler logic of the denoise loop, reduced to what the fused CUDA update (sd_cfg_step) consumes:
per U-Net call a timestep, the multistep weights over the eps history, and the two scalars of
``x_prev = cx * x_in + ce * e``.  Follows the schedulers the reference's pipeline resolves to (SURVEY.md
Appendix B): PNDM with skip_prk_steps (PLMS, the CompVis/stable-diffusion-v1-4 default that
evalscripts/generate-images-sd.py:13-15 therefore uses) and DDIM (eta = 0)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def alphas_cumprod() -> np.ndarray:
    import torch
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2     # scaled_linear, fp32 torch ops like the reference stack
    return torch.cumprod(1.0 - betas, dim=0).numpy().astype(np.float64)


@dataclass
class StepPlan:
    t: int                 # timestep fed to the U-Net
    coeffs: tuple          # weights over (current eps, history[-1], history[-2], history[-3])
    cx: float
    ce: float
    use_saved_sample: bool  # PLMS second call: x_in is the sample saved at the first call
    save_sample: bool       # PLMS first call: remember x
    append_eps: bool        # whether the guided eps joins the history


class PNDMPlan:
    """steps_offset = 1, set_alpha_to_one = False, skip_prk_steps = True: S + 1 U-Net calls for S steps."""

    def __init__(self, steps: int):
        self.ac = alphas_cumprod()
        self.r = 1000 // steps
        ts = (np.arange(steps) * self.r).round().astype(np.int64) + 1
        self.timesteps = np.concatenate([ts[:-1], ts[-2:-1], ts[-1:]])[::-1].copy()

    def _scalars(self, t, prev):
        a = self.ac[t]
        ap = self.ac[prev] if prev >= 0 else self.ac[0]
        b, bp = 1 - a, 1 - ap
        return float((ap / a) ** 0.5), float(-(ap - a) / (a * bp ** 0.5 + (a * b * ap) ** 0.5))

    def plans(self):
        out, n_hist, counter = [], 0, 0
        for t in self.timesteps:
            t = int(t)
            t_unet = t
            prev = t - self.r
            append = counter != 1
            if append:
                n_hist = min(n_hist, 3) + 1
            else:
                prev, t = t, t + self.r
            if n_hist == 1 and counter == 0:
                coeffs, use_saved, save = (1.0, 0.0, 0.0, 0.0), False, True
            elif n_hist == 1 and counter == 1:
                coeffs, use_saved, save = (0.5, 0.5, 0.0, 0.0), True, False      # (eps + ets[-1]) / 2, eps not appended
            elif n_hist == 2:
                coeffs, use_saved, save = (1.5, -0.5, 0.0, 0.0), False, False
            elif n_hist == 3:
                coeffs, use_saved, save = (23 / 12, -16 / 12, 5 / 12, 0.0), False, False
            else:
                coeffs, use_saved, save = (55 / 24, -59 / 24, 37 / 24, -9 / 24), False, False
            cx, ce = self._scalars(t, prev)
            out.append(StepPlan(t_unet, coeffs, cx, ce, use_saved, save, append))
            counter += 1
        return out


class DDIMPlan:
    def __init__(self, steps: int):
        self.ac = alphas_cumprod()
        self.r = 1000 // steps
        self.timesteps = ((np.arange(steps) * self.r).round()[::-1].copy().astype(np.int64)) + 1

    def plans(self):
        out = []
        for t in self.timesteps:
            t = int(t)
            prev = t - self.r
            a = self.ac[t]
            ap = self.ac[prev] if prev >= 0 else self.ac[0]
            cx = float((ap / a) ** 0.5)
            ce = float((1 - ap) ** 0.5 - (ap * (1 - a) / a) ** 0.5)
            out.append(StepPlan(t, (1.0, 0.0, 0.0, 0.0), cx, ce, False, False, False))
        return out


def make_plan(name: str, steps: int):
    if name == "pndm":
        return PNDMPlan(steps).plans()
    if name == "ddim":
        return DDIMPlan(steps).plans()
    raise ValueError(name)
