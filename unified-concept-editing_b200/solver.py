"""Tensor-level host API of the UCE edit solver (thin layer over the C ABI, include/uce_b200.h).

Sits where the reference does its arithmetic (trainscripts/uce_sd_erase.py:45-82,
trainscripts/uce_sd_debias.py:114-140): concept rows and projection weights in, edited
projection weights out.  torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch

from . import _native as N

APPLY_AUTO, APPLY_SIMT, APPLY_TCGEN05_FUSED, APPLY_TCGEN05_TWO_GEMM, APPLY_TCGEN05_KSPLIT = 0, 1, 4, 5, 7


def _ptr(t: torch.Tensor) -> int:
    return t.data_ptr()


class EditSolver:
    """One workspace on one GPU for text dimension ``K`` and at most ``max_rows`` concept rows."""

    def __init__(self, K: int, max_rows: int, device="cuda:0"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("the UCE solver runs on CUDA only (there is no CPU path)")
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device visible: the UCE hot path cannot run")
        self.device = dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())
        self.K, self.max_rows = int(K), int(max_rows)
        h = C.c_void_p()
        N.check(N.lib().uce_ws_create(self.device.index, self.K, self.max_rows, C.byref(h)))
        self._h = h
        self._keep = []   # keeps pointer arrays alive until the stream has consumed them

    def close(self):
        if getattr(self, "_h", None):
            N.lib().uce_ws_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ configuration
    def set_apply_impl(self, impl: int) -> int:
        return N.lib().uce_ws_set_apply_impl(self._h, int(impl))

    def set_factor_impl(self, impl: int) -> int:
        """0 auto (low-latency single-CTA factor when it applies), 1 general blocked path."""
        return N.lib().uce_ws_set_factor_impl(self._h, int(impl))

    def set_debug(self, on: bool) -> int:
        return N.lib().uce_ws_set_debug(self._h, int(bool(on)))

    def set_profile(self, on: bool) -> int:
        return N.lib().uce_ws_set_profile(self._h, int(bool(on)))

    def timings(self):
        """(factor_ms, apply_stage1_ms, apply_stage2_ms) of the last factor/apply (profile mode)."""
        ms = (C.c_float * 3)()
        N.check(N.lib().uce_ws_timings(self._h, ms))
        return tuple(ms)

    # ------------------------------------------------------------------ phase 1
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check_rows(self, C_rows, G_rows, scales, n_edit):
        n = C_rows.shape[0]
        if C_rows.dim() != 2 or C_rows.shape[1] != self.K:
            raise ValueError(f"C must be [n,{self.K}], got {tuple(C_rows.shape)}")
        if not (0 <= n_edit <= n):
            raise ValueError("n_edit out of range")
        if n_edit and (G_rows is None or tuple(G_rows.shape) != (n_edit, self.K)):
            raise ValueError(f"G must be [{n_edit},{self.K}]")
        if len(scales) != n:
            raise ValueError("one scale per concept row")
        if n > self.max_rows:
            raise ValueError(f"{n} concept rows exceed the workspace capacity {self.max_rows}")

    def factor(self, C_rows: torch.Tensor, G_rows: torch.Tensor | None, scales: Sequence[float], n_edit: int, lamb: float):
        """Shared factor (uce_factor_dev_f32).  C_rows [n,K]: edit rows first, then preserve rows."""
        self._check_rows(C_rows, G_rows, scales, n_edit)
        Cd = C_rows.to(self.device, torch.float32).contiguous()
        Gd = G_rows.to(self.device, torch.float32).contiguous() if n_edit else None
        sc = (C.c_float * len(scales))(*[float(s) for s in scales])
        with torch.cuda.device(self.device):
            N.check(N.lib().uce_factor_dev_f32(self._h, C.c_void_p(_ptr(Cd)), C.c_void_p(_ptr(Gd)) if n_edit else None, sc,
                                               Cd.shape[0], int(n_edit), float(lamb), self._stream()))
        self._keep = [Cd, Gd]

    # ------------------------------------------------------------------ phase 2
    def apply(self, W_old: Sequence[torch.Tensor], W_new: Sequence[torch.Tensor] | None = None):
        """Batched apply (uce_apply_dev_f32).  Returns the list of edited weights (new tensors unless W_new given)."""
        L = len(W_old)
        for w in W_old:
            if w.device != self.device or w.dtype != torch.float32 or not w.is_contiguous() or w.dim() != 2 or w.shape[1] != self.K:
                raise ValueError("W_old must be contiguous fp32 [d,K] tensors on the solver's device")
        if W_new is None:
            W_new = [torch.empty_like(w) for w in W_old]
        po = (C.c_void_p * L)(*[_ptr(w) for w in W_old])
        pn = (C.c_void_p * L)(*[_ptr(w) for w in W_new])
        dd = (C.c_int * L)(*[int(w.shape[0]) for w in W_old])
        with torch.cuda.device(self.device):
            N.check(N.lib().uce_apply_dev_f32(self._h, po, pn, dd, L, self._stream()))
        return list(W_new)

    def check(self):
        """Synchronise the current stream and raise on a deferred numerical failure."""
        with torch.cuda.device(self.device):
            N.check(N.lib().uce_ws_check(self._h, self._stream()))

    def edit(self, C_rows, G_rows, scales, n_edit, lamb, W_old, W_new=None, check=True):
        """Factor + apply in one call (uce_edit_dev_f32): the library overlaps the part of the apply that does not need the factor."""
        self._check_rows(C_rows, G_rows, scales, n_edit)
        Cd = C_rows.to(self.device, torch.float32).contiguous()
        Gd = G_rows.to(self.device, torch.float32).contiguous() if n_edit else None
        sc = (C.c_float * len(scales))(*[float(s) for s in scales])
        L = len(W_old)
        for w in W_old:
            if w.device != self.device or w.dtype != torch.float32 or not w.is_contiguous() or w.dim() != 2 or w.shape[1] != self.K:
                raise ValueError("W_old must be contiguous fp32 [d,K] tensors on the solver's device")
        if W_new is None:
            W_new = [torch.empty_like(w) for w in W_old]
        po = (C.c_void_p * L)(*[_ptr(w) for w in W_old])
        pn = (C.c_void_p * L)(*[_ptr(w) for w in W_new])
        dd = (C.c_int * L)(*[int(w.shape[0]) for w in W_old])
        with torch.cuda.device(self.device):
            N.check(N.lib().uce_edit_dev_f32(self._h, C.c_void_p(_ptr(Cd)), C.c_void_p(_ptr(Gd)) if n_edit else None, sc, Cd.shape[0],
                                             int(n_edit), float(lamb), po, pn, dd, L, self._stream()))
        self._keep = [Cd, Gd]
        if check:
            self.check()
        return list(W_new)

    def edit_two_calls(self, C_rows, G_rows, scales, n_edit, lamb, W_old, W_new=None, check=True):
        """The same edit through uce_factor_dev_f32 + uce_apply_dev_f32 (everything serial on the current stream)."""
        self.factor(C_rows, G_rows, scales, n_edit, lamb)
        out = self.apply(W_old, W_new)
        if check:
            self.check()
        return out

    # ------------------------------------------------------------------ host-buffer path
    @staticmethod
    def host_arena(dims: Sequence[int], K: int, pin: bool = True):
        """One (pinned) host buffer holding every projection back to back, and its per-projection views ``[d_l, K]``.  A caller that keeps
        W_old / W_new in such arenas lets ``edit_host`` move each pipeline group as ONE copy per direction instead of one per
        projection (the 32 + 32 small copies of SD-1.4 cost the PCIe link ~0.5 ms more than flat ones, profiles/r01_e2e_floor.txt)."""
        total = sum(int(d) * int(K) for d in dims)
        buf = torch.empty(total, dtype=torch.float32)
        if pin:
            buf = buf.pin_memory()
        views, off = [], 0
        for d in dims:
            n = int(d) * int(K)
            views.append(buf[off:off + n].view(int(d), int(K)))
            off += n
        return buf, views

    def edit_host(self, C_rows, G_rows, scales, n_edit, lamb, W_old: Sequence[torch.Tensor], W_new: Sequence[torch.Tensor]):
        """Whole edit on HOST tensors (uce_edit_host_f32): copies are inside the call."""
        self._check_rows(C_rows, G_rows, scales, n_edit)
        key = (tuple(w.data_ptr() for w in W_old), tuple(w.data_ptr() for w in W_new), tuple(int(w.shape[0]) for w in W_old))
        cached = getattr(self, "_host_args", None)
        if cached is None or cached[0] != key:                      # pointer tables are rebuilt only when the caller's buffers change
            for t in list(W_old) + list(W_new):
                if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous():
                    raise ValueError("edit_host takes contiguous fp32 CPU tensors")
            L = len(W_old)
            cached = (key, (C.c_void_p * L)(*key[0]), (C.c_void_p * L)(*key[1]), (C.c_int * L)(*key[2]), L)
            self._host_args = cached
        for t in [C_rows] + ([G_rows] if n_edit else []):
            if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("edit_host takes contiguous fp32 CPU tensors")
        _, po, pn, dd, L = cached
        sc = (C.c_float * len(scales))(*[float(s) for s in scales])
        N.check(N.lib().uce_edit_host_f32(self._h, C.c_void_p(_ptr(C_rows)), C.c_void_p(_ptr(G_rows)) if n_edit else None, sc,
                                          C_rows.shape[0], int(n_edit), float(lamb), po, pn, dd, L))
        return list(W_new)

    # ------------------------------------------------------------------ introspection
    def info(self) -> dict:
        v = [C.c_int() for _ in range(6)]
        N.check(N.lib().uce_ws_info(self._h, *[C.byref(x) for x in v]))
        keys = ["mode", "rank", "dense", "sys_n", "launches_factor", "launches_apply"]
        d = {k: x.value for k, x in zip(keys, v)}
        d["mode_name"] = {0: "none", 1: "dual", 2: "primal"}[d["mode"]]
        return d

    def debug_read(self, which: int) -> torch.Tensor:
        i = self.info()
        n, K, r = i["sys_n"], self.K, i["rank"]
        if which in (0, 1):
            out = torch.empty(n, n, dtype=torch.float64)
        elif which in (2, 3):
            out = torch.empty(r, K, dtype=torch.float32)
        elif which == 4:
            out = torch.empty(K, K, dtype=torch.float32)
        else:
            raise ValueError(which)
        N.check(N.lib().uce_ws_debug_read(self._h, which, C.c_void_p(out.data_ptr()), out.numel() * out.element_size()))
        return out
