"""Host API of the B200 U-Net denoise engine (include/sd_unet_b200.h): the eps-prediction network the
reference reaches through ``pipe(...)`` (evalscripts/generate-images-sd.py:37-42), classifier-free guidance and
the scheduler update.  Weights are addressed by their diffusers state-dict names."""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as N
from .unet_spec import SD14, param_shapes


class SDUnetConfig(C.Structure):
    _fields_ = [("in_channels", C.c_int), ("out_channels", C.c_int), ("n_levels", C.c_int), ("block_out_channels", C.c_int * 4),
                ("layers_per_block", C.c_int), ("down_has_attn", C.c_int * 4), ("up_has_attn", C.c_int * 4),
                ("cross_attention_dim", C.c_int), ("context_len", C.c_int), ("heads", C.c_int), ("norm_groups", C.c_int),
                ("temb_dim", C.c_int)]


SD_SIGNATURES = {
    "sd_last_error": (C.c_char_p, []),
    "sd_unet_create": (C.c_int, [C.c_int, C.POINTER(SDUnetConfig), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "sd_unet_destroy": (C.c_int, [C.c_void_p]),
    "sd_unet_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_long), C.c_int]),
    "sd_unet_finalize": (C.c_int, [C.c_void_p]),
    "sd_unet_set_context": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "sd_unet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sd_unet_set_timestep": (C.c_int, [C.c_void_p, C.c_float]),
    "sd_cfg_step": (C.c_int, [C.c_void_p, C.c_long, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float),
                              C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sd_unet_inventory": (C.c_int, [C.POINTER(SDUnetConfig), C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_long), C.POINTER(C.c_int)]),
    "sd_unet_launch_count": (C.c_int, [C.c_void_p]),
    "sd_unet_context_launch_count": (C.c_int, [C.c_void_p]),
    "sd_unet_read_tap": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]),
}
_bound = False


def _lib():
    global _bound
    L = N.lib()
    if not _bound:
        for name, (res, args) in SD_SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _bound = True
    return L


def _check(rc):
    if rc != 0:
        raise N.UCEError(rc, _lib().sd_last_error().decode(errors="replace"))


def _cfg_struct(cfg, context_len=77) -> SDUnetConfig:
    ch = list(cfg["block_out_channels"])
    pad = lambda xs: (list(xs) + [0] * 4)[:4]
    return SDUnetConfig(cfg["in_channels"], cfg["out_channels"], len(ch), (C.c_int * 4)(*pad(ch)), cfg["layers_per_block"],
                        (C.c_int * 4)(*pad(int(x) for x in cfg["down_has_attn"])), (C.c_int * 4)(*pad(int(x) for x in cfg["up_has_attn"])),
                        cfg["cross_attention_dim"], context_len, cfg["heads"], cfg["norm_groups"], cfg["temb_dim"])


def engine_inventory(cfg=SD14, context_len=77) -> dict:
    """{name: shape} of the parameters the native engine expects for ``cfg`` (host-only; no GPU needed)."""
    cs = _cfg_struct(cfg, context_len)
    n = _lib().sd_unet_inventory(C.byref(cs), -1, None, 0, None, None)
    if n < 0:
        _check(n)
    out = {}
    buf, shp, nd = C.create_string_buffer(256), (C.c_long * 4)(), C.c_int()
    for i in range(n):
        rc = _lib().sd_unet_inventory(C.byref(cs), i, buf, 256, shp, C.byref(nd))
        if rc < 0:
            _check(rc)
        out[buf.value.decode()] = tuple(shp[: nd.value])
    return out


class UNetEngine:
    """U-Net for ``batch`` samples per call (2 x images under classifier-free guidance) at latent size H x W."""

    def __init__(self, cfg=SD14, batch=2, H=64, W=64, device="cuda:0", context_len=77):
        dev = torch.device(device)
        if dev.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("the U-Net engine runs on CUDA only (there is no CPU path)")
        self.device = dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())
        self.cfg, self.batch, self.H, self.W, self.context_len = dict(cfg), batch, H, W, context_len
        self._cs = _cfg_struct(cfg, context_len)
        h = C.c_void_p()
        _check(_lib().sd_unet_create(self.device.index, C.byref(self._cs), batch, H, W, C.byref(h)))
        self._h = h
        self._shapes = param_shapes(cfg)

    def close(self):
        if getattr(self, "_h", None):
            _lib().sd_unet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_state_dict(self, state, strict=True):
        """Upload parameters by diffusers name.  strict=False accepts a subset (e.g. only the UCE-edited
        attn2.to_k/to_v tensors, like generate-images-sd.py:19) and ignores unknown keys."""
        for name, t in state.items():
            if name not in self._shapes:
                if strict:
                    raise KeyError(name)
                continue
            a = t.detach().to("cpu", torch.float32).contiguous()
            shp = (C.c_long * a.dim())(*a.shape)
            _check(_lib().sd_unet_set_weight(self._h, name.encode(), C.c_void_p(a.data_ptr()), shp, a.dim()))
        if strict:
            missing = [k for k in self._shapes if k not in state]
            if missing:
                raise KeyError(f"missing parameters: {missing[:3]} ...")

    def finalize(self):
        with torch.cuda.device(self.device):
            _check(_lib().sd_unet_finalize(self._h))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check_ctx(self, ctx: torch.Tensor):
        assert ctx.device == self.device and ctx.dtype == torch.float32 and ctx.is_contiguous()
        assert tuple(ctx.shape) == (self.batch, self.context_len, self.cfg["cross_attention_dim"])

    def set_context(self, ctx: torch.Tensor):
        """Text context [batch,context_len,ctx_dim] fp32 of the following steps: its cross-attention K / V^T projections are
        computed once here instead of at every denoise step (the prompt embedding is constant over a row,
        evalscripts/generate-images-sd.py:37-42)."""
        self._check_ctx(ctx)
        with torch.cuda.device(self.device):
            _check(_lib().sd_unet_set_context(self._h, C.c_void_p(ctx.data_ptr()), self._stream()))

    def forward(self, x: torch.Tensor, t: float, ctx: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
        """eps [batch,4,H,W] fp32 = UNet(x [batch,4,H,W] fp32, t, ctx [batch,context_len,ctx_dim] fp32) — device tensors.
        ctx=None reuses the context given to set_context()."""
        assert x.device == self.device and x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape) == (self.batch, 4, self.H, self.W)
        if ctx is not None:
            self._check_ctx(ctx)
        if out is None:
            out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _check(_lib().sd_unet_forward(self._h, C.c_void_p(x.data_ptr()), float(t), C.c_void_p(ctx.data_ptr()) if ctx is not None else None,
                                          C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def set_timestep(self, t: float):
        """Timestep read by the next replay of a CUDA graph captured around forward()."""
        _check(_lib().sd_unet_set_timestep(self._h, float(t)))

    def launch_count(self) -> int:
        """Kernels one denoise step enqueues (context projections excluded: see context_launch_count)."""
        return _lib().sd_unet_launch_count(self._h)

    def context_launch_count(self) -> int:
        return _lib().sd_unet_context_launch_count(self._h)

    def read_tap(self, name: str) -> torch.Tensor:
        cap = self.batch * max(self.cfg["block_out_channels"]) * 2 * self.H * self.W
        buf = torch.empty(cap, dtype=torch.float32)
        dims = (C.c_int * 4)()
        _check(_lib().sd_unet_read_tap(self._h, name.encode(), C.c_void_p(buf.data_ptr()), cap, dims))
        d = list(dims)
        n = d[0] * d[1] * d[2] * d[3]
        return buf[:n].view(*d).clone() if d[2] > 1 or d[3] > 1 else buf[:n].view(d[0], d[1]).clone()


def cfg_step(eps2, gs, x_in, x_out, coeffs, cx, ce, hist=(), eps_out=None):
    """Fused guidance + scheduler update (sd_cfg_step) on fp32 device tensors; eps2 = [uncond | text]."""
    n = x_in.numel()
    c = (C.c_float * 4)(*[float(v) for v in (list(coeffs) + [0, 0, 0, 0])[:4]])
    hp = [C.c_void_p(h.data_ptr()) for h in hist] + [None] * (3 - len(hist))
    _check(_lib().sd_cfg_step(C.c_void_p(eps2.data_ptr()), n, float(gs), C.c_void_p(eps_out.data_ptr()) if eps_out is not None else None,
                              hp[0], hp[1], hp[2], c, float(cx), float(ce), C.c_void_p(x_in.data_ptr()), C.c_void_p(x_out.data_ptr()),
                              C.c_void_p(torch.cuda.current_stream(x_in.device).cuda_stream)))
    return x_out
