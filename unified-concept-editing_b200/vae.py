"""Host API of the B200 VAE decoder engine (include/sd_vae_b200.h): ``vae.decode(latents / scaling_factor)`` followed by the
``(x / 2 + 0.5).clamp(0, 1)`` -> uint8 conversion that ends the reference's ``pipe(...)`` call
(evalscripts/generate-images-sd.py:37-46; explicit in evalscripts/concept_algebra.py:126-135).  Weights are addressed by the
diffusers state-dict names of ``pipe.vae`` (``post_quant_conv.*``, ``decoder.*``; encoder / quant_conv keys are ignored).

Opt-in in round 1: ``generate_images`` uses it when UCE_VAE_ENGINE=1."""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as N
from .unet import _lib as _sd_lib
from .vae_spec import SD14_VAE, decoder_param_shapes


class SDVaeConfig(C.Structure):
    _fields_ = [("latent_channels", C.c_int), ("out_channels", C.c_int), ("n_levels", C.c_int), ("block_out_channels", C.c_int * 4),
                ("layers_per_block", C.c_int), ("norm_groups", C.c_int), ("scaling_factor", C.c_float)]


VAE_SIGNATURES = {
    "sd_vae_create": (C.c_int, [C.c_int, C.POINTER(SDVaeConfig), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "sd_vae_destroy": (C.c_int, [C.c_void_p]),
    "sd_vae_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_long), C.c_int]),
    "sd_vae_finalize": (C.c_int, [C.c_void_p]),
    "sd_vae_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sd_vae_launch_count": (C.c_int, [C.c_void_p]),
    "sd_vae_inventory": (C.c_int, [C.POINTER(SDVaeConfig), C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_long), C.POINTER(C.c_int)]),
    "sd_vae_read_tap": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]),
}
_bound = False


def _lib():
    global _bound
    L = _sd_lib()                   # binds sd_last_error as well
    if not _bound:
        for name, (res, args) in VAE_SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _bound = True
    return L


def _check(rc):
    if rc != 0:
        raise N.UCEError(rc, _lib().sd_last_error().decode(errors="replace"))


def _cfg_struct(cfg) -> SDVaeConfig:
    ch = list(cfg["block_out_channels"])
    return SDVaeConfig(cfg["latent_channels"], cfg["out_channels"], len(ch), (C.c_int * 4)(*(ch + [0] * 4)[:4]), cfg["layers_per_block"],
                       cfg["norm_groups"], float(cfg["scaling_factor"]))


def engine_inventory(cfg=SD14_VAE) -> dict:
    """{name: shape} of the parameters the native engine expects for ``cfg`` (host-only; no GPU needed)."""
    cs = _cfg_struct(cfg)
    n = _lib().sd_vae_inventory(C.byref(cs), -1, None, 0, None, None)
    if n < 0:
        _check(n)
    out = {}
    buf, shp, nd = C.create_string_buffer(256), (C.c_long * 4)(), C.c_int()
    for i in range(n):
        rc = _lib().sd_vae_inventory(C.byref(cs), i, buf, 256, shp, C.byref(nd))
        if rc < 0:
            _check(rc)
        out[buf.value.decode()] = tuple(shp[: nd.value])
    return out


class VAEDecoderEngine:
    """Decoder for ``batch`` latents of size h x w per call."""

    def __init__(self, cfg=SD14_VAE, batch=1, h=64, w=64, device="cuda:0"):
        dev = torch.device(device)
        if dev.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("the VAE decoder engine runs on CUDA only (there is no CPU path)")
        self.device = dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())
        self.cfg, self.batch, self.h, self.w = dict(cfg), batch, h, w
        up = 1 << (len(cfg["block_out_channels"]) - 1)
        self.H, self.W = h * up, w * up
        self._cs = _cfg_struct(cfg)
        hd = C.c_void_p()
        _check(_lib().sd_vae_create(self.device.index, C.byref(self._cs), batch, h, w, C.byref(hd)))
        self._h = hd
        self._shapes = decoder_param_shapes(cfg)

    def close(self):
        if getattr(self, "_h", None):
            _lib().sd_vae_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_state_dict(self, state, strict=True):
        """Upload the decoder parameters of a ``pipe.vae.state_dict()``; keys outside the decoder are ignored unless strict."""
        for name, t in state.items():
            if name not in self._shapes:
                if strict and not name.startswith(("encoder.", "quant_conv.")):
                    raise KeyError(name)
                continue
            a = t.detach().to("cpu", torch.float32).contiguous()
            shp = (C.c_long * a.dim())(*a.shape)
            _check(_lib().sd_vae_set_weight(self._h, name.encode(), C.c_void_p(a.data_ptr()), shp, a.dim()))
        missing = [k for k in self._shapes if k not in state]
        if missing:
            raise KeyError(f"missing decoder parameters: {missing[:3]} ...")

    def finalize(self):
        with torch.cuda.device(self.device):
            _check(_lib().sd_vae_finalize(self._h))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def decode(self, latents: torch.Tensor, want_image=False):
        """latents [batch,4,h,w] fp32 on the device (NOT yet divided by scaling_factor) -> uint8 [batch,H,W,3] on the device
        (and, with want_image, also the fp32 [batch,3,H,W] decoder output)."""
        assert latents.device == self.device and latents.dtype == torch.float32 and latents.is_contiguous()
        assert tuple(latents.shape) == (self.batch, 4, self.h, self.w)
        rgb = torch.empty((self.batch, self.H, self.W, 3), dtype=torch.uint8, device=self.device)
        img = torch.empty((self.batch, 3, self.H, self.W), dtype=torch.float32, device=self.device) if want_image else None
        with torch.cuda.device(self.device):
            _check(_lib().sd_vae_decode(self._h, C.c_void_p(latents.data_ptr()), C.c_void_p(img.data_ptr()) if img is not None else None,
                                        C.c_void_p(rgb.data_ptr()), self._stream()))
        return (rgb, img) if want_image else rgb

    def launch_count(self) -> int:
        return _lib().sd_vae_launch_count(self._h)

    def read_tap(self, name: str) -> torch.Tensor:
        """Intermediate (``mid``, ``up.i``) as fp32 NCHW on the host; recorded only if UCE_VAE_TAPS was set before finalize()."""
        cap = self.batch * max(self.cfg["block_out_channels"]) * self.H * self.W
        buf = torch.empty(cap, dtype=torch.float32)
        dims = (C.c_int * 4)()
        _check(_lib().sd_vae_read_tap(self._h, name.encode(), C.c_void_p(buf.data_ptr()), cap, dims))
        d = list(dims)
        return buf[: d[0] * d[1] * d[2] * d[3]].view(*d).clone()
