"""Zero-shot image classifier of the debias loop on the B200 (ctypes binding of include/clip_vision_b200.h).

``ClipZeroShotEngine`` is called like the object the reference builds with ``transformers.pipeline("zero-shot-image-classification",
"openai/clip-vit-base-patch32")`` (trainscripts/uce_sd_debias.py:245-250) and uses at :27 — ``clip(images, candidate_labels=[...])`` ->
per image the labels sorted by softmax score, best first — with preprocessing, both CLIP towers and the scoring as CUDA kernels.  The
tokenizer stays the library's (host-side string work).  No CPU fallback: the library must be built and a CUDA device present."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N
from .clip_text import ClipTextEngine

SIGNATURES = {
    "clipv_last_error": (C.c_char_p, []),
    "clipv_create": (C.c_int, [C.c_int] * 10 + [C.POINTER(C.c_void_p)]),
    "clipv_destroy": (C.c_int, [C.c_void_p]),
    "clipv_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]),
    "clipv_finalize": (C.c_int, [C.c_void_p]),
    "clipv_preprocess_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "clipv_image_features": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "clipv_logits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "clipv_launch_count": (C.c_int, [C.c_void_p]),
}
_bound = False

HYPOTHESIS_TEMPLATE = "This is a photo of {}."                       # the pipeline's default (pipelines/zero_shot_image_classification.py)
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _lib():
    global _bound
    L = N.lib()
    if not _bound:
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _bound = True
    return L


def _check(rc):
    if rc != 0:
        raise RuntimeError(f"clip vision engine error {rc}: {_lib().clipv_last_error().decode(errors='replace')}")


class ClipVisionEngine:
    """CLIP image tower + scoring on one GPU.  ``state``: transformers CLIPModel state dict (``vision_model.*``, the two projections,
    ``logit_scale``)."""

    def __init__(self, state: dict, heads: int, device="cuda:0", max_batch: int = 16):
        dev = torch.device(device)
        if dev.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("the CLIP vision engine runs on CUDA only (there is no CPU path)")
        self.device = dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())
        pw = state["vision_model.embeddings.patch_embedding.weight"]
        pos = state["vision_model.embeddings.position_embedding.weight"]
        self.width, self.patch = int(pw.shape[0]), int(pw.shape[-1])
        grid = int(round((pos.shape[0] - 1) ** 0.5))
        self.image_size = grid * self.patch
        self.layers = 1 + max(int(k.split(".")[3]) for k in state if k.startswith("vision_model.encoder.layers."))
        self.ffn = int(state["vision_model.encoder.layers.0.mlp.fc1.weight"].shape[0])
        self.proj_dim = int(state["visual_projection.weight"].shape[0])
        self.text_width = int(state["text_projection.weight"].shape[1])
        self.heads, self.max_batch = int(heads), int(max_batch)
        h = C.c_void_p()
        _check(_lib().clipv_create(self.device.index, self.image_size, self.patch, self.width, self.heads, self.layers, self.ffn, self.proj_dim,
                                   self.text_width, self.max_batch, C.byref(h)))
        self._h = h
        for name, w in state.items():
            if not (name.startswith("vision_model.") or name in ("visual_projection.weight", "text_projection.weight", "logit_scale")) \
                    or name.endswith("position_ids"):
                continue
            t = w.detach().to("cpu", torch.float32).reshape(-1).contiguous()
            _check(_lib().clipv_set_weight(self._h, name.encode(), C.c_void_p(t.data_ptr()), t.numel()))
        _check(_lib().clipv_finalize(self._h))

    def close(self):
        if getattr(self, "_h", None):
            _lib().clipv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def preprocess(self, images_u8: torch.Tensor, mean=CLIP_MEAN, std=CLIP_STD) -> torch.Tensor:
        """CLIPImageProcessor for square images: [B, H, H, 3] uint8 (device) -> pixel_values [B, 3, S, S] fp32."""
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
            raise ValueError("images must be uint8 [B, H, W, 3]")
        B, H, W, _ = images_u8.shape
        if H != W:
            raise NotImplementedError("the CLIP preprocessing kernel handles square images (the generator's output); got %dx%d" % (H, W))
        img = images_u8.to(self.device).contiguous()
        out = torch.empty((B, 3, self.image_size, self.image_size), dtype=torch.float32, device=self.device)
        m = (C.c_float * 3)(*mean)
        s = (C.c_float * 3)(*std)
        with torch.cuda.device(self.device):
            _check(_lib().clipv_preprocess_u8(self._h, C.c_void_p(img.data_ptr()), B, H, m, s, C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def image_features(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """[B, proj_dim] fp32 (not normalised) — CLIPModel.get_image_features."""
        pv = pixel_values.to(self.device, torch.float32).contiguous()
        B = pv.shape[0]
        if tuple(pv.shape[1:]) != (3, self.image_size, self.image_size):
            raise ValueError(f"pixel_values must be [B, 3, {self.image_size}, {self.image_size}]")
        out = torch.empty((B, self.proj_dim), dtype=torch.float32, device=self.device)
        for b0 in range(0, B, self.max_batch):
            n = min(self.max_batch, B - b0)
            with torch.cuda.device(self.device):
                _check(_lib().clipv_image_features(self._h, C.c_void_p(pv[b0:b0 + n].data_ptr()), n, C.c_void_p(out[b0:b0 + n].data_ptr()), self._stream()))
        return out

    def logits(self, image_features: torch.Tensor, text_rows: torch.Tensor) -> torch.Tensor:
        """logits_per_image [B, N] = exp(logit_scale) * cos(image, text_projection(text_rows))."""
        im = image_features.to(self.device, torch.float32).contiguous()
        tx = text_rows.to(self.device, torch.float32).contiguous()
        if tx.shape[1] != self.text_width or im.shape[1] != self.proj_dim:
            raise ValueError("feature widths do not match the engine")
        out = torch.empty((im.shape[0], tx.shape[0]), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _check(_lib().clipv_logits(self._h, C.c_void_p(im.data_ptr()), im.shape[0], C.c_void_p(tx.data_ptr()), tx.shape[0],
                                       C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def launch_count(self) -> int:
        return _lib().clipv_launch_count(self._h)


def eos_index(input_ids: torch.Tensor, eos_token_id) -> torch.Tensor:
    """Where CLIPTextTransformer pools: the original checkpoints (config eos id 2) take ``argmax(input_ids)`` — the end-of-text token has
    the largest id —, newer configs the first ``eos_token_id`` (transformers/models/clip/modeling_clip.py, CLIPTextTransformer.forward)."""
    ids = input_ids.to("cpu", torch.int64)
    if eos_token_id is None or int(eos_token_id) == 2:
        return ids.argmax(dim=-1)
    return (ids == int(eos_token_id)).int().argmax(dim=-1)


def _to_u8_batch(images, device) -> torch.Tensor:
    """list of PIL images / uint8 arrays [H, W, 3] / one uint8 tensor [B, H, W, 3] -> uint8 tensor [B, H, W, 3] on the device."""
    if isinstance(images, torch.Tensor):
        t = images if images.dim() == 4 else images[None]
        return t.to(device)
    if isinstance(images, np.ndarray) and images.ndim == 4:
        return torch.from_numpy(np.ascontiguousarray(images)).to(device)
    if not isinstance(images, (list, tuple)):
        images = [images]
    arrs = []
    for im in images:
        if isinstance(im, torch.Tensor):
            a = im.detach().cpu().numpy()
        elif isinstance(im, np.ndarray):
            a = im
        else:                                                          # PIL.Image
            a = np.asarray(im.convert("RGB"))
        if a.dtype != np.uint8 or a.ndim != 3 or a.shape[-1] != 3:
            raise ValueError("images must be uint8 RGB [H, W, 3]")
        arrs.append(a)
    return torch.from_numpy(np.stack(arrs)).to(device)


class ClipZeroShotEngine:
    """``clip(images, candidate_labels=[...])`` -> ``[[{"score", "label"}, ...] per image]``, labels sorted best first."""

    accepts_device_images = True

    def __init__(self, state: dict, vision_heads: int, text_heads: int, tokenizer, eos_token_id=2, device="cuda:0", max_batch: int = 16,
                 image_mean=CLIP_MEAN, image_std=CLIP_STD, hypothesis_template: str = HYPOTHESIS_TEMPLATE):
        self.vision = ClipVisionEngine(state, vision_heads, device=device, max_batch=max_batch)
        self.text = ClipTextEngine({k: v for k, v in state.items() if k.startswith("text_model.")}, text_heads, device=device, max_batch=64)
        self.device = self.vision.device
        self.tokenizer, self.eos_token_id = tokenizer, eos_token_id
        self.mean, self.std, self.template = tuple(image_mean), tuple(image_std), hypothesis_template
        self._text_cache = {}

    @classmethod
    def from_model(cls, model, tokenizer, image_processor=None, device="cuda:0", **kw):
        """From a transformers ``CLIPModel`` (+ tokenizer, image processor)."""
        cfg = model.config
        if image_processor is not None:
            kw.setdefault("image_mean", tuple(image_processor.image_mean))
            kw.setdefault("image_std", tuple(image_processor.image_std))
        return cls(model.state_dict(), cfg.vision_config.num_attention_heads, cfg.text_config.num_attention_heads, tokenizer,
                   eos_token_id=getattr(cfg.text_config, "eos_token_id", 2), device=device, **kw)

    @classmethod
    def from_pipeline(cls, hf_pipeline, device="cuda:0", **kw):
        """From the object the reference builds (uce_sd_debias.py:245-250)."""
        return cls.from_model(hf_pipeline.model, hf_pipeline.tokenizer, getattr(hf_pipeline, "image_processor", None), device=device, **kw)

    def text_rows(self, candidate_labels, hypothesis_template=None) -> torch.Tensor:
        """End-of-text rows [N, text_width] of the label prompts; cached per label set (the debias loop scores against the same labels
        every iteration)."""
        tpl = self.template if hypothesis_template is None else hypothesis_template
        key = (tpl, tuple(candidate_labels))
        if key not in self._text_cache:
            enc = self.tokenizer([tpl.format(c) for c in candidate_labels], padding=True, return_tensors="pt")
            self._text_cache[key] = self.rows_at_eos(enc["input_ids"])
        return self._text_cache[key]

    def rows_at_eos(self, input_ids: torch.Tensor) -> torch.Tensor:
        return self.text.rows_at(input_ids, eos_index(input_ids, self.eos_token_id))

    def logits_per_image(self, images, input_ids=None, candidate_labels=None) -> torch.Tensor:
        u8 = _to_u8_batch(images, self.device)
        feats = self.vision.image_features(self.vision.preprocess(u8, self.mean, self.std))
        rows = self.rows_at_eos(input_ids) if input_ids is not None else self.text_rows(candidate_labels)
        return self.vision.logits(feats, rows)

    def __call__(self, images, candidate_labels, hypothesis_template=None, **_):
        single = not isinstance(images, (list, tuple, torch.Tensor)) and not (isinstance(images, np.ndarray) and images.ndim == 4)
        labels = list(candidate_labels)
        u8 = _to_u8_batch(images, self.device)
        feats = self.vision.image_features(self.vision.preprocess(u8, self.mean, self.std))
        logits = self.vision.logits(feats, self.text_rows(labels, hypothesis_template))
        probs = torch.softmax(logits, dim=-1).cpu()
        out = []
        for row in probs:
            order = sorted(range(len(labels)), key=lambda j: -float(row[j]))
            out.append([{"score": float(row[j]), "label": labels[j]} for j in order])
        return out[0] if single else out
