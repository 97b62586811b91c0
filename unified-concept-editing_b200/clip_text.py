"""Batched CLIP text encoder on the B200 (ctypes binding of include/clip_text_b200.h).

The reference encodes its concepts one prompt per forward (trainscripts/uce_sd_erase.py:26-42); this engine runs all distinct prompts
of an edit through ONE fp32 forward and returns the kept rows (``attention_mask.sum() - 2``).  There is no CPU fallback: the library
must be built and a CUDA device present."""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as N

SIGNATURES = {
    "clipt_last_error": (C.c_char_p, []),
    "clipt_create": (C.c_int, [C.c_int] * 8 + [C.POINTER(C.c_void_p)]),
    "clipt_destroy": (C.c_int, [C.c_void_p]),
    "clipt_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]),
    "clipt_finalize": (C.c_int, [C.c_void_p]),
    "clipt_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "clipt_concept_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "clipt_launch_count": (C.c_int, [C.c_void_p]),
}
_bound = False


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
        raise RuntimeError(f"clip text engine error {rc}: {_lib().clipt_last_error().decode(errors='replace')}")


class ClipTextEngine:
    """CLIPTextModel forward on one GPU.  ``state`` is the transformers state dict (names ``text_model.*``)."""

    def __init__(self, state: dict, heads: int, device="cuda:0", max_batch: int = 128):
        dev = torch.device(device)
        if dev.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("the CLIP text engine runs on CUDA only (there is no CPU path)")
        self.device = dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())
        tok = state["text_model.embeddings.token_embedding.weight"]
        pos = state["text_model.embeddings.position_embedding.weight"]
        layers = 1 + max(int(k.split(".")[3]) for k in state if k.startswith("text_model.encoder.layers."))
        ffn = state["text_model.encoder.layers.0.mlp.fc1.weight"].shape[0]
        self.vocab, self.width, self.heads, self.layers, self.max_pos, self.max_batch = tok.shape[0], tok.shape[1], int(heads), layers, pos.shape[0], int(max_batch)
        h = C.c_void_p()
        _check(_lib().clipt_create(self.device.index, self.vocab, self.width, self.heads, layers, ffn, self.max_pos, self.max_batch, C.byref(h)))
        self._h = h
        for name, w in state.items():
            if not name.startswith("text_model.") or name.endswith("position_ids"):
                continue
            t = w.detach().to("cpu", torch.float32).contiguous()
            _check(_lib().clipt_set_weight(self._h, name.encode(), C.c_void_p(t.data_ptr()), t.numel()))
        _check(_lib().clipt_finalize(self._h))

    def close(self):
        if getattr(self, "_h", None):
            _lib().clipt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def encode(self, input_ids: torch.Tensor) -> torch.Tensor:
        """last_hidden_state [B, T, width] fp32 on the device for token ids [B, T] (any integer tensor, host or device)."""
        ids = input_ids.detach().to("cpu", torch.int32).contiguous()
        B, T = ids.shape
        out = torch.empty((B, T, self.width), dtype=torch.float32, device=self.device)
        for b0 in range(0, B, self.max_batch):
            n = min(self.max_batch, B - b0)
            chunk = ids[b0:b0 + n].contiguous()
            with torch.cuda.device(self.device):
                _check(_lib().clipt_encode(self._h, C.c_void_p(chunk.data_ptr()), n, T, C.c_void_p(out[b0:b0 + n].data_ptr()), self._stream()))
        return out

    def rows_at(self, input_ids: torch.Tensor, row_index: torch.Tensor) -> torch.Tensor:
        """[B, width]: ``last_hidden_state[b, row_index[b], :]`` of every prompt."""
        ids = input_ids.detach().to("cpu", torch.int32).contiguous()
        idx = row_index.detach().to("cpu", torch.int32).contiguous()
        B, T = ids.shape
        out = torch.empty((B, self.width), dtype=torch.float32, device=self.device)
        for b0 in range(0, B, self.max_batch):
            n = min(self.max_batch, B - b0)
            chunk, ichunk = ids[b0:b0 + n].contiguous(), idx[b0:b0 + n].contiguous()
            with torch.cuda.device(self.device):
                _check(_lib().clipt_concept_rows(self._h, C.c_void_p(chunk.data_ptr()), C.c_void_p(ichunk.data_ptr()), n, T,
                                                 C.c_void_p(out[b0:b0 + n].data_ptr()), self._stream()))
        return out

    def concept_rows(self, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        """[B, width]: row ``attention_mask.sum() - 2`` of every prompt (trainscripts/uce_sd_erase.py:34-42)."""
        return self.rows_at(input_ids, attention_mask.detach().to("cpu").sum(dim=1) - 2)

    def launch_count(self) -> int:
        return _lib().clipt_launch_count(self._h)
