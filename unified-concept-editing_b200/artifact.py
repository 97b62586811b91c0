"""The UCE artifact on disk (include/uce_b200.h, csrc/artifact.cu): the safetensors file of the reference
(trainscripts/uce_sd_erase.py:85-88 writes it, evalscripts/generate-images-sd.py:17-19 reads it), without the `safetensors`
package on the path: a native writer that is byte-identical to ``safetensors.torch.save_file`` for the same dictionary, fed
straight from ONE pinned staging buffer (asynchronous device-to-host copies, a single stream wait), and a native reader."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _native as N


def save_artifact(state: Dict[str, torch.Tensor], path: str) -> None:
    """Write ``state`` (fp32, two-dimensional tensors: the edited attn2.to_k / to_v weights) as ``path``.
    Device tensors are gathered into one pinned host buffer with non-blocking copies."""
    names = sorted(state)
    for k in names:
        t = state[k]
        if t.dtype != torch.float32 or t.dim() != 2:
            raise ValueError(f"artifact tensors are fp32 [out_features, in_features] (uce_sd_erase.py:117), got {k}: {t.dtype} {tuple(t.shape)}")
    total = sum(state[k].numel() for k in names)
    on_gpu = any(state[k].is_cuda for k in names)
    stage = torch.empty(total, dtype=torch.float32, pin_memory=on_gpu and torch.cuda.is_available())
    views, off = [], 0
    for k in names:
        t = state[k].detach()
        v = stage[off: off + t.numel()].view(t.shape)
        v.copy_(t, non_blocking=True)
        views.append(v)
        off += t.numel()
    for dev in {state[k].device for k in names if state[k].is_cuda}:
        torch.cuda.synchronize(dev)
    n = len(names)
    c_names = (C.c_char_p * n)(*[k.encode() for k in names])
    c_data = (C.c_void_p * n)(*[v.data_ptr() for v in views])
    rows = (C.c_long * n)(*[v.shape[0] for v in views])
    cols = (C.c_long * n)(*[v.shape[1] for v in views])
    N.check(N.lib().uce_artifact_write_f32(path.encode(), n, c_names, c_data, rows, cols))


def load_artifact(path: str) -> Dict[str, torch.Tensor]:
    """Read every tensor of a safetensors file written by the reference or by save_artifact (fp32) into CPU tensors."""
    lib = N.lib()
    h = C.c_void_p()
    N.check(lib.uce_artifact_open(path.encode(), C.byref(h)))
    try:
        out = {}
        for i in range(lib.uce_artifact_count(h)):
            name, dtype, ndim, shape = C.c_char_p(), C.c_char_p(), C.c_int(), (C.c_long * 8)()
            N.check(lib.uce_artifact_entry(h, i, C.byref(name), C.byref(dtype), C.byref(ndim), shape))
            t = torch.empty([shape[d] for d in range(ndim.value)], dtype=torch.float32)
            N.check(lib.uce_artifact_read_f32(h, i, C.c_void_p(t.data_ptr()), t.numel()))
            out[name.value.decode()] = t
        return out
    finally:
        lib.uce_artifact_close(h)
