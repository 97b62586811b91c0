"""PNG output of the generation driver (evalscripts/generate-images-sd.py:45-46 saves every image with PIL) through the native
parallel-deflate writer of libuce_b200 (csrc/png.cu)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


def save_png(path: str, image, level: int = 6, threads: int = 0) -> None:
    """``image``: a PIL image or an array / tensor [H, W, 3] uint8 (RGB).  Lossless."""
    if hasattr(image, "convert") and hasattr(image, "size"):            # PIL.Image
        image = np.asarray(image.convert("RGB"))
    elif hasattr(image, "detach"):                                       # torch tensor
        image = image.detach().cpu().numpy()
    a = np.ascontiguousarray(image)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
        raise ValueError(f"save_png takes [H, W, 3] uint8 RGB, got {a.dtype} {a.shape}")
    N.check(N.lib().uce_png_write_rgb8(str(path).encode(), a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1], int(level), int(threads)))
