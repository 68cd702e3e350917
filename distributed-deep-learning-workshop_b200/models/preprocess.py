"""ONE preprocessing definition for training and inference (fixes the reference's train/serve skew).

Reference `preprocess(content, label_idx)` (P1/02:119-126 and four more copies): decode JPEG -> bilinear resize to
224x224 -> MobileNetV2 `preprocess_input` (x/127.5 - 1).  Here: `decode_image` (bytes -> uint8 HWC, CPU, PIL) and
`resize_uint8` run in the loader workers; the float conversion x/127.5-1 runs on the GPU (fused into the engine's
first kernel, or `preprocess_tensor` for autograd models).
"""
from __future__ import annotations

import io
from typing import Tuple

import numpy as np
import torch

IMG_HEIGHT = 224
IMG_WIDTH = 224
IMG_CHANNELS = 3


def decode_image(content: bytes, size: Tuple[int, int] = (IMG_HEIGHT, IMG_WIDTH)) -> np.ndarray:
    """JPEG/PNG bytes (or raw uint8 H*W*3 bytes of exactly the target size) -> uint8 [H, W, 3], bilinear resize."""
    h, w = size
    if isinstance(content, (bytes, bytearray, memoryview)) and len(content) == h * w * 3:
        return np.frombuffer(content, dtype=np.uint8).reshape(h, w, 3)
    from PIL import Image

    img = Image.open(io.BytesIO(bytes(content))).convert("RGB")
    if img.size != (w, h):
        img = img.resize((w, h), Image.BILINEAR)
    return np.asarray(img, dtype=np.uint8)


def _arrow_of(values):
    """The Arrow (chunked) array behind a pandas Series / Arrow array, or None when the input is not Arrow-backed."""
    try:
        import pandas as pd
        import pyarrow as pa
    except Exception:
        return None
    if isinstance(values, (pa.Array, pa.ChunkedArray)):
        return values
    arr = getattr(values, "array", None)
    if arr is not None and hasattr(arr, "__arrow_array__") and isinstance(getattr(values, "dtype", None), pd.ArrowDtype):
        return arr.__arrow_array__()
    return None


def decode_batch(values, size: Tuple[int, int] = (IMG_HEIGHT, IMG_WIDTH), threads: int = 8) -> np.ndarray:
    """A column of image payloads -> uint8 [n, H, W, 3].

    Fast path: an Arrow-backed column whose payloads are all raw H*W*3 uint8 images is returned as ONE zero-copy view
    of the column's data buffer (which the scoring workers place in pinned host memory, so the view is also the source
    of the host->device copy).  Otherwise every row goes through `decode_image` (JPEG/PNG decode + bilinear resize, PIL
    releases the GIL) on a small thread pool."""
    h, w = size
    row = h * w * 3
    arr = _arrow_of(values)
    if arr is not None:
        import pyarrow as pa

        chunks = arr.chunks if isinstance(arr, pa.ChunkedArray) else [arr]
        views = []
        for c in chunks:
            if len(c) == 0:
                continue
            if not (pa.types.is_binary(c.type) or pa.types.is_large_binary(c.type)) or c.null_count:
                views = None
                break
            odt = np.int64 if pa.types.is_large_binary(c.type) else np.int32
            bufs = c.buffers()
            offs = np.frombuffer(bufs[1], dtype=odt)[c.offset:c.offset + len(c) + 1]
            if int(offs[-1] - offs[0]) != len(c) * row or (len(c) > 1 and not np.all(np.diff(offs) == row)):
                views = None
                break
            views.append(np.frombuffer(bufs[2], dtype=np.uint8, count=len(c) * row, offset=int(offs[0])).reshape(len(c), h, w, 3))
        if views is not None:
            if not views:
                return np.zeros((0, h, w, 3), np.uint8)
            return views[0] if len(views) == 1 else np.concatenate(views, 0)
        values = arr.to_pylist()
    vals = list(values)
    if not vals:
        return np.zeros((0, h, w, 3), np.uint8)
    out = np.empty((len(vals), h, w, 3), np.uint8)

    def one(i):
        out[i] = decode_image(vals[i], size)

    if threads > 1 and len(vals) >= 2 * threads:
        import concurrent.futures as cf

        with cf.ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, range(len(vals))))
    else:
        for i in range(len(vals)):
            one(i)
    return out


def preprocess(content: bytes, label_idx: int, size: Tuple[int, int] = (IMG_HEIGHT, IMG_WIDTH)):
    """Reference-signature helper: (bytes, label) -> (float32 image in [-1, 1] HWC, label)."""
    img = decode_image(content, size).astype(np.float32) / 127.5 - 1.0
    return img, label_idx


def preprocess_tensor(x: torch.Tensor) -> torch.Tensor:
    """uint8 [B, H, W, 3] -> float32 [B, 3, H, W] in [-1, 1] (channels_last memory) for autograd models."""
    if x.dtype == torch.uint8:
        x = x.float().mul_(1.0 / 127.5).sub_(1.0)
    if x.dim() == 4 and x.shape[-1] == 3:
        x = x.permute(0, 3, 1, 2)
    return x
