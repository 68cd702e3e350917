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
