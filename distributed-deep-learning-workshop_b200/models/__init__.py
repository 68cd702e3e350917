"""Model zoo + the `build_model` / `preprocess` helpers of the reference (C13, C14).

    build_model(img_height, img_width, img_channels, num_classes)      (P1/02:159, P1/03:159)
    build_model(dropout=0.5)                                            (P2/01:92, P2/02:116)
    build_model()                                                       (P2/03:125)

`arch='resnet50'` (TARGET, default on a B200): the hand-scheduled sm_100a engine, fully trainable, bf16.
`arch='mobilenetv2'` (REF parity): frozen MobileNetV2 base + GAP + Dropout + Dense head - on a GPU the native
`MobileNetV2Engine` (depthwise / stem kernels + tcgen05 pointwise convs with folded-BN epilogues), on CPU (or with
`arch='mobilenetv2_torch'`, or `freeze_base=False`) the torch.nn module.
`arch='resnet50_torch'`: torchvision-architecture ResNet-50 as a plain nn.Module (CPU plumbing runs).
"""
from __future__ import annotations

from typing import Optional

import torch

from .preprocess import (IMG_CHANNELS, IMG_HEIGHT, IMG_WIDTH, decode_batch, decode_image, preprocess, preprocess_tensor)

CLASSES = ["daisy", "dandelion", "roses", "sunflowers", "tulips"]  # reference P2/03:62


def build_model(img_height: int = IMG_HEIGHT, img_width: int = IMG_WIDTH, img_channels: int = IMG_CHANNELS,
                num_classes: int = 5, dropout: float = 0.5, arch: Optional[str] = None, batch_size: int = 32,
                freeze_base: bool = True, device=None, seed: int = 0, **engine_kwargs):
    if isinstance(img_height, float) and img_width == IMG_WIDTH:  # build_model(dropout) positional form
        dropout, img_height = img_height, IMG_HEIGHT
    if img_channels != 3:
        raise ValueError("only 3-channel images are supported")
    if arch is None:
        arch = "resnet50" if (torch.cuda.is_available() and img_height == img_width and img_height % 32 == 0) else "mobilenetv2"
    if arch == "mobilenetv2_head":   # SURVEY.md 7.1's spelling of the reference's model (frozen MobileNetV2 base + trainable head)
        arch = "mobilenetv2"
    if arch == "resnet50":
        from .resnet_engine import ResNet50Engine

        return ResNet50Engine(batch=batch_size, num_classes=num_classes, device=device, image_size=img_height,
                              dropout=dropout, seed=seed, **engine_kwargs)
    if arch == "mobilenetv2" and freeze_base and torch.cuda.is_available() and img_height == img_width and img_height % 32 == 0 \
            and (device is None or torch.device(device).type == "cuda"):
        # the reference's own model on native kernels (frozen base = inference-mode layers; trainable Dense head)
        from .mobilenet_engine import MobileNetV2Engine

        return MobileNetV2Engine(batch=batch_size, num_classes=num_classes, device=device, image_size=img_height,
                                 dropout=dropout, seed=seed)
    if arch in ("mobilenetv2", "mobilenetv2_torch"):
        from .mobilenet import FrozenBaseClassifier, MobileNetV2Base

        g = torch.random.fork_rng(devices=[])
        with g:
            torch.manual_seed(seed)
            base = MobileNetV2Base()
            return FrozenBaseClassifier(base, base.out_channels, num_classes, dropout, freeze_base)
    if arch == "resnet50_torch":
        import torchvision

        with torch.random.fork_rng(devices=[]):
            torch.manual_seed(seed)
            return torchvision.models.resnet50(weights=None, num_classes=num_classes)
    raise ValueError(f"unknown arch {arch!r}")


__all__ = ["build_model", "preprocess", "preprocess_tensor", "decode_image", "decode_batch", "CLASSES", "IMG_HEIGHT", "IMG_WIDTH",
           "IMG_CHANNELS"]
