"""The reference's own model (C14): MobileNetV2 base (frozen) + GlobalAveragePooling + Dropout + Dense(num_classes)
(P1/02:159-178; `build_model(dropout)` P2/01:92-108).  Pure torch.nn - it is the *parity* model (6,405 trainable
parameters => a ~25 KB all-reduce), not the performance target; it runs through the module backend
of the Trainer, on GPU or CPU.  No pretrained weights exist offline, so the base is random-init."""
from __future__ import annotations

import torch
from torch import nn


def _make_divisible(v, divisor=8):
    return max(divisor, int(v + divisor / 2) // divisor * divisor)


class ConvBNReLU6(nn.Sequential):
    def __init__(self, cin, cout, k=3, stride=1, groups=1):
        super().__init__(nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, groups=groups, bias=False),
                         nn.BatchNorm2d(cout), nn.ReLU6(inplace=True))


class InvertedResidual(nn.Module):
    def __init__(self, cin, cout, stride, expand):
        super().__init__()
        hidden = int(round(cin * expand))
        self.use_res = stride == 1 and cin == cout
        layers = []
        if expand != 1:
            layers.append(ConvBNReLU6(cin, hidden, 1))
        layers += [ConvBNReLU6(hidden, hidden, 3, stride, groups=hidden), nn.Conv2d(hidden, cout, 1, bias=False),
                   nn.BatchNorm2d(cout)]
        self.conv = nn.Sequential(*layers)

    def forward(self, x):
        return x + self.conv(x) if self.use_res else self.conv(x)


class MobileNetV2Base(nn.Module):
    """include_top=False feature extractor: [B,3,H,W] -> [B,1280,H/32,W/32]."""

    def __init__(self, width_mult: float = 1.0):
        super().__init__()
        cfg = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]
        cin = _make_divisible(32 * width_mult)
        self.out_channels = _make_divisible(1280 * max(1.0, width_mult))
        feats = [ConvBNReLU6(3, cin, 3, 2)]
        for t, c, n, s in cfg:
            cout = _make_divisible(c * width_mult)
            for i in range(n):
                feats.append(InvertedResidual(cin, cout, s if i == 0 else 1, t))
                cin = cout
        feats.append(ConvBNReLU6(cin, self.out_channels, 1))
        self.features = nn.Sequential(*feats)

    def forward(self, x):
        return self.features(x)


class FrozenBaseClassifier(nn.Module):
    """Sequential[base (frozen, inference-mode BN), GlobalAveragePooling2D, Dropout(p), Dense(num_classes)] - logits."""

    def __init__(self, base: nn.Module, feat_channels: int, num_classes: int, dropout: float = 0.5,
                 freeze_base: bool = True):
        super().__init__()
        self.base = base
        self.freeze_base = freeze_base
        if freeze_base:
            for p in self.base.parameters():
                p.requires_grad_(False)
        self.dropout = nn.Dropout(dropout)
        self.fc = nn.Linear(feat_channels, num_classes)

    def train(self, mode: bool = True):
        super().train(mode)
        if self.freeze_base:
            self.base.eval()  # frozen base => BatchNorm in inference mode (Keras `layer.trainable = False`)
        return self

    def forward(self, x):
        if self.freeze_base:
            with torch.no_grad():
                f = self.base(x)
        else:
            f = self.base(x)
        return self.fc(self.dropout(f.mean(dim=(2, 3))))
