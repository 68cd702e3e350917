"""Keras-style loss objects for `Trainer.compile(loss=...)`.

The reference compiles every model with `tf.keras.losses.SparseCategoricalCrossentropy(from_logits=True)` (P1/02:203,
P1/03:223,346, P2/01:165, P2/02:226, P2/03:301); the models here output logits and the fused softmax-cross-entropy kernel
(csrc/elementwise.cu `softmax_ce`) / `torch.nn.functional.cross_entropy` consume logits, so that is the one form offered."""
from __future__ import annotations


class SparseCategoricalCrossentropy:
    name = "sparse_categorical_crossentropy"

    def __init__(self, from_logits: bool = False, name: str = "sparse_categorical_crossentropy", **_ignored):
        if not from_logits:
            raise ValueError("the models output logits: use SparseCategoricalCrossentropy(from_logits=True) "
                             "(what the reference workflow does)")
        self.from_logits = True
        self.name = name

    def __repr__(self) -> str:
        return "SparseCategoricalCrossentropy(from_logits=True)"


__all__ = ["SparseCategoricalCrossentropy"]
