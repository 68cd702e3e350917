"""Fused optimizers (reference: `tf.keras.optimizers.Adam` P1/02:201, `Adadelta|Adam` looked up
by name in the HPO objective P2/01:154-155)."""
from .fused import SGD, Adam, Adadelta, FlatOptimizer, get, OPTIMIZERS

__all__ = ["SGD", "Adam", "Adadelta", "FlatOptimizer", "get", "OPTIMIZERS"]
