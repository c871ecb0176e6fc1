"""Minimal stand-ins for the gpytorch lazy tensors ``preds/gps.py`` touches (gpytorch is absent)."""
import torch


class DiagLazyTensor:
    """``DiagLazyTensor(d).evaluate()`` = ``diag_embed(d)`` (preds/gps.py:108,440)."""

    def __init__(self, diag):
        self._diag = diag

    def evaluate(self):
        return torch.diag_embed(self._diag)

    def diag(self):
        return self._diag

    def to(self, *a, **k):
        return DiagLazyTensor(self._diag.to(*a, **k))


class BlockDiagLazyTensor:  # noqa
    pass


class KroneckerProductLazyTensor:  # noqa
    pass


class NonLazyTensor:  # noqa
    pass


def __getattr__(name):
    return type(name, (), {})
