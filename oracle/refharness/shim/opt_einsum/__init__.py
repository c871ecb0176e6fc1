import torch


def contract(expr, *operands, backend=None, **kw):
    return torch.einsum(expr, *operands)
