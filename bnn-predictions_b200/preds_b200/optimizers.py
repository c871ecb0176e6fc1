"""``GGN`` bundle with the signature of ``preds/optimizers.py:8-25``."""
from .gradients import Jacobians


def GGN(model, likelihood, data, target=None, ret_f=False):
    Js, f = Jacobians(model, data)
    if target is not None:
        rs = likelihood.residual(target.to(f.device), f)
    Hess = likelihood.Hessian(f)
    m, p = Js.shape[:2]
    if Js.dim() == 2:
        Hess = Hess.reshape(m, 1, 1)
        if target is not None:
            rs = rs.reshape(m, 1)
        Js = Js.reshape(m, p, 1)
    if target is not None:
        return (Js, Hess, rs, f) if ret_f else (Js, Hess, rs)
    return Js, Hess, f
