"""Per-sample Jacobians, API of ``preds/gradients.py:20-46``.

The reference runs K forward+backward passes under BackPACK ``BatchGrad`` and concatenates
``grad_batch``; here one forward and one K-column backward produce the layer factors
(a_l, g_l) and ``bnnp_jacobians`` writes Js[n, p, k] = g_l[n,k,i] * a_l[n,j] once.  The hot path
(``Laplace``) never calls this: it consumes the factors directly.
"""
from ._engine import net_for


def Jacobians(model, data):
    """Returns (Js [N,P,K] -- [N,P] when the model has one output -- and f as the model returns it:
    [N,K], or [N] for a model that flattens its single output, preds/models.py:165), detached."""
    net = net_for(model)
    Js, f = net.materialise_jacobians(data)
    if net.K > 1:
        return Js, f
    return Js.reshape(Js.shape[0], Js.shape[1]), (f.reshape(-1) if net.flat_output else f)
