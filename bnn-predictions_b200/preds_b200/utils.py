"""Classification metrics with the signatures of ``preds/utils.py:6-52`` (``acc``, ``macc``, ``nll_cls``, ``ece``):
one pass of ``bnnp_metric_cls`` over the mean predictive instead of a dozen torch reductions and
boolean-mask gathers per metric.  ``classification_metrics`` (additive) returns all three from a single launch.
"""
import torch

from . import _cabi as cabi
from ._cabi import lib, check, ptr, stream
from .likelihoods import BernoulliLh, CategoricalLh


def _reduce(probs, y, code, bins):
    cabi.require_device(probs)
    probs = cabi.f32(probs, probs.device)
    if code == cabi.LIK_CATEGORICAL:
        N, K = probs.shape
        yy = y.to(device=probs.device).long().contiguous()
    else:
        probs = probs.reshape(-1)
        N, K = probs.shape[0], 1
        yy = y.to(device=probs.device, dtype=torch.float32).reshape(-1).contiguous()
    if yy.shape[0] != N:
        raise ValueError('labels and predictions differ in length')
    bounds = torch.linspace(0, 1, bins + 1).to(probs.device)          # preds/utils.py:36
    out = torch.empty(3 + 3 * bins, dtype=torch.float64, device=probs.device)
    check(lib.bnnp_metric_cls(code, ptr(probs), ptr(yy), N, K, ptr(bounds), bins, ptr(out), stream()),
          'bnnp_metric_cls')
    return out.cpu(), N


def _ece_from(out, N, bins):
    cnt, conf, hit = out[3::3][:bins], out[4::3][:bins], out[5::3][:bins]
    nz = cnt > 0
    return float((torch.abs(conf[nz] / cnt[nz] - hit[nz] / cnt[nz]) * (cnt[nz] / N)).sum())


def _code(likelihood, default):
    if type(likelihood) is CategoricalLh:
        return cabi.LIK_CATEGORICAL
    if type(likelihood) is BernoulliLh:
        return cabi.LIK_BERNOULLI
    return default


def acc(g, y, likelihood=None):
    """preds/utils.py:6-11: multiclass accuracy for CategoricalLh, else (g >= 0.5) == y."""
    out, N = _reduce(g, y, _code(likelihood, cabi.LIK_BERNOULLI), 0)
    return float(out[1]) / N


def macc(g, y):
    """preds/utils.py:14-16."""
    out, N = _reduce(g, y, cabi.LIK_CATEGORICAL, 0)
    return float(out[1]) / N


def nll_cls(p, y, likelihood):
    """preds/utils.py:19-28: average negative log-likelihood of the labels under the predictive probabilities."""
    code = _code(likelihood, None)
    if code is None:
        raise ValueError('Only Bernoulli and Categorical likelihood.')
    out, N = _reduce(p, y, code, 0)
    return -float(out[0]) / N


def ece(probs, labels, likelihood=None, bins=10):
    """preds/utils.py:31-52: expected calibration error over ``bins`` equal-width confidence bins."""
    out, N = _reduce(probs, labels, _code(likelihood, cabi.LIK_CATEGORICAL), bins)
    return _ece_from(out, N, bins)


def classification_metrics(p, y, likelihood, bins=10):
    """{'nll', 'acc', 'ece'} from one kernel launch (additive)."""
    code = _code(likelihood, None)
    if code is None:
        raise ValueError('Only Bernoulli and Categorical likelihood.')
    out, N = _reduce(p, y, code, bins)
    return {'nll': -float(out[0]) / N, 'acc': float(out[1]) / N, 'ece': _ece_from(out, N, bins)}
