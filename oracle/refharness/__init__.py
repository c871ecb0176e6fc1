"""Harness that imports the UNMODIFIED reference ``preds`` package in the build container.

TEST INFRASTRUCTURE ONLY -- used by ``oracle/gen_golden.py`` and by the container-only tests
that pin ``oracle/port.py`` against the reference itself.  ``/root/reference`` does not exist
on the GPU box, so nothing under ``-m gpu``, ``smoke()`` or ``bench.py`` may call this.

What it does:
* puts stand-ins for the absent third-party leaves (``backpack``, ``deepobs``, ``gpytorch``,
  ``opt_einsum``) and the reference root on ``sys.path``;
* re-creates ``torch.symeig`` (removed from torch; the reference calls it at
  ``preds/kron.py:155,161``) on top of ``torch.linalg.eigh(UPLO='U')``;
* turns off ``torch.distributions`` argument validation (torch 1.5, which the reference
  pins, did not validate; see ``preds/predictives.py:70-76``).
"""
import os
import sys

REFERENCE_ROOT = os.environ.get('BNNP_REFERENCE_ROOT', '/root/reference')
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shim')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'preds'))


def _symeig(M, eigenvectors=False, upper=True):
    import torch
    L, W = torch.linalg.eigh(M, UPLO='U' if upper else 'L')
    return L, W


def load():
    """Import and return the reference ``preds`` package (modules loaded as attributes)."""
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    import torch
    for p in (REFERENCE_ROOT, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    torch.symeig = _symeig
    torch.distributions.Distribution.set_default_validate_args(False)
    import preds
    import preds.laplace, preds.gradients, preds.optimizers, preds.likelihoods  # noqa
    import preds.kron, preds.predictives, preds.models, preds.utils  # noqa
    return preds


class inject_normals:
    """Context manager: make ``torch.randn`` and ``MultivariateNormal.sample`` consume the
    given standard-normal tensors (in call order) so reference sampling is reproducible."""

    def __init__(self, draws):
        self.draws = list(draws)

    def __enter__(self):
        import torch
        import torch.distributions.multivariate_normal as mvn
        import torch.distributions.normal as nrm
        self._randn = torch.randn
        self._std = mvn._standard_normal
        self._nstd = nrm._standard_normal
        harness = self

        def take(shape):
            t = harness.draws.pop(0)
            assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape))
            return t.clone()

        def randn(*size, **kw):
            if len(size) == 1 and not isinstance(size[0], int):
                size = tuple(size[0])
            return take(size)

        def std(shape, dtype, device):
            return take(tuple(shape))

        def normal(mean, std_, *a, **k):
            # Normal.sample draws through torch.normal(loc.expand(shape), scale.expand(shape)) (preds/laplace.py:199)
            return mean + std_ * take(tuple(mean.shape))

        self._normal = torch.normal
        torch.randn = randn
        torch.normal = normal
        mvn._standard_normal = std
        nrm._standard_normal = std
        return self

    def __exit__(self, *a):
        import torch
        import torch.distributions.multivariate_normal as mvn
        import torch.distributions.normal as nrm
        torch.randn = self._randn
        torch.normal = self._normal
        mvn._standard_normal = self._std
        nrm._standard_normal = self._nstd
        return False
