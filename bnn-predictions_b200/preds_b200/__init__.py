"""preds_b200 -- B200-native drop-in for the Laplace-GGN path of ``preds`` (aleximmer/BNN-predictions).

    from preds_b200.laplace import Laplace
    from preds_b200.likelihoods import CategoricalLh
    from preds_b200.models import MLPS

mirrors ``preds.laplace`` / ``preds.likelihoods`` / ``preds.models`` / ``preds.kron`` /
``preds.gradients`` / ``preds.optimizers`` / ``preds.predictives`` / ``preds.gps.gp_quantities`` for that path.  Importing the
package loads ``lib/libbnnp_b200.so`` (built by ``build.py``); there is no CPU fallback.
"""
from . import _cabi  # noqa: F401  (fails loudly if the CUDA library is missing)
from .laplace import Laplace, FunctionaLaplace  # noqa: F401
from .gps import gp_quantities  # noqa: F401
from .kron import Kron, symeig  # noqa: F401
from .gradients import Jacobians  # noqa: F401
from .optimizers import GGN, LaplaceGGN, get_diagonal_ggn  # noqa: F401
from .likelihoods import CategoricalLh, GaussianLh, BernoulliLh  # noqa: F401
from .models import MLPS, MLP, SiMLP, CIFAR10Net, CIFAR100Net  # noqa: F401
