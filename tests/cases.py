"""Golden-case registry: rebuilds the model / likelihood / data of each ``tests/golden/*.npz``
fixture (generated from the unmodified reference by ``oracle/gen_golden.py``)."""
import os

import numpy as np
import torch
import torch.nn as nn

from oracle import port

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

FULL_CASES = ['mlp_cls', 'mlp_relu3', 'mlp_reg', 'mlp_flat', 'mlp_bern', 'cnn_tiny']
MLP_CASES = ['mlp_cls', 'mlp_relu3', 'mlp_reg', 'mlp_flat', 'mlp_bern']

_MLPS = {
    'mlp_cls': dict(input_size=2, hidden_sizes=[10, 10], output_size=2, activation='tanh'),
    'mlp_relu3': dict(input_size=5, hidden_sizes=[12], output_size=3, activation='relu'),
    'mlp_reg': dict(input_size=3, hidden_sizes=[8], output_size=1, activation='tanh'),
    'mlp_flat': dict(input_size=16, hidden_sizes=[8, 6], output_size=4, activation='tanh',
                     flatten=True, bias=False),
    'mlp_bern': dict(input_size=4, hidden_sizes=[7], output_size=1, activation='tanh'),
}


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False))


def make_lh(desc, lib=port):
    desc = str(desc)
    if desc == 'categorical':
        return lib.CategoricalLh()
    if desc == 'bernoulli':
        return lib.BernoulliLh()
    if desc.startswith('gaussian:'):
        return lib.GaussianLh(float(desc.split(':')[1]))
    raise ValueError(desc)


def make_port_model(name):
    if name in _MLPS:
        return port.make_mlps(**_MLPS[name])
    if name == 'cnn_tiny':
        net = nn.Sequential(port.SamePad2d(3, 1), nn.Conv2d(2, 4, 3), nn.ReLU(),
                            port.SamePad2d(3, 2), nn.MaxPool2d(3, 2),
                            nn.Conv2d(4, 5, 3), nn.Tanh(), nn.Flatten(), nn.Linear(45, 3))
        net.output_size = 3
        return net
    raise ValueError(name)


def make_product_model(name, models):
    """Same architectures through the product package's own model zoo."""
    if name in _MLPS:
        return models.MLPS(**_MLPS[name])
    if name == 'cnn_tiny':
        net = nn.Sequential()
        net.add_module('conv1', models.tfconv2d(2, 4, 3, tf_padding_type='same'))
        net.add_module('relu1', nn.ReLU())
        net.add_module('pool1', models.tfmaxpool2d(3, stride=2, tf_padding_type='same'))
        net.add_module('conv2', models.tfconv2d(4, 5, 3))
        net.add_module('tanh2', nn.Tanh())
        net.add_module('flatten', nn.Flatten())
        net.add_module('dense', nn.Linear(45, 3))
        net.output_size = 3
        return net
    raise ValueError(name)


def case_data(g):
    nb = int(g['n_batches'])
    batches = [(torch.from_numpy(g['X%d' % i]), torch.from_numpy(g['y%d' % i])) for i in range(nb)]
    return batches, torch.from_numpy(g['Xte'])


def kron_eps_from(g):
    eps, i = [], 0
    while 'eps_kron_%d' % i in g:
        eps.append(torch.from_numpy(g['eps_kron_%d' % i]))
        i += 1
    return eps


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ---- LaplaceGGN path (experiments/classification.py) ---------------------------
GGN_CASES = {
    'ggn_simlp_cls': dict(input_size=6, output_size=3, n_layers=2, n_units=8, activation='tanh'),
    'ggn_simlp_bern': dict(input_size=4, output_size=1, n_layers=1, n_units=7, activation='tanh'),
}


def ggn_prior(g, P):
    """(prior precision as the constructor takes it, prior mean) of a LaplaceGGN golden case."""
    pp = torch.from_numpy(g['prior_prec'])
    pm = torch.from_numpy(g['prior_mu'])
    return (float(pp) if pp.dim() == 0 else pp), (float(pm) if pm.dim() == 0 else pm)


# ---- function-space inference (preds/gps.py:664-927, preds/laplace.py:138-207) --------
GP_CASES = ['gp_mlp', 'gp_cnn']


def make_gp_port_model(name):
    if name == 'gp_mlp':
        return port.make_mlps(5, [8, 6], 3, 'tanh')
    net = nn.Sequential(port.SamePad2d(3, 1), nn.Conv2d(2, 4, 3), nn.ReLU(), port.SamePad2d(3, 2), nn.MaxPool2d(3, 2),
                        nn.Conv2d(4, 5, 3), nn.Tanh(), nn.Flatten(), nn.Linear(45, 4))
    net.output_size = 4
    return net


def make_gp_product_model(name, models):
    if name == 'gp_mlp':
        return models.MLPS(5, [8, 6], 3, 'tanh')
    net = nn.Sequential()
    net.add_module('conv1', models.tfconv2d(2, 4, 3, tf_padding_type='same'))
    net.add_module('relu1', nn.ReLU())
    net.add_module('pool1', models.tfmaxpool2d(3, stride=2, tf_padding_type='same'))
    net.add_module('conv2', models.tfconv2d(4, 5, 3))
    net.add_module('tanh2', nn.Tanh())
    net.add_module('flatten', nn.Flatten())
    net.add_module('dense', nn.Linear(45, 4))
    net.output_size = 4
    return net
