"""Model zoo the Jacobians are taken over: ``MLPS`` (preds/models.py:91-122), ``CIFAR10Net``
(:239-285), ``CIFAR100Net`` (:182-236), the functional-activation ``MLP`` / ``SiMLP`` (:124-176) of the
UCI classification driver, with deepobs' ``tfconv2d`` / ``tfmaxpool2d`` building
blocks re-created here (TF-"same" padding as a ``ZeroPad2d`` set by a forward-pre-hook) so that
models built against the reference run unchanged.  These are plain ``nn.Module`` containers; the
compute goes through the descriptor in ``_engine.py``.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn


def _pair(x):
    return (x, x) if isinstance(x, int) else tuple(x)


def _tf_same_hook(kernel_size, stride):
    (kh, kw), (sh, sw) = _pair(kernel_size), _pair(stride)

    def hook(module, inp):
        H, W = inp[0].shape[2], inp[0].shape[3]
        ph = max((math.ceil(H / sh) - 1) * sh + kh - H, 0)
        pw = max((math.ceil(W / sw) - 1) * sw + kw - W, 0)
        module[0].padding = (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2)
    return hook


def tfconv2d(in_channels, out_channels, kernel_size, stride=1, dilation=1, groups=1, bias=True,
             tf_padding_type=None):
    mods = [nn.ZeroPad2d(0)] if tf_padding_type == 'same' else []
    mods.append(nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=0,
                          dilation=dilation, groups=groups, bias=bias))
    seq = nn.Sequential(*mods)
    if tf_padding_type == 'same':
        seq.register_forward_pre_hook(_tf_same_hook(kernel_size, stride))
    return seq


def tfmaxpool2d(kernel_size, stride=None, dilation=1, return_indices=False, ceil_mode=False,
                tf_padding_type=None):
    stride = kernel_size if stride is None else stride
    mods = [nn.ZeroPad2d(0)] if tf_padding_type == 'same' else []
    mods.append(nn.MaxPool2d(kernel_size, stride=stride, padding=0, dilation=dilation,
                             return_indices=return_indices, ceil_mode=ceil_mode))
    seq = nn.Sequential(*mods)
    if tf_padding_type == 'same':
        seq.register_forward_pre_hook(_tf_same_hook(kernel_size, stride))
    return seq


class MLP(nn.Module):
    """preds/models.py:124-165: activations are plain functions (``self.act``), hidden layers live in
    ``self.hidden_layers`` and the head in ``self.output_layer``; a single output is returned flattened
    ([N] instead of [N,1]).  The engine reads exactly these three attributes (``_engine._functional_mlp``)."""
    _ACTS = {'relu': torch.relu, 'tanh': torch.tanh, 'sigmoid': torch.sigmoid, 'selu': F.elu}

    def __init__(self, input_size, hidden_sizes, output_size, activation='tanh', **kwargs):
        super().__init__()
        self.input_size = input_size
        self.hidden_sizes = hidden_sizes
        self.output_size = 1 if output_size is None else output_size
        self.act = self._ACTS[activation]
        widths = [input_size] + list(hidden_sizes)
        if hidden_sizes:
            self.hidden_layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(widths[:-1], widths[1:]))
        else:
            self.hidden_layers = []
        self.output_layer = nn.Linear(widths[-1], self.output_size)

    def forward(self, x):
        h = x.view(-1, self.input_size)
        for layer in self.hidden_layers:
            h = self.act(layer(h))
        z = self.output_layer(h)
        return z.flatten() if self.output_size == 1 else z


class SiMLP(MLP):
    """preds/models.py:168-176: width x depth parametrisation of ``MLP``."""

    def __init__(self, input_size, output_size, n_layers, n_units, activation='tanh', **kwargs):
        super().__init__(input_size, [n_units] * n_layers, output_size, activation, **kwargs)


class MLPS(nn.Sequential):
    def __init__(self, input_size, hidden_sizes, output_size, activation='tanh', flatten=False,
                 bias=True):
        super().__init__()
        self.input_size = input_size
        self.hidden_sizes = hidden_sizes
        self.activation = activation
        self.output_size = output_size if output_size is not None else 1
        if activation not in ('relu', 'tanh'):
            raise ValueError('invalid activation')
        act = nn.ReLU if activation == 'relu' else nn.Tanh
        if flatten:
            self.add_module('flatten', nn.Flatten())
        if len(hidden_sizes) == 0:
            self.add_module('lin_layer', nn.Linear(input_size, self.output_size, bias=bias))
            return
        widths = [input_size] + list(hidden_sizes)
        for i in range(len(hidden_sizes)):
            self.add_module('layer%d' % (i + 1), nn.Linear(widths[i], widths[i + 1], bias=bias))
            self.add_module('%s%d' % (activation, i + 1), act())
        self.add_module('out_layer', nn.Linear(widths[-1], self.output_size, bias=bias))


def _init_3c3d(net):
    for m in net.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.constant_(m.bias, 0.0)
            nn.init.xavier_normal_(m.weight)
        if isinstance(m, nn.Linear):
            nn.init.constant_(m.bias, 0.0)
            nn.init.xavier_uniform_(m.weight)


class CIFAR10Net(nn.Sequential):
    """DeepOBS ``net_cifar10_3c3d``."""

    def __init__(self, in_channels=3, n_out=10, use_tanh=False):
        super().__init__()
        self.output_size = n_out
        activ = nn.Tanh if use_tanh else nn.ReLU
        spec = [(in_channels, 64, 5, None), (64, 96, 3, None), (96, 128, 3, 'same')]
        for i, (cin, cout, k, pad) in enumerate(spec, 1):
            self.add_module('conv%d' % i, tfconv2d(cin, cout, k, tf_padding_type=pad))
            self.add_module('relu%d' % i, nn.ReLU())
            self.add_module('maxpool%d' % i, tfmaxpool2d(3, stride=2, tf_padding_type='same'))
        self.add_module('flatten', nn.Flatten())
        self.add_module('dense1', nn.Linear(3 * 3 * 128, 512))
        self.add_module('relu4', activ())
        self.add_module('dense2', nn.Linear(512, 256))
        self.add_module('relu5', activ())
        self.add_module('dense3', nn.Linear(256, n_out))
        _init_3c3d(self)


class CIFAR100Net(nn.Sequential):
    """DeepOBS ``net_cifar100_allcnnc`` (All-CNN-C)."""

    def __init__(self, in_channels=3, n_out=100):
        super().__init__()
        assert n_out in (10, 100)
        self.output_size = n_out
        spec = [(in_channels, 96, 3, 1, 'same'), (96, 96, 3, 1, 'same'), (96, 96, 3, 2, 'same'),
                (96, 192, 3, 1, 'same'), (192, 192, 3, 1, 'same'), (192, 192, 3, 2, 'same'),
                (192, 192, 3, 1, None), (192, 192, 1, 1, 'same'), (192, n_out, 1, 1, 'same')]
        for i, (cin, cout, k, s, pad) in enumerate(spec, 1):
            self.add_module('conv%d' % i, tfconv2d(cin, cout, k, stride=(s, s), tf_padding_type=pad))
            if i < len(spec):
                self.add_module('relu%d' % i, nn.ReLU())
        self.add_module('avg', nn.AvgPool2d(kernel_size=(6, 6)))
        self.add_module('flatten', nn.Flatten())
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.constant_(m.bias, 0.1)
                nn.init.xavier_normal_(m.weight)
