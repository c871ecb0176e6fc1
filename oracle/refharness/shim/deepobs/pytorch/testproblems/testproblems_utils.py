"""Stand-in for deepobs 1.1.2 ``testproblems_utils`` (``requirements.txt:5`` of the reference).

TEST INFRASTRUCTURE ONLY.  TF-"same" padding is a separate ``ZeroPad2d`` whose padding is set
by a forward-pre-hook: out = ceil(in / stride), pad = max((out-1)*stride + k - in, 0),
left/top = pad // 2, the remainder right/bottom.
"""
import math
import torch
from torch import nn


def _pair(x):
    return (x, x) if isinstance(x, int) else tuple(x)


def _same_pad_hook(kernel_size, stride):
    kh, kw = _pair(kernel_size)
    sh, sw = _pair(stride)

    def hook(module, inp):
        pad_mod = module[0]
        H, W = inp[0].shape[2], inp[0].shape[3]
        ph = max((math.ceil(H / sh) - 1) * sh + kh - H, 0)
        pw = max((math.ceil(W / sw) - 1) * sw + kw - W, 0)
        pad_mod.padding = (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2)
    return hook


def tfconv2d(in_channels, out_channels, kernel_size, stride=1, dilation=1, groups=1, bias=True,
             tf_padding_type=None):
    modules = []
    if tf_padding_type == 'same':
        modules.append(nn.ZeroPad2d(0))
    modules.append(nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=0,
                             dilation=dilation, groups=groups, bias=bias))
    seq = nn.Sequential(*modules)
    if tf_padding_type == 'same':
        seq.register_forward_pre_hook(_same_pad_hook(kernel_size, stride))
    return seq


def tfmaxpool2d(kernel_size, stride=None, dilation=1, return_indices=False, ceil_mode=False,
                tf_padding_type=None):
    stride = kernel_size if stride is None else stride
    modules = []
    if tf_padding_type == 'same':
        modules.append(nn.ZeroPad2d(0))
    modules.append(nn.MaxPool2d(kernel_size, stride=stride, padding=0, dilation=dilation,
                                return_indices=return_indices, ceil_mode=ceil_mode))
    seq = nn.Sequential(*modules)
    if tf_padding_type == 'same':
        seq.register_forward_pre_hook(_same_pad_hook(kernel_size, stride))
    return seq


def _truncated_normal_init(tensor, mean=0, stddev=1):
    with torch.no_grad():
        nn.init.trunc_normal_(tensor, mean=mean, std=stddev, a=mean - 2 * stddev,
                              b=mean + 2 * stddev)
    return tensor
