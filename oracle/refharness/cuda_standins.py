"""Stand-ins that let the UNMODIFIED function-space code of the reference (``preds/gps.py:664-927``,
``preds/laplace.py:138-207``) run in the CPU-only build container.

TEST INFRASTRUCTURE ONLY (golden generation).  ``gp_quantities`` hard-codes CUDA: ``'cuda:{}'.format(device_ids[0])``,
``pin_memory``, ``.cuda()`` and the ``nn.parallel`` scatter / replicate / parallel_apply / gather primitives.  None of
that is arithmetic: inside ``cpu_as_cuda()`` every ``'cuda*'`` device means the CPU, pinning is a no-op and the parallel
primitives run their single replica in place, so the reference's own einsums, Cholesky factorisations and solves produce
the fixtures.
"""
import contextlib

import torch
from torch import nn


def _dev(d):
    if isinstance(d, str) and d.startswith('cuda'):
        return 'cpu'
    if isinstance(d, torch.device) and d.type == 'cuda':
        return torch.device('cpu')
    return d


def _map_args(args, kwargs):
    args = tuple(_dev(a) for a in args)
    kwargs = {k: _dev(v) for k, v in kwargs.items()}
    return args, kwargs


def _replicate(module, device_ids, detach=False):
    return [module for _ in device_ids]


def _scatter(inputs, device_ids, dim=0):
    assert len(device_ids) == 1
    return [inputs]


def _parallel_apply(modules, inputs, kwargs_tup=None, devices=None):
    outs = []
    for i, (m, inp) in enumerate(zip(modules, inputs)):
        kw = kwargs_tup[i] if kwargs_tup is not None else {}
        if not isinstance(inp, (list, tuple)):
            inp = (inp,)
        outs.append(m(*inp, **kw))
    return outs


def _gather(outputs, target_device, dim=0):
    assert len(outputs) == 1
    return outputs[0]


@contextlib.contextmanager
def cpu_as_cuda():
    saved = (torch.Tensor.to, nn.Module.to, torch.Tensor.cuda, torch.Tensor.pin_memory, torch.zeros, torch.tensor,
             nn.parallel.replicate, nn.parallel.scatter, nn.parallel.parallel_apply, nn.parallel.gather)
    t_to, m_to, _, _, zeros, tensor = saved[:6]

    def tensor_to(self, *a, **k):
        a, k = _map_args(a, k)
        return t_to(self, *a, **k)

    def module_to(self, *a, **k):
        a, k = _map_args(a, k)
        return m_to(self, *a, **k)

    def zeros_(*a, **k):
        k.pop('pin_memory', None)
        a, k = _map_args(a, k)
        return zeros(*a, **k)

    def tensor_(*a, **k):
        a, k = _map_args(a, k)
        return tensor(*a, **k)

    torch.Tensor.to, nn.Module.to = tensor_to, module_to
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.zeros, torch.tensor = zeros_, tensor_
    nn.parallel.replicate, nn.parallel.scatter = _replicate, _scatter
    nn.parallel.parallel_apply, nn.parallel.gather = _parallel_apply, _gather
    try:
        yield
    finally:
        (torch.Tensor.to, nn.Module.to, torch.Tensor.cuda, torch.Tensor.pin_memory, torch.zeros, torch.tensor,
         nn.parallel.replicate, nn.parallel.scatter, nn.parallel.parallel_apply, nn.parallel.gather) = saved
