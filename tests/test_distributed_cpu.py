"""World-size-2 test (gloo, CPU) of the N>1 host logic: which loader batches a rank owns, the
global example count behind the KFAC M/N weight, and the single all-reduce that joins the factors
(prior added once, after the reduce).  The per-rank partial sums come from the CPU oracle here --
the CUDA kernels are covered by the -m gpu tests -- so this checks exactly the code that
`Laplace.infer(shard=...)` runs around them (preds_b200/parallel.py)."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port_no, out_dir):
    for p in (ROOT, os.path.join(ROOT, 'bnn-predictions_b200')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from oracle import port
    from preds_b200 import parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port_no))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(71)
    model = port.make_mlps(3, [6], 2, 'tanh')
    lh = port.CategoricalLh()
    g = torch.Generator().manual_seed(15)
    X = torch.randn(50, 3, generator=g)
    y = torch.randint(2, (50,), generator=g)
    batches = [(X[:16], y[:16]), (X[16:32], y[16:32]), (X[32:41], y[32:41]), (X[41:], y[41:])]
    mine = [b for i, b in enumerate(batches) if parallel.owns_batch(i, 'batches', rank, world)]
    P = port.n_params(model)
    prior = 0.7
    # full: local sum on a zero buffer, one all-reduce, prior once
    acc = port.ggn_full_accumulate(model, lh, mine, torch.zeros(P, P))
    parallel.reduce_sum_(acc, port.expand_prior_precision(prior, P, 'full'))
    # diag
    dacc = torch.zeros(P)
    for Xb, _ in mine:
        dacc += lh.hess_factor() * port.diag_ggn_batch(model, lh, Xb)
    parallel.reduce_sum_(dacc, port.expand_prior_precision(prior, P, 'diag'))
    # kron: global N, then one flat bucket
    N = parallel.global_count(sum(len(b[0]) for b in mine), torch.device('cpu'))
    kron = port.KronState(model, prior)
    for Xb, yb in mine:
        kron.update(port.kflr_batch(model, lh, Xb), factor=lh.hess_factor(), batch_factor=len(yb) / N)
    parallel.reduce_factors_(kron.qhs)
    # predictive: shard test inputs, gather outputs
    Xte = torch.randn(7, 3, generator=g)
    out = parallel.sharded_predictive(lambda xs, S: xs.unsqueeze(0).repeat(S, 1, 1) * 2.0, Xte, 3)
    torch.save({'H': acc, 'd': dacc, 'N': N, 'qhs': kron.qhs, 'out': out, 'n_mine': len(mine)},
               os.path.join(out_dir, 'r%d.pt' % rank))
    dist.destroy_process_group()


def test_world_size_2_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle import port
    torch.manual_seed(71)
    model = port.make_mlps(3, [6], 2, 'tanh')
    lh = port.CategoricalLh()
    g = torch.Generator().manual_seed(15)
    X = torch.randn(50, 3, generator=g)
    y = torch.randint(2, (50,), generator=g)
    batches = [(X[:16], y[:16]), (X[16:32], y[16:32]), (X[32:41], y[32:41]), (X[41:], y[41:])]
    full = port.infer(model, 0.7, lh, batches, 'full')
    diag = port.infer(model, 0.7, lh, batches, 'diag')
    kron = port.infer(model, 0.7, lh, batches, 'kron')['kron']
    Xte = torch.randn(7, 3, generator=g)
    res = [torch.load(os.path.join(str(tmp_path), 'r%d.pt' % r), weights_only=False) for r in range(world)]
    assert [r['n_mine'] for r in res] == [2, 2]
    for r in res:
        assert r['N'] == 50
        assert torch.allclose(r['H'], full['precision'], rtol=1e-5, atol=1e-6)
        assert torch.allclose(r['d'], diag['precision'], rtol=1e-5, atol=1e-6)
        for qa, qb in zip(r['qhs'], kron.qhs):
            for a, b in zip(qa, qb):
                assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
        assert torch.equal(r['out'], Xte.unsqueeze(0).repeat(3, 1, 1) * 2.0)
    # every rank ends with bit-identical sums (no broadcast needed before the factorisation)
    assert torch.equal(res[0]['H'], res[1]['H'])
