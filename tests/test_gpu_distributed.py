"""Two-GPU run of the sharded path (NCCL): `Laplace.infer(shard='batches')` on 2 ranks must agree
with the single-process result and leave bit-identical posteriors on both ranks; the sharded GLM
predictive must agree with the unsharded one.  Skipped on boxes with fewer than 2 GPUs."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _spawn(fn, out_dir, nprocs=2):
    """mp.spawn with a fresh rendezvous port; a port that another process grabbed between the probe and the bind
    ("address already in use") is retried, anything else propagates."""
    for attempt in range(3):
        try:
            mp.spawn(fn, args=(nprocs, _free_port(), out_dir), nprocs=nprocs, join=True)
            return
        except Exception as exc:
            if attempt == 2 or 'address already in use' not in str(exc).lower():
                raise


def _problem():
    from preds_b200 import MLPS
    torch.manual_seed(71)
    model = MLPS(4, [9, 7], 3, 'tanh')
    g = torch.Generator().manual_seed(15)
    X = torch.randn(90, 4, generator=g)
    y = torch.randint(3, (90,), generator=g)
    batches = [(X[:32], y[:32]), (X[32:64], y[32:64]), (X[64:80], y[64:80]), (X[80:], y[80:])]
    Xte = torch.randn(9, 4, generator=g)
    eps = torch.randn(5, 9, 3, generator=g)
    return model, batches, Xte, eps


def _worker(rank, world, port_no, out_dir):
    for p in (ROOT, os.path.join(ROOT, 'bnn-predictions_b200')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from preds_b200 import Laplace, CategoricalLh, parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port_no))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    model, batches, Xte, eps = _problem()
    model = model.cuda()
    res = {}
    for cov in ('full', 'diag', 'kron'):
        lap = Laplace(model, 0.6, CategoricalLh())
        lap.infer(batches, cov_type=cov, shard='batches')
        if cov == 'kron':
            res[cov] = [[m.cpu() for m in qh] for qh in lap.Sigma_chol.qhs]
        else:
            res[cov] = lap.precision.cpu()
        lo, hi = parallel.shard_bounds(len(Xte), rank, world)
        out = parallel.sharded_predictive(
            lambda xs, S: lap.predictive_samples_glm(xs, n_samples=S, eps=eps[:, lo:hi]), Xte, 5)
        res['glm_' + cov] = out.cpu()
    torch.save(res, os.path.join(out_dir, 'r%d.pt' % rank))
    dist.destroy_process_group()


def test_two_gpu_sharded_infer_and_predictive(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    from preds_b200 import Laplace, CategoricalLh
    _spawn(_worker, str(tmp_path))
    res = [torch.load(os.path.join(str(tmp_path), 'r%d.pt' % r), weights_only=False) for r in range(2)]
    model, batches, Xte, eps = _problem()
    model = model.cuda()
    for cov in ('full', 'diag', 'kron'):
        lap = Laplace(model, 0.6, CategoricalLh())
        lap.infer(batches, cov_type=cov)
        ref_glm = lap.predictive_samples_glm(Xte, n_samples=5, eps=eps).cpu()
        for r in res:
            if cov == 'kron':
                for qa, qb in zip(r[cov], lap.Sigma_chol.qhs):
                    for a, b in zip(qa, qb):
                        assert torch.allclose(a, b.cpu(), rtol=1e-4, atol=1e-6)
            else:
                assert torch.allclose(r[cov], lap.precision.cpu(), rtol=1e-4, atol=1e-6)
            assert torch.allclose(r['glm_' + cov], ref_glm, rtol=1e-3, atol=1e-5)
        if cov != 'kron':
            assert torch.equal(res[0][cov], res[1][cov])      # identical sums on every rank


def _exchange_worker(rank, world, port_no, out_dir):
    for p in (ROOT, os.path.join(ROOT, 'bnn-predictions_b200')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from preds_b200 import Laplace, CategoricalLh, MLPS, parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port_no))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    torch.manual_seed(11)
    model = MLPS(300, [40, 24], 7, 'tanh').cuda()           # P = 13 199: tensor-core engine, several row chunks
    g = torch.Generator().manual_seed(12)
    X, y = torch.randn(1500, 300, generator=g), torch.randint(7, (1500,), generator=g)
    batches = [(X[i:i + 250], y[i:i + 250]) for i in range(0, 1500, 250)]
    prior_vec = torch.rand(13199, generator=g) + 0.5
    res = {}
    _, hdl = parallel.symmetric_zeros(8, torch.device('cuda', rank))
    modes = ['nccl'] + (['p2p'] if hdl is not None else []) + (['nvls'] if hdl is not None and int(hdl.multicast_ptr) else [])
    res['modes'] = modes
    for mode in modes:
        for pname, prior in (('scalar', 0.6), ('vector', prior_vec)):
            lap = Laplace(model, prior, CategoricalLh())
            lap.infer(batches, cov_type='full', shard='batches', factorise=False, exchange=mode)
            res[mode + '_' + pname] = lap.precision.cpu()
            res[mode + '_chunks'] = len(lap.exchange_stats.events)
            lap.precision = None
    # the one-off factorisation with the inverse split between the ranks (dense.spd_inverse_sharded, P >= dense.MIN_P)
    lap = Laplace(model, 0.6, CategoricalLh())
    lap.infer(batches, cov_type='full', shard='batches')
    res['Sigma_rows'] = lap.Sigma[::97].cpu()
    res['Sigma_chol_rows'] = lap.Sigma_chol[::97].cpu()
    lap.release_posterior()
    # a rank without examples still takes part (rank 1 owns nothing when there is one batch)
    lap = Laplace(model, 0.6, CategoricalLh())
    lap.infer(batches[:1], cov_type='full', shard='batches', factorise=False)
    res['one_batch'] = lap.precision.cpu()
    lap = Laplace(model, 0.6, CategoricalLh())
    lap.infer(batches[:1], cov_type='kron', shard='batches', factorise=False)
    res['one_batch_kron'] = [[m.cpu() for m in qh] for qh in lap.Sigma_chol.qhs]
    torch.save(res, os.path.join(out_dir, 'x%d.pt' % rank))
    dist.destroy_process_group()


def test_two_gpu_lower_triangle_exchange(tmp_path):
    """The full-covariance exchange (own kernel over peer memory: NVLS multimem / P2P, and the packed NCCL path), row
    chunks overlapped with the accumulation: every mode must give the single-process H (1e-4), identical bits on both
    ranks, the prior exactly once on the diagonal; a rank without examples must not hang the others."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    from preds_b200 import Laplace, CategoricalLh, MLPS
    _spawn(_exchange_worker, str(tmp_path))
    res = [torch.load(os.path.join(str(tmp_path), 'x%d.pt' % r), weights_only=False) for r in range(2)]
    torch.manual_seed(11)
    model = MLPS(300, [40, 24], 7, 'tanh').cuda()
    g = torch.Generator().manual_seed(12)
    X, y = torch.randn(1500, 300, generator=g), torch.randint(7, (1500,), generator=g)
    batches = [(X[i:i + 250], y[i:i + 250]) for i in range(0, 1500, 250)]
    prior_vec = torch.rand(13199, generator=g) + 0.5
    print('exchange modes available:', res[0]['modes'])
    for pname, prior in (('scalar', 0.6), ('vector', prior_vec)):
        lap = Laplace(model, prior, CategoricalLh())
        lap.infer(batches, cov_type='full', factorise=False)
        ref = lap.precision.cpu()
        for mode in res[0]['modes']:
            a, b = res[0][mode + '_' + pname], res[1][mode + '_' + pname]
            assert torch.equal(a, b), (mode, pname)
            assert torch.equal(a, a.T), (mode, pname)
            assert float((a - ref).abs().max() / ref.abs().max()) < 1e-4, (mode, pname)
            # the prior entered exactly once: compare the diagonal on its own scale
            assert float((a.diagonal() - ref.diagonal()).abs().max() / ref.diagonal().abs().max()) < 1e-4, (mode, pname)
            assert res[0][mode + '_chunks'] > 1
    lap = Laplace(model, 0.6, CategoricalLh())
    lap.infer(batches, cov_type='full')
    Sref, Lref = lap.Sigma[::97].cpu(), lap.Sigma_chol[::97].cpu()
    assert torch.equal(res[0]['Sigma_rows'], res[1]['Sigma_rows'])            # identical bits on every rank
    assert float((res[0]['Sigma_rows'] - Sref).abs().max() / Sref.abs().max()) < 1e-4
    assert float((res[0]['Sigma_chol_rows'] - Lref).abs().max() / Lref.abs().max()) < 1e-4
    lap = Laplace(model, 0.6, CategoricalLh())
    lap.infer(batches[:1], cov_type='full', factorise=False)
    ref = lap.precision.cpu()
    for r in res:
        assert float((r['one_batch'] - ref).abs().max() / ref.abs().max()) < 1e-4
    lap = Laplace(model, 0.6, CategoricalLh())
    lap.infer(batches[:1], cov_type='kron', factorise=False)
    for r in res:
        for qa, qb in zip(r['one_batch_kron'], lap.Sigma_chol.qhs):
            for a, b in zip(qa, qb):
                assert torch.allclose(a, b.cpu(), rtol=1e-4, atol=1e-6)


def _factor_worker(rank, world, port_no, out_dir):
    for p in (ROOT, os.path.join(ROOT, 'bnn-predictions_b200')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from preds_b200 import factor
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port_no))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    g = torch.Generator(device='cuda').manual_seed(3)
    P = 5003
    A = torch.randn(P, 700, device='cuda', generator=g)
    H = A @ A.T
    H.diagonal().add_(float(H.diagonal().mean()) * 0.02)
    dist.broadcast(H, 0)                                   # (two GPUs may round the product differently)
    chol1, Sigma1 = factor.posterior_from_precision(H)
    chol2, Sigma2 = factor.posterior_from_precision(H, shared=True)
    L1 = factor.cholesky(H)
    L2 = factor.cholesky(H, shared=True)
    torch.save({'same_L': bool(torch.equal(L1, L2)), 'same_Sigma': bool(torch.equal(Sigma1, Sigma2)),
                'same_chol': bool(torch.equal(chol1, chol2)), 'Sigma_rows': Sigma2[::53].cpu(),
                'chol_rows': chol2[::53].cpu()}, os.path.join(out_dir, 'f%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_shared_factorisation(tmp_path):
    """The one-off factorisation shared between two ranks (tiles of every product dealt out, panels / planes / Sigma
    gathered): bit-identical to the single-process chain on both ranks."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    _spawn(_factor_worker, str(tmp_path))
    res = [torch.load(os.path.join(str(tmp_path), 'f%d.pt' % r), weights_only=False) for r in range(2)]
    for r in res:
        assert r['same_L'] and r['same_Sigma'] and r['same_chol'], {k: v for k, v in r.items() if k.startswith('same')}
    assert torch.equal(res[0]['Sigma_rows'], res[1]['Sigma_rows'])
    assert torch.equal(res[0]['chol_rows'], res[1]['chol_rows'])
