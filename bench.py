#!/usr/bin/env python
"""Benchmark of the Laplace-GGN hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py --gpus N --steps K --warmup W          # our arm (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W   # reference CPU arm (oracle port)

Workload (BASELINE.json configs[4], the largest single-GPU configuration): full-covariance GGN of
MLPS(784,[75],10,'tanh',flatten=True), P = 59 635, K = 10, synthetic N(0,1) inputs of MNIST shape.
One step = one loader batch of B examples through the hot path: forward, K-column backward,
likelihood Hessian, tensor-core accumulation  H += sum_n J_n^T Lambda_n J_n  into the resident
14.2 GB precision matrix.  `value` is device-timed with the batches resident in HBM; `e2e` is the
same metric through `Laplace.infer(...)` with pinned HOST batches (H2D + a D2H read per step).
The one-off factorisation is not part of this metric (timed separately, north star).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, 'bnn-predictions_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = 'ggn_accum_examples_per_s'
UNIT = 'examples/s'
MODEL = dict(input_size=784, hidden_sizes=[75], output_size=10, activation='tanh', flatten=True)
P_PARAMS, K_OUT = 59635, 10
WORKLOAD = 'cfg5 full-cov GGN, MLPS(784,[75],10,tanh) P=59635 K=10, synthetic MNIST-shape N(0,1)'


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i] == 'Active' for r in self.rows)]
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def make_batches(n_batches, B, pinned=False):
    out = []
    for i in range(n_batches):
        g = torch.Generator().manual_seed(15 + i)
        X = torch.randn(B, 1, 28, 28, generator=g)
        y = torch.randint(K_OUT, (B,), generator=g)
        if pinned:
            X, y = X.pin_memory(), y.pin_memory()
        out.append((X, y))
    return out


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------
def cpu_full_ggn_rate(sample, repeats, warmup):
    """examples/s of the reference's full-GGN accumulation (Jacobians + einsum, preds/laplace.py:39-42)
    restated in oracle/port.py, float32 torch on all host threads, on `sample` examples per step."""
    from oracle import port
    torch.manual_seed(71)
    model = port.make_mlps(**MODEL)
    lh = port.CategoricalLh()
    g = torch.Generator().manual_seed(15)
    X = torch.randn(sample, 1, 28, 28, generator=g)
    y = torch.randint(K_OUT, (sample,), generator=g)
    H = torch.zeros(P_PARAMS, P_PARAMS)
    times = []
    for it in range(warmup + repeats):
        t0 = time.perf_counter()
        port.ggn_full_accumulate(model, lh, [(X, y)], H)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sample / statistics.mean(times), statistics.mean(times)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    sample = args.cpu_sample
    rate, sec = cpu_full_ggn_rate(sample, args.steps, args.warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'batch': sample},
        'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                         'sample': '%d examples per step (Jacobians + einsum into the 14.2 GB H), oracle/port.py' % sample},
        'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def cpu_glm_rate(Sigma_dev, n_inputs, n_samples):
    """(inputs x samples)/s of the reference's GLM predictive at cfg5 on the host cores (oracle/port.py: Jacobians,
    einsum('nkp,pq,ncq->nkc') with the full covariance, K x K Cholesky, sampling + softmax; preds/predictives.py:50-76)
    on `n_inputs` inputs x `n_samples` samples, as experiments/imginference.py:45,391 batches it.  The reference also
    recomputes Sigma = L L^T (2 P^3 flops) on EVERY call (preds/laplace.py:111); that is left out here (it would
    take minutes at P = 59 635), so the figure flatters the CPU path."""
    from oracle import port
    torch.manual_seed(71)
    model = port.make_mlps(**MODEL)
    Sigma = Sigma_dev.cpu()
    g = torch.Generator().manual_seed(16)
    X = torch.randn(n_inputs, 1, 28, 28, generator=g)
    eps = torch.randn(n_samples, n_inputs, K_OUT, generator=g)
    t0 = time.perf_counter()
    Js, f = port.jacobians(model, X)
    Jt = Js.transpose(1, 2)
    f_var = torch.einsum('nkp,pq,ncq->nkc', Jt, Sigma, Jt)
    L = torch.linalg.cholesky(f_var)
    out = torch.softmax(f.unsqueeze(0) + torch.einsum('nkl,snl->snk', L, eps), dim=-1)
    float(out.mean())
    sec = time.perf_counter() - t0
    return n_inputs * n_samples / sec, sec


def strong_scaling(dev, rank, world, args):
    """Strong-scaling records of the configurations as BASELINE.json states them: the TOTAL work is fixed and sharded
    over the ranks, through the public API with pinned host batches (H2D inside), the exchange step included:
      cfg5  60 000 examples, full covariance  -> Laplace.infer(shard='presharded'), own lower-triangle exchange kernel
      cfg4  50 000 examples, CIFAR10Net KFAC  -> Laplace.infer(cov_type='kron', shard='presharded'), NCCL bucket
    Each is run twice (the first run warms allocator / communicator); the second is timed by wall clock between
    barriers, max over ranks.  N = 1 gives the baseline the driver's per-N files are compared with."""
    import torch.distributed as dist
    from preds_b200 import Laplace, CategoricalLh, MLPS, CIFAR10Net, parallel
    out = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def my_batches(total, bs, shape, classes, seed):
        lo, hi = parallel.shard_bounds(total, rank, world)
        g = torch.Generator().manual_seed(seed + rank)
        X = torch.randn((hi - lo,) + shape, generator=g).pin_memory()
        y = torch.randint(classes, (hi - lo,), generator=g).pin_memory()
        return [(X[i:i + bs], y[i:i + bs]) for i in range(0, hi - lo, bs)], hi - lo

    shard = 'presharded' if world > 1 else None
    try:
        torch.manual_seed(71)
        model = MLPS(**MODEL).to(dev)
        batches, n_mine = my_batches(60000, 2500, (1, 28, 28), K_OUT, 1500)
        lap = Laplace(model, 1.0, CategoricalLh())
        lap.ggn_engine = args.engine
        rec = None
        for rep in range(2):
            barrier()
            t0 = time.perf_counter()
            lap.infer(batches, cov_type='full', shard=shard, factorise=False)
            trace = float(lap.precision.diagonal().sum())
            barrier()
            sec = maxr(time.perf_counter() - t0)
            rec = {'examples_total': 60000, 'examples_this_rank': n_mine, 'loader_batch': 2500, 'seconds': sec,
                   'examples_per_s': 60000 / sec, 'accumulate_ms_device': maxr(lap.timings['accumulate_ms']),
                   'trace_H': trace}
            if world > 1 and lap.exchange_stats is not None:
                ms, nbytes = lap.exchange_stats.stats()
                ms = maxr(ms)
                rec.update({'exchange_mode': lap.exchange_stats.mode, 'exchange_chunks': len(lap.exchange_stats.events),
                            'allreduce_ms': ms, 'allreduce_bytes': nbytes,
                            'allreduce_busbw_GBs': 2 * (world - 1) / world * nbytes / (ms * 1e-3) / 1e9,
                            'allreduce_note': 'sum over the row chunks of the CUDA-event time of the own exchange kernel '
                                              '(lower triangle only); all but the last chunk overlap the accumulation'})
            lap.release_posterior()
        out['cfg5_full'] = rec
        del lap, batches, model
    except Exception as exc:
        out['cfg5_full'] = {'error': repr(exc)[:300]}
    try:
        torch.manual_seed(71)
        model = CIFAR10Net(3, 10, use_tanh=True).to(dev)
        batches, n_mine = my_batches(50000, 250, (3, 32, 32), 10, 2500)
        lap = Laplace(model, 1.0, CategoricalLh())
        rec = None
        for rep in range(2):
            barrier()
            t0 = time.perf_counter()
            lap.infer(batches, cov_type='kron', shard=shard, factorise=False)
            chk = float(sum(m.double().sum() for qh in lap.Sigma_chol.qhs for m in qh))
            barrier()
            sec = maxr(time.perf_counter() - t0)
            rec = {'examples_total': 50000, 'examples_this_rank': n_mine, 'loader_batch': 250, 'seconds': sec,
                   'examples_per_s': 50000 / sec, 'accumulate_ms_device': maxr(lap.timings['accumulate_ms']),
                   'checksum_factors': chk}
            if world > 1 and lap.exchange_stats:
                xt = lap.exchange_stats
                ms = maxr(xt['e0'].elapsed_time(xt['e1']))
                rec.update({'allreduce_ms': ms, 'allreduce_bytes': xt['bytes'],
                            'allreduce_note': 'one NCCL all-reduce of the flat bucket of all Kronecker factors'})
        out['cfg4_kron'] = rec
    except Exception as exc:
        out['cfg4_kron'] = {'error': repr(exc)[:300]}
    return out


def run_ours(args):
    import torch.distributed as dist
    from preds_b200 import Laplace, CategoricalLh, MLPS, parallel
    from preds_b200._cabi import lib, check, ptr, stream
    from preds_b200._engine import net_for

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B, K, W = args.batch, args.steps, args.warmup

    torch.manual_seed(71)
    model = MLPS(**MODEL).to(dev)
    lh = CategoricalLh()
    net = net_for(model)
    n_distinct = 4
    host = make_batches(n_distinct, B, pinned=True)
    for i in range(n_distinct):        # different examples on every rank (sharded data set)
        host[i][0].add_(0.01 * rank)
    resident = [(X.to(dev), y.to(dev)) for X, y in host]
    lap = Laplace(model, 1.0, lh)
    lap.ggn_engine = args.engine
    hdl = None
    if world > 1:
        H, hdl = parallel.symmetric_zeros(P_PARAMS, dev)
    if world == 1 or H is None:
        H = torch.zeros(P_PARAMS, P_PARAMS, device=dev)

    def job(n_steps, exchange):
        """The hot path over n_steps resident batches through the product's own accumulation routine; with N > 1 the
        path's one exchange step (own kernel over NVLink, row chunks overlapped with the last batch) and the mirror."""
        ex = parallel.LowerExchange(H, hdl, 1.0) if exchange else None
        lap._accumulate_full(((resident[i % n_distinct][0], None) for i in range(n_steps)), H, exchange=ex)
        return ex

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    job(W, world > 1)                  # warm-up: kernels, allocator, and (N > 1) communicator + exchange kernel
    barrier()
    H.zero_()
    lib.bnnp_prof_enable(1)
    launches0 = lib.bnnp_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    ex = job(K, world > 1)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = int(lib.bnnp_launch_count() - launches0)
    # dominant kernel: CUDA-event records of every gram launch (recorded on the launching stream)
    cap = 4096
    pms, pex, pus = (C.c_float * cap)(), (C.c_double * cap)(), (C.c_double * cap)()
    nrec = lib.bnnp_prof_read(pms, pex, pus, cap)
    lib.bnnp_prof_enable(0)
    exchange = None
    if ex is not None:
        xms, xbytes = ex.stats()
        exchange = {'mode': ex.mode, 'chunks': len(ex.events), 'allreduce_ms': xms, 'allreduce_bytes': xbytes,
                    'busbw_GBs': 2 * (world - 1) / world * xbytes / (xms * 1e-3) / 1e9,
                    'note': 'own kernel over NVLink peer memory, lower triangle of H only, prior fused; all but the last '
                            'row chunk overlap the accumulation of the last batch'}
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = K * B * world / (ms * 1e-3)
    checksum = float(H.diagonal().double().sum())

    # ---- e2e through the public API with pinned host batches (H2D inside, D2H read per step) ----
    e2e_steps = max(1, min(K, args.e2e_steps))
    shard = 'presharded' if world > 1 else None     # N > 1: every rank infers on its own examples, exchange included
    for _ in range(2):      # warm-up: the second call makes the allocator hold both 14.2 GB blocks (old + new precision)
        lap.infer([host[0]], cov_type='full', factorise=False, shard=shard)
        float(lap.precision.diagonal().sum())
        lap.release_posterior()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        lap.infer([host[i % n_distinct]], cov_type='full', factorise=False, shard=shard)
        float(lap.precision.diagonal().sum())                  # device -> host read of the step's result
        lap.release_posterior()
    if world > 1:
        dist.barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = e2e_steps * B * world / e2e_s

    # ---- strong scaling of the configurations as stated (total work fixed) ----
    strong = None
    if not args.no_strong:
        strong = strong_scaling(dev, rank, world, args)

    # ---- secondary: the one-off factorisation (timed separately) and the GLM predictive ----------
    secondary = None
    Sigma_for_cpu = None
    if not args.no_glm:
        H.diagonal().add_(1.0 if world == 1 else 0.0)   # prior precision (N > 1: the exchange kernel already added it)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        if world > 1:
            # the shared chain gathers panels (fp32) and planes (bf16) through NCCL: warm both paths of the communicator
            for dt in (torch.float32, torch.bfloat16):
                wbuf = torch.zeros(1 << 27, dtype=dt, device=dev)
                dist.all_reduce(wbuf)
                del wbuf
        barrier()
        ev[0].record()
        Sigma_chol, Sigma = lap._factorise_full(H, sharded=world > 1)
        ev[1].record()
        del H
        lap.mu = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        lap.Sigma_chol, lap._Sigma, lap.inferred = Sigma_chol, Sigma, True
        nte, S = args.glm_inputs, args.glm_samples
        Xte_host = torch.randn(nte, 1, 28, 28, generator=torch.Generator().manual_seed(16 + rank)).pin_memory()
        Xte = Xte_host.to(dev)
        lap.predictive_samples_glm(Xte[:256], n_samples=8, seed=1)          # warm-up (also builds the Sigma split)
        barrier()
        lib.bnnp_prof_enable(1)
        ev[2].record()
        out = lap.predictive_samples_glm(Xte, n_samples=S, seed=2)
        ev[3].record()
        torch.cuda.synchronize()
        gms, gus = (C.c_float * 1024)(), (C.c_double * 1024)()
        gn = lib.bnnp_glm_prof_read(gms, gus, 1024)
        lib.bnnp_prof_enable(0)
        out.mean(0).cpu()                                                   # warm-up of the reduction kernels (first call: ~19 ms)
        barrier()
        t0 = time.perf_counter()
        out = lap.predictive_samples_glm(Xte_host, n_samples=S, seed=3)     # host inputs -> H2D inside
        pmean = out.mean(0).cpu()                                           # what the drivers keep (imginference.py:45)
        e2e_glm_s = time.perf_counter() - t0
        glm_ms = ev[2].elapsed_time(ev[3])
        del out
        # the same quantity with the mean over samples taken inside the sampling kernel (no [S,N,K] tensor)
        lap.predictive_mean_glm(Xte[:256], n_samples=8, seed=1)
        barrier()
        t0 = time.perf_counter()
        pmean_fused = lap.predictive_mean_glm(Xte_host, n_samples=S, seed=3).cpu()
        e2e_fused_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([glm_ms, e2e_glm_s, e2e_fused_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            glm_ms, e2e_glm_s, e2e_fused_s = float(t[0]), float(t[1]), float(t[2])
        pk, _ = peaks()
        fact_ms = ev[0].elapsed_time(ev[1])
        if world > 1:
            t = torch.tensor([fact_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            fact_ms = float(t[0])
        fact_tf = 4.0 * lap.P ** 3 / 6.0 * 12.0 / (fact_ms * 1e-3) / 1e12
        glm_roof = None
        if gn > 0:
            kms = sum(gms[i] for i in range(gn))
            kflops = sum(gus[i] for i in range(gn))
            peak = pk.get('bf16_tflops_sustained', pk['bf16_tflops'])
            glm_roof = {'bound': 'tensor', 'kernel': 'glm_fvar_pair_kernel (TMA-fed CTA-pair tcgen05.mma)', 'achieved': kflops / (kms * 1e-3) / 1e12,
                        'peak': peak, 'unit': 'TFLOP/s', 'frac': kflops / (kms * 1e-3) / 1e12 / peak, 'traffic': None,
                        'kernel_ms': kms, 'kernel_launches': gn, 'kernel_share_of_call': kms / glm_ms,
                        'flops_note': 'useful = 3 split-bf16 passes x 2 x inputs x elements of Sigma on/below the block diagonal '
                                      '(f_var = J Sigma J^T evaluated as T[n,u,u\'] = a^T Sigma_uu\' a\' for u\' <= u); '
                                      'as written the reference einsum costs 2 K P^2 + 2 K^2 P per input'}
        secondary = {
            'factorise_full_ms': fact_ms,
            'factorise_note': 'one-off chol(H), Sigma = H^-1, chol(Sigma) at P=59635 on the library\'s own three-term split-bf16 '
                              'tensor-core GEMM (csrc/gemm3.cu, preds_b200/factor.py; timed separately, north star (3)); '
                              'N > 1: the tiles of every product are dealt out between the ranks and gathered, every rank ends '
                              'with the bits of the single-GPU chain',
            'factorise_roofline': {'bound': 'tensor', 'kernel': 'gemm3_kernel (6 bf16 MMAs per product, fp32-grade)',
                                   'achieved': fact_tf, 'peak': world * pk.get('bf16_tflops_sustained', pk['bf16_tflops']),
                                   'unit': 'TFLOP/s', 'frac': fact_tf / (world * pk.get('bf16_tflops_sustained', pk['bf16_tflops'])),
                                   'fp32_equivalent_tflops': fact_tf / 6.0,
                                   'flops_note': 'whole chain incl. the library calls on the 1024 x 1024 diagonal blocks, gathers '
                                                 'and splits: 4 x P^3/6 MACs (two Cholesky, triangular inverse, Gram) x 12 bf16 '
                                                 'flop per MAC / wall time of the chain, whole job (all ranks)'},
            'glm_predictive': {
                'metric': 'glm_predictive_input_samples_per_s', 'value': nte * S * world / (glm_ms * 1e-3),
                'unit': '(inputs x samples)/s', 'inputs_per_gpu': nte, 'samples': S, 'ms': glm_ms,
                'roofline': glm_roof,
                'e2e_value': nte * S * world / e2e_glm_s,
                'e2e_note': 'pinned host inputs -> predictive_samples_glm -> .mean(0) read back',
                'e2e_fused_mean_value': nte * S * world / e2e_fused_s,
                'e2e_fused_mean_note': 'pinned host inputs -> predictive_mean_glm (mean over samples inside the sampler) read back',
                'fused_vs_materialised_max_abs_diff': float((pmean_fused - pmean).abs().max()),
                'prob_rows_sum_to_one': float((pmean.sum(1) - 1).abs().max()),
                'scaling': 'test inputs sharded across ranks, no collective',
            },
        }
        if world == 1 and rank == 0 and not args.no_cpu_baseline:
            Sigma_for_cpu = Sigma

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    if Sigma_for_cpu is not None:
        try:
            torch.set_num_threads(os.cpu_count())
            rate, sec = cpu_glm_rate(Sigma_for_cpu, 160, 100)
            secondary['glm_predictive']['cpu_baseline'] = {
                'value': rate, 'unit': '(inputs x samples)/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                'sample': '160 inputs x 100 samples (8 batches of 20 as experiments/imginference.py:45,391), Jacobians + einsum with the full '
                          '[P,P] covariance + sampling, oracle/port.py, %.1f s; the reference\'s per-call Sigma = L L^T '
                          '(2 P^3 flops) excluded' % sec}
        except Exception as exc:
            secondary['glm_predictive']['cpu_baseline'] = {'error': repr(exc)[:200]}
        Sigma_for_cpu = None

    if secondary is not None and world == 1 and not args.no_other_configs:
        lap = None
        torch.cuda.empty_cache()
        secondary['configs'] = other_configs(dev)
    if secondary is None:
        secondary = {}
    secondary['strong_scaling'] = strong
    secondary['exchange'] = exchange

    pk, pk_kind = peaks()
    recs = [(pms[i], pex[i], pus[i]) for i in range(max(nrec, 0))]
    roofline = None
    if recs:
        top = max(r[2] for r in recs)
        dom = [r for r in recs if r[2] >= 0.5 * top]           # the whole (layer1, layer1) block launches
        avg_ms = statistics.mean(r[0] for r in dom)
        useful = statistics.mean(r[2] for r in dom)
        executed = statistics.mean(r[1] for r in dom)
        peak = pk.get('bf16_tflops_sustained', pk['bf16_tflops'])
        ach = useful / (avg_ms * 1e-3) / 1e12
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, 'profiles', 'gram_wide_traffic.json')))
            traffic, traffic_src = tj['dram_bytes_per_launch'], tj.get('source')
        except Exception:
            pass
        roofline = {
            'bound': 'tensor', 'kernel': 'gram_wide_kernel (CTA-pair tcgen05.mma.cta_group::2, one unit x two column tiles)',
            'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s',
            'frac': ach / peak, 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': pk_kind + ' bf16 dense (sustained)',
            'kernel_ms_per_launch': avg_ms, 'kernel_share_of_step': sum(r[0] for r in recs) / ms,
            'flops_note': 'achieved = useful bf16 MMA flops (3 split passes x 2 x examples x elements of H on or below its '
                          'diagonal in the launch\'s rows, padding and the above-diagonal part of crossing tiles excluded) '
                          '/ CUDA-event kernel time',
            'executed_tflops_incl_padding': executed / (avg_ms * 1e-3) / 1e12,
            'einsum_as_written_tflops': (2.0 * K_OUT * P_PARAMS ** 2 + 2.0 * P_PARAMS * K_OUT ** 2) * B / (ms / K * 1e-3) / 1e12,
        }

    cpu = None
    if not args.no_cpu_baseline and world == 1:        # reported at N=1 only (rank 0)
        torch.set_num_threads(os.cpu_count())
        rate, sec = cpu_full_ggn_rate(4 * args.cpu_sample, 1, 1)
        cpu = {'value': rate, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': '%d examples, Jacobians + einsum into the 14.2 GB H (oracle/port.py), %.1f s' % (4 * args.cpu_sample, sec)}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16x2-split (fp32-grade) tensor cores, f32 accumulate', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'batch_per_gpu': B, 'engine': 'tcgen05' if args.engine != 1 else 'simt',
                   'l2': 'working set >> L2: the 14.2 GB H is read-modify-written every step',
                   'timed_region': 'K steps + exchange of the lower triangle of H (N>1, own NVLink kernel, overlapped with the last step) + symmetrise'},
        'roofline': roofline, 'cpu_baseline': cpu,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': B * 784 * 4 + B * 8,
                'd2h_bytes_per_step': 4, 'steps': e2e_steps,
                'call': "Laplace.infer([(X_pinned_host, y)], cov_type='full', factorise=False%s) + float(precision.trace)"
                        % (", shard='presharded'" if world > 1 else '')},
        'gpu_launches': launches, 'clocks': clocks, 'checksum_trace_H': checksum, 'secondary': secondary,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def other_configs(dev):
    """Device-timed rates of the other BASELINE.json configurations (3: MNIST-shape MLP kron/diag + GLM predictive on
    10 000 inputs x 1000 samples; 4: CIFAR10Net KFAC accumulation), reported under `secondary.configs`.  Never fatal for
    the headline line."""
    from preds_b200 import Laplace, CategoricalLh, MLPS, CIFAR10Net
    out = []

    def timed(fn, reps=3):
        best = None
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1)
            best = t if best is None else min(best, t)
        return best

    g = torch.Generator().manual_seed(15)
    try:
        torch.manual_seed(71)
        model = MLPS(784, [1024, 512, 256, 128], 10, flatten=True).to(dev)
        N, bs, nte, S = 4096, 256, 10000, 1000
        X = torch.randn(N, 1, 28, 28, generator=g)
        y = torch.randint(10, (N,), generator=g)
        batches = [(X[i:i + bs].to(dev), y[i:i + bs].to(dev)) for i in range(0, N, bs)]
        Xte = torch.randn(nte, 1, 28, 28, generator=g).to(dev)
        for cov in ('kron', 'diag'):
            lap = Laplace(model, 1.0, CategoricalLh())
            lap.infer(batches[:2], cov_type=cov, factorise=False)
            acc_ms = timed(lambda: lap.infer(batches, cov_type=cov, factorise=False))
            lap.infer(batches, cov_type=cov)
            lap.predictive_samples_glm(Xte[:64], n_samples=8, seed=1)
            glm_ms = timed(lambda: lap.predictive_samples_glm(Xte, n_samples=S, seed=2), reps=2)
            # SURVEY 8d work per unit: layer widths (n_l -> m_l) of the cfg3 MLP, K = 10
            dims = [(784, 1024), (1024, 512), (512, 256), (256, 128), (128, 10)]
            Kc = 10
            fwd = 2.0 * sum(n * m for n, m in dims)
            bwd = 2.0 * Kc * sum(n * m for n, m in dims[1:])
            fbytes = 4.0 * sum(n + Kc * m for n, m in dims)                       # 4 sum(size a + K size g) per example / input
            pkk, _ = peaks()
            tc_peak = pkk.get('bf16_tflops_sustained', pkk['bf16_tflops']) / 3.0      # fp32-grade = 3 bf16 MMA passes
            if cov == 'kron':
                acc_flop = sum(2.0 * n * n + 2.0 * Kc * m * m for n, m in dims) + fwd + bwd
                glm_flop = sum(2.0 * n * n + 2.0 * Kc * m * m + 2.0 * m * n + Kc * Kc * m for n, m in dims) + fwd + bwd
            else:
                acc_flop = sum(2.0 * m * n for n, m in dims) + fwd + bwd
                glm_flop = acc_flop
            acc_tf = acc_flop * N / (acc_ms * 1e-3) / 1e12
            glm_tf = glm_flop * nte / (glm_ms * 1e-3) / 1e12
            roof = {'accum': {'bound': 'tensor', 'achieved': acc_tf, 'peak': tc_peak, 'unit': 'TFLOP/s fp32-equivalent',
                              'frac': acc_tf / tc_peak, 'flop_per_example': acc_flop},
                    'glm': {'bound': 'tensor', 'achieved': glm_tf, 'peak': tc_peak, 'unit': 'TFLOP/s fp32-equivalent',
                            'frac': glm_tf / tc_peak, 'flop_per_input': glm_flop}}
            if cov == 'diag':           # north star: HBM-bound on the factor stream -- report against both rooflines
                roof['accum_hbm'] = {'bound': 'hbm', 'achieved': fbytes * N / (acc_ms * 1e-3) / 1e9, 'peak': pkk['hbm_gbs'],
                                     'unit': 'GB/s', 'frac': fbytes * N / (acc_ms * 1e-3) / 1e9 / pkk['hbm_gbs'],
                                     'bytes_per_example': fbytes}
                roof['glm_hbm'] = {'bound': 'hbm', 'achieved': fbytes * nte / (glm_ms * 1e-3) / 1e9, 'peak': pkk['hbm_gbs'],
                                   'unit': 'GB/s', 'frac': fbytes * nte / (glm_ms * 1e-3) / 1e9 / pkk['hbm_gbs'],
                                   'bytes_per_input': fbytes}
            out.append({'config': 'cfg3 MLP 784-1024-512-256-128-10 (P=1494154), loader batch 256', 'cov_type': cov,
                        'accum_examples_per_s': N / acc_ms * 1e3, 'glm_input_samples_per_s': nte * S / glm_ms * 1e3,
                        'glm_inputs': nte, 'glm_samples': S, 'roofline': roof})
        del model, lap, batches, Xte
        torch.manual_seed(71)
        model = CIFAR10Net(3, 10, use_tanh=True).to(dev)
        N = 2048
        X = torch.randn(N, 3, 32, 32, generator=g)
        y = torch.randint(10, (N,), generator=g)
        batches = [(X[i:i + 256].to(dev), y[i:i + 256].to(dev)) for i in range(0, N, 256)]
        lap = Laplace(model, 1.0, CategoricalLh())
        lap.infer(batches[:2], cov_type='kron', factorise=False)
        acc_ms = timed(lambda: lap.infer(batches, cov_type='kron', factorise=False))
        lap.infer(batches[:2], cov_type='diag', factorise=False)
        dacc_ms = timed(lambda: lap.infer(batches, cov_type='diag', factorise=False))
        pkk, _ = peaks()
        tc_peak = pkk.get('bf16_tflops_sustained', pkk['bf16_tflops']) / 3.0
        # SURVEY 8d, diag: per-example conv flops 2 K L m n (conv1 64x75 L=784, conv2 96x576 L=144, conv3 128x864 L=36) + dense
        d4_flop = 2.0 * 10 * (784 * 64 * 75 + 144 * 96 * 576 + 36 * 128 * 864) + 2.0 * (1152 * 512 + 512 * 256 + 256 * 10)
        d4_tf = d4_flop * N / (dacc_ms * 1e-3) / 1e12
        out.append({'config': 'cfg4 CIFAR10Net(3,10) (P=895210), loader batch 256', 'cov_type': 'diag',
                    'accum_examples_per_s': N / dacc_ms * 1e3,
                    'roofline': {'accum': {'bound': 'tensor', 'achieved': d4_tf, 'peak': tc_peak, 'unit': 'TFLOP/s fp32-equivalent',
                                           'frac': d4_tf / tc_peak, 'flop_per_example': d4_flop,
                                           'note': 'per-row weight-Jacobian products only (2 K L m n); the forward / K-column '
                                                   'backward passes that feed them are not counted'}}})
        c4_tf = 0.53e9 * N / (acc_ms * 1e-3) / 1e12                                 # ~0.53 GFLOP per example
        out.append({'config': 'cfg4 CIFAR10Net(3,10) (P=895210), loader batch 256', 'cov_type': 'kron',
                    'accum_examples_per_s': N / acc_ms * 1e3,
                    'roofline': {'accum': {'bound': 'tensor', 'achieved': c4_tf, 'peak': tc_peak, 'unit': 'TFLOP/s fp32-equivalent',
                                           'frac': c4_tf / tc_peak, 'flop_per_example': 0.53e9},
                                 'input_stream_GBs': 3 * 32 * 32 * 4.0 * N / (acc_ms * 1e-3) / 1e9}})
        del model, lap, batches
        torch.cuda.empty_cache()
        # function-space inference (experiments/imginference.py:343-372: subset of M training inputs against
        # the validation / test inputs) on the cfg5 network: NTK Grams per output + the GP predictive
        from preds_b200 import FunctionaLaplace
        torch.manual_seed(71)
        model = MLPS(784, [75], 10, 'tanh', flatten=True).to(dev)
        M, nte = 1600, 10000
        Xtr = torch.randn(M, 1, 28, 28, generator=g)
        Xte = torch.randn(nte, 1, 28, 28, generator=g)
        lap = FunctionaLaplace(model, 1.0 * 60000 / M, CategoricalLh())
        lap.infer([(Xtr[:64], None)], [(Xte[:64], None)])
        inf_ms = timed(lambda: lap.infer([(Xtr, None)], [(Xte, None)]), reps=2)
        lap.predictive_samples(n_samples=8, batch_size=8, seed=1)
        pred_ms = timed(lambda: lap.predictive_samples(n_samples=1000, batch_size=1000, seed=2), reps=2)
        P = lap.P
        from preds_b200 import gps as _gps
        _gps.STRUCTURED = False
        lap.infer([(Xtr[:64], None)], [(Xte[:64], None)])
        mat_ms = timed(lambda: lap.infer([(Xtr, None)], [(Xte, None)]), reps=2)
        _gps.STRUCTURED = True
        pairs = float(M * M + M * nte)
        out.append({'config': 'function-space Laplace, MLPS(784,[75],10) P=59635, M=1600 training x 10000 test inputs, 10 outputs',
                    'infer_ms': inf_ms, 'kernel_entries_per_s': 10.0 * (pairs + nte) / inf_ms * 1e3,
                    'as_written_tflops_equivalent': 10.0 * pairs * P * 2.0 / (inf_ms * 1e-3) / 1e12,
                    'structured_macs_per_entry': (785 + 76) / 10.0 + 75 + 10,
                    'infer_ms_materialised_jacobians': mat_ms,
                    'materialised_bf16_tflops': 10.0 * pairs * P * 6.0 / (mat_ms * 1e-3) / 1e12,
                    'predictive_ms': pred_ms, 'predictive_input_samples_per_s': nte * 1000 / pred_ms * 1e3,
                    'note': 'infer = kernels K_train / K_train_test / K_test + f_gp for all outputs.  Linear-only networks: one '
                            'forward + one K-column backward per chunk, per layer A~ A~^T once and (G_c G_c^T) o (A~ A~^T) per '
                            'output on the TMA-fed tensor-core GEMM (Hadamard in the epilogue), J never formed; '
                            'infer_ms_materialised_jacobians = the general route (conv nets): one-column Jacobians + J J^T, '
                            'P walked in 4096-wide segments, 3 bf16 MMAs per product.  predictive = GP moments '
                            '(preds/gps.py:411-426) + 1000-sample Monte-Carlo mean'})
    except Exception as exc:       # secondary numbers must not cost the headline
        out.append({'error': repr(exc)[:200]})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=6)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=16384)
    ap.add_argument('--engine', type=int, default=2)
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--cpu-sample', type=int, default=24)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-glm', action='store_true')
    ap.add_argument('--no-other-configs', action='store_true')
    ap.add_argument('--no-strong', action='store_true')
    ap.add_argument('--glm-inputs', type=int, default=10000)
    ap.add_argument('--glm-samples', type=int, default=1000)
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
