"""``Laplace`` with the public surface of ``preds/laplace.py:18-135``:

    Laplace(model, prior_prec, likelihood).infer(loader, cov_type='full'|'kron'|'diag')
    .predictive_samples_glm(X, n_samples) / .predictive_samples_bnn(X, n_samples) / .Sigma

Everything between the loader and the posterior runs in the CUDA library: one forward and one
K-column backward per batch give the layer factors (a_l, g_l); the GGN / KFAC / diagonal
accumulators contract them without materialising a Jacobian.  Additive extensions (not in the
reference): ``eps=`` / ``seed=`` on the predictives to inject the normal draws, ``shard=`` on
``infer`` for the one-process-per-GPU data-parallel mode, and ``timings`` (CUDA-event ms split
into accumulation and the one-off factorisation).
"""
import ctypes as C

import numpy as np
import torch
from torch.nn.utils import parameters_to_vector

from . import _cabi as cabi
from ._cabi import lib, check, ptr, stream
from ._engine import net_for, ptr_array
from .kron import Kron
from .predictives import functional_sampling_predictive, nn_sampling_predictive, functional_mean_predictive
from . import parallel
from . import dense
from . import factor
from . import gps
from .likelihoods import CategoricalLh

# Full GGN: loader batches are concatenated into super-batches of up to this many examples per
# accumulation call (bounded by FULL_SCRATCH_FLOATS of kernel scratch).  Inside the library one
# accumulator never sums more than 4096 examples (hierarchical fp32 summation through H).
FULL_SUPERBATCH = 16384
FULL_SCRATCH_FLOATS = 1 << 30
# kron / diag: loader batches are likewise concatenated (G sums over examples and A is weighted by
# M/N, so the factors do not depend on how the examples are batched) up to this many examples and
# this much layer-factor workspace.
FACTOR_SUPERBATCH = 4096
FACTOR_WORKSPACE_FLOATS = 1 << 30


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


class Laplace:

    def __init__(self, model, prior_prec, likelihood):
        self.model = model
        self.prior_prec = prior_prec
        self.likelihood = likelihood
        self.mu = None
        self.Sigma_chol = None
        self.device = next(model.parameters()).device
        self.inferred = False
        self.P = len(parameters_to_vector(model.parameters()))
        self._Sigma = None
        self.precision = None
        self.timings = {}
        self.exchange_stats = None   # parallel.LowerExchange of the last sharded full-covariance infer
        self._symmetric = None       # (accumulator, handle) when the precision lives in symmetric memory
        self.ggn_engine = 0          # 0 auto, 1 SIMT fp32, 2 tcgen05 split-bf16

    # ------------------------------------------------------------------------------ infer
    def infer(self, train_loader, cov_type='full', dampen_kron=False, shard=None, factorise=True, exchange='auto'):
        """Compute the posterior (preds/laplace.py:30-82).

        shard: None   -- single process, exactly the reference's semantics;
               'batches'    -- under torch.distributed, rank r takes loader batches r, r+W, ...;
               'presharded' -- the loader already yields this rank's examples only.
        In both sharded modes the accumulated factors are summed with ONE exchange step and the
        prior is added once, so every rank ends with the same posterior.  For the full covariance the exchange is
        the library's own kernel over NVLink peer memory (lower triangle only, row chunks overlapped with the
        accumulation of the last super-batch; ``exchange`` = 'auto' | 'nvls' | 'p2p' | 'nccl', see
        ``parallel.LowerExchange``); kron / diag exchange a few MB through NCCL.

        factorise=False stops after the accumulation (``self.precision`` / the ``Kron`` factors
        are set, ``Sigma_chol`` is not): the one-off factorisation is timed separately.
        """
        cabi.require_device(next(self.model.parameters()))
        self.mu = parameters_to_vector(self.model.parameters()).detach()
        self.model.eval()
        self._Sigma = None
        self.exchange_stats = None
        self._symmetric = None
        dist = _dist() if shard else None
        if shard and shard not in ('batches', 'presharded'):
            raise ValueError("shard must be None, 'batches' or 'presharded'")
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)

        def my_batches():
            for i, (X, y) in enumerate(train_loader):
                if dist and not parallel.owns_batch(i, shard, rank, world):
                    continue
                yield X, y

        t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0.record()
        if cov_type == 'full':
            # [P,P] fp32 accumulator (+ the factor, the inverse, one temporary while factorising, and the bf16 hi/lo copies
            # of Sigma the GLM predictive keeps): say so before the allocator does, the reference would simply run out of
            # memory inside einsum / cholesky
            need = 4 * self.P * self.P * (5 if factorise else 1)
            free = torch.cuda.mem_get_info(self.device)[0] + torch.cuda.memory_reserved(self.device) \
                - torch.cuda.memory_allocated(self.device)
            if need > free:
                raise cabi.BnnpError("cov_type='full' with P = %d needs about %.1f GB of device memory (%.1f GB available); "
                                     "use cov_type='kron' or 'diag'" % (self.P, need / 1e9, free / 1e9))
            # a scalar / [P] prior only ever touches the diagonal: it is added there in place (the reference builds a dense
            # [P,P] diag matrix, preds/laplace.py:115-135 -- 14 GB at config 5 to hold P numbers)
            prior = self.prior_prec
            dense_prior = torch.is_tensor(prior) and prior.ndim == 2
            if torch.is_tensor(prior) and prior.ndim > 2:
                raise ValueError('Invalid shape for prior precision')
            if torch.is_tensor(prior) and prior.ndim == 0:
                prior = float(prior)
            diag_prior = None if dense_prior else prior
            ex = None
            if dist:
                acc, hdl = (None, None) if exchange == 'nccl' else parallel.symmetric_zeros(self.P, self.device)
                self._symmetric = (acc, hdl) if acc is not None else None
                if acc is None:
                    acc = torch.empty(self.P, self.P, device=self.device)
                    check(lib.bnnp_zero_lower(ptr(acc), self.P, stream()), 'bnnp_zero_lower')
                ex = parallel.LowerExchange(acc, hdl, diag_prior, mode=exchange)
            elif dense_prior:           # accumulate on top of the prior; never mutate a caller-owned prior matrix
                acc = prior.to(device=self.device, dtype=torch.float32).clone().contiguous()
            else:
                # only the lower triangle is accumulated into (and zeroed); the mirror rewrites the upper one
                acc = torch.empty(self.P, self.P, device=self.device)
                check(lib.bnnp_zero_lower(ptr(acc), self.P, stream()), 'bnnp_zero_lower')
            self._accumulate_full(my_batches(), acc, exchange=ex)
            if ex is not None:
                self.exchange_stats = ex
                if dense_prior:
                    acc += prior.to(device=self.device, dtype=torch.float32)
            elif not dense_prior:
                pv = diag_prior if not torch.is_tensor(diag_prior) else diag_prior.to(device=self.device, dtype=torch.float32)
                acc.diagonal().add_(pv)
            self.precision = acc
            t1.record()
            if factorise:
                self.Sigma_chol, self._Sigma = self._factorise_full(acc, sharded=bool(dist))
        elif cov_type == 'diag':
            _, hess_factor = self.likelihood.nn_loss()
            prior = self.expand_prior_precision(cov_type)
            acc = torch.zeros(self.P, device=self.device) if dist else prior.clone().to(torch.float32)
            self._accumulate_diag(my_batches(), acc, float(hess_factor))
            if dist:
                parallel.reduce_sum_(acc, prior)
            self.precision = acc
            t1.record()
            if factorise:
                self.Sigma_chol = torch.sqrt(1 / acc)
        elif cov_type == 'kron':
            _, hess_factor = self.likelihood.nn_loss()
            assert np.isscalar(self.prior_prec)
            precision = Kron(self.model, self.prior_prec, dampen=dampen_kron)
            if type(train_loader) is list:
                n_local = sum(len(a[0]) for i, a in enumerate(train_loader)
                              if not dist or parallel.owns_batch(i, shard, rank, world))
                N = parallel.global_count(n_local, self.device, enabled=bool(dist))
            else:
                N = len(train_loader.dataset)
                N = parallel.global_count(N, self.device, enabled=bool(dist) and shard == 'presharded')
            self._accumulate_kron(my_batches(), precision, float(hess_factor), N)
            if dist:
                xt = {}
                parallel.reduce_factors_(precision.qhs, timing=xt)
                self.exchange_stats = xt
            t1.record()
            if factorise:
                precision.decompose()
            self.Sigma_chol = precision
        else:
            t1.record()
        t2.record()
        t2.synchronize()
        self.timings = {'accumulate_ms': t0.elapsed_time(t1), 'factorise_ms': t1.elapsed_time(t2)}
        net_for(self.model).release()
        self.inferred = bool(factorise)

    def release_posterior(self):
        """Drop the posterior and hand a symmetric-memory accumulator back to ``parallel`` for the next sharded
        ``infer`` of the same size (e.g. a prior-precision sweep).  The caller must not use ``precision`` afterwards.
        Collective in the sharded mode: every rank calls it."""
        if self._symmetric is not None:
            parallel.symmetric_release(*self._symmetric)
        self._symmetric = None
        self.precision = self.Sigma_chol = self._Sigma = None
        self.exchange_stats = None
        self.inferred = False

    def _accumulate_full(self, batches, H, grad=None, exchange=None):
        """H += sum_n J_n^T Lambda_n J_n over the batches (lower triangle, mirrored at the end).  With
        ``grad`` ([P]) also grad += sum_n J_n^T r_n from the same factors (LaplaceGGN.post_process,
        preds/optimizers.py:95-96); the targets are then read from the batches.

        ``exchange`` (a ``parallel.LowerExchange``, data-parallel mode): the LAST super-batch is accumulated row chunk
        by row chunk of H, and every finished chunk is handed to the exchange (own kernel over NVLink on a side stream)
        while the tensor-core kernel works on the next chunk; H is mirrored after the exchange has joined."""
        net = net_for(self.model)

        def superbatches():
            """Loader batches concatenated into super-batches of up to `limit` examples (examples per launch from the
            scratch the engine needs per example)."""
            pend, pend_y, n_pend, limit = [], [], 0, None
            for X, y in batches:
                X = X.to(self.device, non_blocking=True)
                if grad is not None:
                    y = y.to(self.device)
                if limit is None:
                    net.prepare(X[:1])
                    pl1 = net.plan(64, net.K)
                    per = max(1, int(lib.bnnp_ggn_full_scratch_floats(C.byref(pl1), 64, self.ggn_engine)) // 64)
                    limit = max(64, min(FULL_SUPERBATCH, FULL_SCRATCH_FLOATS // per))
                while len(X) > 0:
                    take = min(len(X), limit - n_pend)
                    pend.append(X[:take])
                    if grad is not None:
                        pend_y.append(y[:take])
                        y = y[take:]
                    n_pend += take
                    X = X[take:]
                    if n_pend >= limit:
                        yield (torch.cat(pend) if len(pend) > 1 else pend[0]), \
                            ((torch.cat(pend_y) if len(pend_y) > 1 else pend_y[0]) if grad is not None else None)
                        pend, pend_y, n_pend = [], [], 0
            if pend:
                yield (torch.cat(pend) if len(pend) > 1 else pend[0]), \
                    ((torch.cat(pend_y) if len(pend_y) > 1 else pend_y[0]) if grad is not None else None)

        def run(X, y, last):
            net.prepare(X[:1])
            if not net.linear_only:
                raise cabi.BnnpError("cov_type='full' is implemented for Linear-only networks "
                                     '(the reference is infeasible beyond them: [P,P] storage)')
            X2d, pl, ws, f = net.jacobian_factors(X)
            N = X2d.shape[0]
            Lam = self.likelihood._hessian_nkk(f)
            sc = net.scratch(lib.bnnp_ggn_full_scratch_floats(C.byref(pl), N, self.ggn_engine))
            if last and exchange is not None:
                bounds = (C.c_int64 * (parallel.EXCHANGE_CHUNKS + 1))()
                n = lib.bnnp_ggn_full_row_chunks(net.ops, net.n_ops, C.byref(pl), N, self.ggn_engine,
                                                 parallel.EXCHANGE_CHUNKS, bounds)
                check(min(n, 0), 'bnnp_ggn_full_row_chunks')
                if not exchange.agree(n > 1):        # some rank cannot chunk (no examples / SIMT engine): one exchange at the end
                    n, bounds[1] = 1, self.P
                for c in range(n):
                    check(lib.bnnp_ggn_full_accum_rows(net.ops, net.n_ops, C.byref(pl), N, ptr(ws), ptr(Lam), ptr(H),
                                                       ptr(sc), self.ggn_engine, bounds[c], bounds[c + 1], int(c > 0),
                                                       stream()), 'bnnp_ggn_full_accum_rows')
                    exchange.rows_done(int(bounds[c]), int(bounds[c + 1]))
            else:
                check(lib.bnnp_ggn_full_accum(net.ops, net.n_ops, C.byref(pl), N, ptr(ws), ptr(Lam), ptr(H),
                                              ptr(sc), self.ggn_engine, stream()), 'bnnp_ggn_full_accum')
            if grad is not None:
                rs = self.likelihood.residual(y, f).reshape(N, net.K).contiguous()
                sc = net.scratch(lib.bnnp_jtr_scratch_floats(C.byref(pl), N))
                check(lib.bnnp_jtr_accum(net.ops, net.n_ops, C.byref(pl), N, ptr(ws), ptr(rs), ptr(grad),
                                         ptr(sc), stream()), 'bnnp_jtr_accum')

        prev = None
        for sb in superbatches():            # one super-batch of look-ahead: the last one is exchanged chunk by chunk
            if prev is not None:
                run(prev[0], prev[1], False)
            prev = sb
        if prev is not None:
            run(prev[0], prev[1], True)
        elif exchange is not None:
            exchange.agree(False)            # a rank without examples still takes part in the exchange (H_r = 0)
            exchange.rows_done(0, self.P)
        if exchange is not None:
            exchange.finish()            # joins the side stream; the exchange has mirrored H chunk by chunk
        else:
            check(lib.bnnp_symmetrize_lower(ptr(H), self.P, stream()), 'bnnp_symmetrize_lower')

    def _factorise_full(self, H, sharded=False):
        """chol(H) -> Sigma = H^-1 -> chol(Sigma) (preds/laplace.py:43-45), one-off, on device.  Sigma itself is kept: the
        reference recomputes L L^T on every predictive call (preds/laplace.py:111).

        From factor.MIN_P parameters on, every O(P^3) product of the chain runs on the library's own tensor-core GEMM
        (factor.py / csrc/gemm3.cu: 1.75 s at P = 59 635 against 10.0 s for the three cuSOLVER calls, residual
        |H Sigma - I| 7e-5 against potri's 2.7e-4).  ``sharded`` (data-parallel mode, every rank holds the same H): the tiles
        of every product are dealt out between the ranks and gathered (factor.py), every rank ends with the bits of the
        single-process chain.  factor.ENGINE = 'library' restores the cuSOLVER path, whose inverse is split between the
        ranks when ``sharded`` (dense.spd_inverse_sharded)."""
        if factor.use_tc(H.shape[0]):
            return factor.posterior_from_precision(H, shared=sharded)
        chol = torch.linalg.cholesky(H)
        Sigma = dense.spd_inverse_from_chol(chol, sharded=sharded)
        del chol
        return torch.linalg.cholesky(Sigma), Sigma

    def _superbatches(self, batches):
        """Concatenate loader batches into super-batches bounded by FACTOR_SUPERBATCH examples and
        FACTOR_WORKSPACE_FLOATS of layer-factor workspace; yields (X, n_examples) on the device."""
        net = net_for(self.model)
        pend, n_pend, limit = [], 0, None
        for X, _ in batches:
            X = X.to(self.device)
            if limit is None:
                net.prepare(X[:1])
                per = max(1, int(net.plan(64, net.K).total_floats) // 64)
                limit = max(1, min(FACTOR_SUPERBATCH, FACTOR_WORKSPACE_FLOATS // per))
            while len(X) > 0:
                take = min(len(X), limit - n_pend)
                pend.append(X[:take])
                n_pend += take
                X = X[take:]
                if n_pend >= limit:
                    yield (torch.cat(pend) if len(pend) > 1 else pend[0]), n_pend
                    pend, n_pend = [], 0
        if pend:
            yield (torch.cat(pend) if len(pend) > 1 else pend[0]), n_pend

    def _accumulate_diag(self, batches, diag, hess_factor):
        net = net_for(self.model)
        for X, _ in self._superbatches(batches):
            X2d, pl, ws, f, flags = net.sqrt_factors(X, self.likelihood.code, diag_only=True)
            N = X2d.shape[0]
            sc = net.scratch(lib.bnnp_ggn_diag_scratch_floats(net.ops, net.n_ops, C.byref(pl), N, net.K))
            check(lib.bnnp_ggn_diag_accum_ex(net.ops, net.n_ops, C.byref(pl), N, net.K, ptr(X2d), ptr(ws),
                                             hess_factor, ptr(diag), ptr(sc), flags, stream()), 'bnnp_ggn_diag_accum_ex')

    def _accumulate_kron(self, batches, kron, hess_factor, N_total):
        net = net_for(self.model)
        index = {id(p): i for i, p in enumerate(net.params)}
        table = None
        for X, M in self._superbatches(batches):
            X2d, pl, ws, f, flags = net.sqrt_factors(X, self.likelihood.code, kfac_only=True)
            if table is None:
                ptrs = []
                for op, mod in zip(net.ops, net.mods):
                    if op.kind in (cabi.OP_LINEAR, cabi.OP_CONV2D):
                        qh = kron.qhs[index[id(mod.weight)]]
                        ptrs += [ptr(qh[0]), ptr(qh[1])]
                    else:
                        ptrs += [None, None]
                table = ptr_array(ptrs)
            sc = net.scratch(lib.bnnp_kfac_scratch_floats(net.ops, net.n_ops, C.byref(pl), M, net.K))
            check(lib.bnnp_kfac_accum_ex(net.ops, net.n_ops, C.byref(pl), M, net.K, ptr(X2d), ptr(ws), hess_factor,
                                         M / N_total, table, ptr(sc), flags, stream()), 'bnnp_kfac_accum_ex')
        # a bias shares its layer's output-side factor (BackPACK KFLR: p.kflr = [G] for biases)
        for mod in net.mods:
            if getattr(mod, 'bias', None) is not None and id(mod.bias) in index:
                kron.qhs[index[id(mod.bias)]][0].copy_(kron.qhs[index[id(mod.weight)]][0])

    # ------------------------------------------------------------------------------ predictives
    def predictive_samples_glm(self, X, n_samples=1000, eps=None, seed=None):
        if not self.inferred:
            raise ValueError('Need to infer first.')
        Sigma = self.Sigma if type(self.Sigma_chol) is not Kron else self.Sigma_chol
        return functional_sampling_predictive(X, self.model, self.likelihood, self.mu, Sigma, n_samples,
                                              eps=eps, seed=seed)

    def predictive_mean_glm(self, X, n_samples=1000, eps=None, seed=None):
        """``predictive_samples_glm(X, n_samples).mean(0)`` fused into the sampling kernel ([N,K]); additive."""
        if not self.inferred:
            raise ValueError('Need to infer first.')
        Sigma = self.Sigma_chol if isinstance(self.Sigma_chol, Kron) else self.Sigma
        return functional_mean_predictive(X, self.model, self.likelihood, self.mu, Sigma, mc_samples=n_samples,
                                          eps=eps, seed=seed)

    def predictive_samples_bnn(self, X, n_samples=1000, eps=None):
        if not self.inferred:
            raise ValueError('Need to infer first.')
        return nn_sampling_predictive(X, self.model, self.likelihood, self.mu, self.Sigma_chol, n_samples,
                                      eps=eps)

    @property
    def Sigma(self):
        """preds/laplace.py:104-113; the full covariance is cached instead of recomputed."""
        if type(self.Sigma_chol) is Kron:
            return self.Sigma_chol
        if self.Sigma_chol.ndim == 1:
            return self.Sigma_chol.square()
        elif self.Sigma_chol.ndim == 2:
            if self._Sigma is None:
                self._Sigma = self.Sigma_chol @ self.Sigma_chol.T
            return self._Sigma
        else:
            raise ValueError('yet to be implemented for kron')

    def expand_prior_precision(self, cov_type='full'):
        """Scalar / [P] / [P,P] prior -> diag matrix (full) or vector (diag); :115-135."""
        assert cov_type in ['full', 'diag']
        if np.isscalar(self.prior_prec) or self.prior_prec.ndim == 0:
            prec_diag = torch.ones(self.P, device=self.device) * self.prior_prec
            return torch.diag(prec_diag) if cov_type == 'full' else prec_diag
        elif self.prior_prec.ndim == 1:
            pp = self.prior_prec.to(self.device)
            return pp if cov_type == 'diag' else torch.diag(pp)
        elif self.prior_prec.ndim == 2:
            if cov_type == 'diag':
                raise ValueError('Will not diagonalize non-diagonal prior.')
            return self.prior_prec.to(self.device)
        else:
            raise ValueError('Invalid shape for prior precision')


class FunctionaLaplace:
    """Function-space Laplace of ``preds/laplace.py:138-207``: the GP whose kernel is the NTK ``J J^T / delta`` of the
    trained network on a subset of the training inputs, evaluated at the test inputs.  Same constructor, ``infer`` and
    ``predictive_samples`` signatures; ``gp_quantities`` holds the reference's dictionary with the tensors left on the
    model's device (the reference keeps them on the host and moves them back for every predictive call).  Categorical
    likelihoods only (the reference's Bernoulli branch hands a [m, 1] mean to a [m]-variate normal).  Additive:
    ``eps`` / ``seed`` on ``predictive_samples`` to inject or fix the standard-normal draws."""

    def __init__(self, model, prior_prec, likelihood, cache_Js=True):
        assert np.isscalar(prior_prec)
        self.model = model
        self.prior_prec = prior_prec
        self.likelihood = likelihood
        self.device = next(model.parameters()).device
        self.Kwinvs = None
        self.cache_Js = cache_Js
        self.Js = None
        self.train_loader = None
        self.log_likelihood = None
        self.inferred = False
        self.gp_quantities = None
        self.P = len(parameters_to_vector(model.parameters()))
        # gp_quantities reads these (preds/laplace.py:155-160)
        model.likelihood = likelihood
        log_var = -torch.log(torch.tensor(prior_prec))
        for p in model.parameters():
            setattr(p, 'log_prior_var', log_var)

    def infer(self, train_loadera, train_loaderb, max_ram_gb=10, batch_size_per_gpu=512, print_progress=False):
        cabi.require_device(next(self.model.parameters()))
        self.train_loader = train_loadera
        X1 = torch.cat([batch[0] for batch in train_loadera], dim=0)
        X2 = torch.cat([batch[0] for batch in train_loaderb], dim=0)
        self.gp_quantities = gps.gp_quantities_device(self.model, X1, X2, collapse_groups=True, with_prior_prec=False,
                                                      max_ram_gb=max_ram_gb)
        self.inferred = True

    def predictive_moments(self, indep=False):
        """(f_mu [m, C], f_var [m, C, C]) of the predictive at the inputs of ``train_loaderb`` (additive API)."""
        if not self.inferred or 'f_star_test' not in self.gp_quantities:
            raise ValueError('Need to infer first.')
        if not isinstance(self.likelihood, CategoricalLh):
            raise NotImplementedError('FunctionaLaplace: categorical likelihoods only')
        return gps.gp_predictive_moments(self.gp_quantities, self.prior_prec, indep=indep)

    def predictive_samples(self, n_samples=1000, batch_size=1024, indep=False, eps=None, seed=None):
        """Monte-Carlo mean of the class probabilities at the test inputs, [m, C] (preds/laplace.py:175-207 returns the
        first value of ``mc_preds``)."""
        self.model.eval()
        f_mu, f_var = self.predictive_moments(indep=indep)
        return gps.mc_mean(self.likelihood, f_mu, f_var, n_samples=n_samples, batch_size=batch_size, eps=eps, seed=seed)
