"""Config 4 (BASELINE.json): CIFAR10Net(3,10) KFAC accumulation sharded over the ranks of a torchrun launch
(one process per GPU, loader batches of 256 dealt round-robin, ONE all-reduce of the factor bucket per infer).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/cfg4_kron_dist.py
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
import torch.distributed as dist
from preds_b200 import Laplace, CategoricalLh, CIFAR10Net

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
if world > 1:
    dist.init_process_group('nccl')
torch.manual_seed(71)
model = CIFAR10Net(3, 10, use_tanh=True).cuda()
per_rank = int(sys.argv[1]) if len(sys.argv) > 1 else 4096          # examples per rank (weak scaling)
bs = 256
g = torch.Generator().manual_seed(15 + rank)
X = torch.randn(per_rank, 3, 32, 32, generator=g).cuda(); y = torch.randint(10, (per_rank,), generator=g).cuda()
batches = [(X[i:i + bs], y[i:i + bs]) for i in range(0, per_rank, bs)]
lap = Laplace(model, 1.0, CategoricalLh())
best = None
for it in range(4):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lap.infer(batches, cov_type='kron', shard='presharded' if world > 1 else None, factorise=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t)
    if it > 0:
        best = ms if best is None else min(best, ms)
if rank == 0:
    payload = sum(f.numel() for qh in lap.Sigma_chol.qhs for f in qh)
    print(json.dumps({'cfg': 'cfg4 CIFAR10Net(3,10) kron', 'n_gpus': world, 'examples_per_gpu': per_rank, 'ms': round(best, 2),
                      'examples_per_s': round(per_rank * world / best * 1e3), 'allreduce_floats': payload}))
if world > 1:
    dist.destroy_process_group()
