"""How do the fp32 accumulator adds of tcgen05.mma round?  Operands exactly representable in bf16 (lo term = 0), so the only
error of bnnp_gemm_nt_split against float64 is the accumulation inside the tensor core.  Prints signed mean and rms relative
error against the length of one TMEM accumulation (K / 16 dependent adds)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200._cabi import lib, check, ptr, stream

M = N = 512
for kind in ('positive', 'signed'):
    for K in (16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
        g = torch.Generator().manual_seed(K)
        if kind == 'positive':
            A = (torch.rand(M, K, generator=g) + 0.5)
            B = (torch.rand(N, K, generator=g) + 0.5)
        else:
            A = torch.randn(M, K, generator=g)
            B = torch.randn(N, K, generator=g)
        ah = A.to(torch.bfloat16).cuda(); bh = B.to(torch.bfloat16).cuda()
        al = torch.zeros_like(ah); bl = torch.zeros_like(bh)
        C = torch.zeros(M, N, device='cuda')
        check(lib.bnnp_gemm_nt_split(M, N, K, 1.0, ptr(ah), ptr(al), K, ptr(bh), ptr(bl), K, 0.0, ptr(C), N, None, stream()))
        want = ah.double() @ bh.double().T
        f32 = (ah.float() @ bh.float().T)
        scale = want.abs().mean() if kind == 'positive' else want.abs().pow(2).mean().sqrt()
        e = (C.double() - want) / (want if kind == 'positive' else scale)
        e32 = (f32.double() - want) / (want if kind == 'positive' else scale)
        print('%-8s K %5d adds %4d  tc: mean %+.3e rms %.3e max %.3e | cublas fp32: mean %+.3e rms %.3e' % (
            kind, K, K // 16, float(e.mean()), float(e.pow(2).mean().sqrt()), float(e.abs().max()),
            float(e32.mean()), float(e32.pow(2).mean().sqrt())))
