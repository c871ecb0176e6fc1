"""A KFAC-shaped Gram on the generic tensor-core GEMM: C = A^T A, A [40960, 1024] (cfg3 G-factor of the
1024-wide layer at 4096 examples x 10 columns), split-K, vs torch fp32 matmul for reference timing."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200._cabi import lib, check, ptr, stream
R, m = 40960, 1024
A = torch.randn(R, m, device='cuda')
C = torch.zeros(m, m, device='cuda')
sc = torch.empty(int(lib.bnnp_sgemm_tc_scratch_floats(m, m, R)), device='cuda')
for it in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    check(lib.bnnp_sgemm_tc(1, 0, m, m, R, 1.0, ptr(A), m, ptr(A), m, 0.0, ptr(C), m, ptr(sc), stream()))
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
ref = A.double().T @ A.double()
print('gemm_tc Gram %dx%dx%d: %.3f ms, %.1f TFLOP/s fp32-equivalent (x3 executed bf16), rel err %.2e'
      % (m, m, R, ms, 2.0 * m * m * R / ms / 1e9, float((C.double() - ref).abs().max() / ref.abs().max())))
torch.backends.cuda.matmul.allow_tf32 = False
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
D = A.T @ A; e0.record(); D = A.T @ A; e1.record(); torch.cuda.synchronize()
print('cuBLAS fp32 (no TF32) same Gram: %.3f ms' % e0.elapsed_time(e1))
