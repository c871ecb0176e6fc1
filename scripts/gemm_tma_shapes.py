"""Rate of bnnp_gemm_nt_split (TMA-fed, 3 bf16 MMAs per product) on the product shapes of the cfg3 / cfg4 factor paths."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200._cabi import lib, check, ptr, stream

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

shapes = [('cfg3 fwd L1', 4096, 1024, 784), ('cfg3 bwd L2', 40960, 1024, 512), ('cfg3 bwd L3', 40960, 512, 256),
          ('cfg3 bwd L4', 40960, 256, 128), ('cfg3 G-Gram L1', 1024, 1024, 40960), ('cfg3 G-Gram L2', 512, 512, 40960),
          ('cfg3 A-Gram L1', 784, 784, 4096), ('cfg3 A-Gram L2', 1024, 1024, 4096), ('cfg4 A-Gram conv2', 576, 576, 73728),
          ('cfg4 A-Gram conv3', 864, 864, 18432), ('big square', 8192, 8192, 4096)]
for name, M, N, K in shapes:
    Kp = (K + 63) // 64 * 64
    ah = torch.randn(M, Kp, device='cuda').to(torch.bfloat16); al = (0.004 * torch.randn(M, Kp, device='cuda')).to(torch.bfloat16)
    bh = torch.randn(N, Kp, device='cuda').to(torch.bfloat16); bl = (0.004 * torch.randn(N, Kp, device='cuda')).to(torch.bfloat16)
    C = torch.zeros(M, N, device='cuda')
    sc = torch.empty(int(lib.bnnp_gemm_nt_scratch_floats(M, N, K)), device='cuda')
    ms = timeit(lambda: check(lib.bnnp_gemm_nt_split(M, N, K, 1.0, ptr(ah), ptr(al), Kp, ptr(bh), ptr(bl), Kp, 0.0, ptr(C), N, ptr(sc), stream())))
    print('%-20s M %6d N %5d K %6d: %7.3f ms  %6.0f TFLOP/s bf16' % (name, M, N, K, ms, 6.0 * M * N * K / ms / 1e9), flush=True)
