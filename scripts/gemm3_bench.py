"""Rate of bnnp_gemm3_nt (three-term split, six MMAs per product) on the shapes of the factorisation chain."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bnn-predictions_b200'))
import torch
from preds_b200.factor import _Planes, _gemm3, _tri_inv

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

which = sys.argv[1] if len(sys.argv) > 1 else 'all'
m = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
K = 1024
pa = _Planes(m, K, 'cuda', zero=True)
pa.buf.copy_(torch.randn(pa.buf.numel(), device='cuda').to(torch.bfloat16))
C = torch.zeros(m, m, device='cuda')
if which in ('all', 'syrk'):
    ms = timeit(lambda: _gemm3(m, m, K, -1.0, pa, pa.addr(), pa, pa.addr(), 1.0, C.data_ptr(), m, lower=1))
    macs = (m / 128) * (m / 128 + 1) / 2 * 128 * 128 * K
    print('syrk lower m %d K %d: %.3f ms  %.0f TFLOP/s bf16 (6 passes)  %.0f TFLOP/s fp32-equivalent' % (m, K, ms, macs * 12 / ms / 1e9, macs * 2 / ms / 1e9))
if which in ('all', 'full'):
    ms = timeit(lambda: _gemm3(m, m, K, 1.0, pa, pa.addr(), pa, pa.addr(), 0.0, C.data_ptr(), m))
    macs = float(m) * m * K
    print('full m %d K %d: %.3f ms  %.0f TFLOP/s bf16' % (m, K, ms, macs * 12 / ms / 1e9))
if which in ('all', 'panel'):
    pw = _Planes(K, K, 'cuda', zero=True)
    pw.buf.copy_(torch.randn(pw.buf.numel(), device='cuda').to(torch.bfloat16))
    ms = timeit(lambda: _gemm3(m, K, K, 1.0, pa, pa.addr(), pw, pw.addr(), 0.0, C.data_ptr(), m))
    print('panel m %d n %d K %d: %.3f ms  %.0f TFLOP/s bf16' % (m, K, K, ms, float(m) * K * K * 12 / ms / 1e9))
    src = torch.randn(m, K, device='cuda')
    ms = timeit(lambda: pa.fill(src.data_ptr(), K, m, K))
    print('split3 of [%d, %d]: %.3f ms  (%.0f GB/s)' % (m, K, ms, m * K * 10 / ms / 1e6))
if which in ('all', 'diag'):
    g = torch.randn(K, 2 * K, device='cuda'); D = g @ g.T
    ms = timeit(lambda: torch.linalg.cholesky_ex(D), 10)
    L = torch.linalg.cholesky(D)
    ms2 = timeit(lambda: _tri_inv(L), 10)
    print('diag block %d: cholesky_ex %.3f ms, triangular inverse %.3f ms' % (K, ms, ms2))
if which in ('sigma',):
    # one K segment of Sigma = Y Y^T: operands are 4096-wide column slices of [m, m] planes (row pitch 2 m bytes)
    py = _Planes(m, m, 'cuda', zero=True)
    py.buf.copy_(torch.randn(py.buf.numel(), device='cuda').to(torch.bfloat16))
    k0, kk = m - 8192, 4096
    rows = k0 + kk
    ms = timeit(lambda: _gemm3(rows, rows, kk, 1.0, py, py.addr(0, k0), py, py.addr(0, k0), 1.0, C.data_ptr(), m, lower=1, ktri=2, koff=-k0))
    nt = (rows + 127) // 128
    macs = sum((mt + 1) * 128 * 128 * (kk - max(0, mt * 128 - k0) // 32 * 32) for mt in range(nt))
    print('sigma segment rows %d K %d: %.3f ms  %.0f TFLOP/s bf16' % (rows, kk, ms, macs * 12 / ms / 1e9))
