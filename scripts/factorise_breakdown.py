"""Time the three dense factorisations of Laplace.infer(cov_type='full') at P = 59 635 (cuSOLVER through torch.linalg)."""
import time, torch
P = 59635
torch.manual_seed(0)
g = torch.Generator(device='cuda').manual_seed(0)
A = torch.randn(P, 4096, device='cuda', generator=g)
H = A @ A.T
del A
H.diagonal().add_(float(H.diagonal().mean()) * 0.05)
def t(fn, name):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    print('%-28s %.2f s' % (name, time.perf_counter() - t0), flush=True); return out
L = t(lambda: torch.linalg.cholesky(H), 'cholesky(H)')
del H
S = t(lambda: torch.cholesky_inverse(L, upper=False), 'cholesky_inverse(L)')
del L
S = S.T if not S.is_contiguous() else S
C = t(lambda: torch.linalg.cholesky(S), 'cholesky(Sigma)')
