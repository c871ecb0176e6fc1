"""Build libbnnp_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'lib')
SO = os.path.join(LIB, 'libbnnp_b200.so')
SOURCES = ['net.cu', 'lik.cu', 'ggn.cu', 'gram_tc.cu', 'gemm_tc.cu', 'glm.cu', 'glm_tc.cu', 'bnn.cu', 'linpred.cu', 'metrics.cu',
           'exchange.cu', 'wide_k.cu', 'gemm_tma.cu', 'gemm3.cu', 'conv_diag_tc.cu', 'conv_bwd_tc.cu']
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Wno-deprecated-gpu-targets']
# BNNP_EXPERIMENTS=1 at build time also compiles the kernels kept only for A/B timing (never part of the default build)
if os.environ.get('BNNP_EXPERIMENTS') == '1':
    FLAGS.append('-DBNNP_EXPERIMENTS')


def _stale(obj, src):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    deps.append(os.path.join(HERE, '..', 'include', 'bnnp_b200.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    os.makedirs(LIB, exist_ok=True)
    objs, procs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(LIB, s.replace('.cu', '.o'))
        objs.append(obj)
        if force or _stale(obj, src):
            cmd = ['nvcc'] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for s, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0 or verbose:
            print('== %s ==\n%s' % (s, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    if procs or not os.path.exists(SO):
        subprocess.check_call(['nvcc', '-shared', '-Wno-deprecated-gpu-targets', '-o', SO] + objs + ['-lcudart'])
    return SO


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv, force='-f' in sys.argv))
