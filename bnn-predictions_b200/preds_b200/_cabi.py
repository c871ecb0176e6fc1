"""ctypes binding of libbnnp_b200.so (declared in include/bnnp_b200.h).

There is NO fallback: if the shared library is missing this module raises at import, and every
compute entry point raises if the CUDA call fails.  PyTorch is only used for device memory and
streams: pointers and the current stream are handed to the C ABI.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'lib', 'libbnnp_b200.so')

MAX_OPS = 64
MAX_K = 32            # BNNP_MAX_K of include/bnnp_b200.h (register-resident K x K code)
MAX_K_WIDE = 128      # BNNP_MAX_K_WIDE: likelihood terms / GLM sampler beyond that on warp- / block-per-example kernels
OP_LINEAR, OP_CONV2D, OP_RELU, OP_TANH, OP_MAXPOOL2D, OP_AVGPOOL2D, OP_FLATTEN, OP_SIGMOID = range(1, 9)
LIK_CATEGORICAL, LIK_GAUSSIAN, LIK_BERNOULLI = 0, 1, 2
BWD_FIRST_CONV_POSSUM = 1
BWD_FIRST_CONV_PLANES = 2
ERR_INVALID = -1


class Op(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        'kind', 'in_c', 'in_h', 'in_w', 'out_c', 'out_h', 'out_w', 'kh', 'kw', 'stride_h',
        'stride_w', 'pad_l', 'pad_r', 'pad_t', 'pad_b', 'has_bias')] + [
        ('weight', C.c_void_p), ('bias', C.c_void_p),
        ('theta_off_w', C.c_int64), ('theta_off_b', C.c_int64)]


class OpPlan(C.Structure):
    _fields_ = [('act_off', C.c_int64), ('act_ld', C.c_int64), ('grad_off', C.c_int64),
                ('grad_ld', C.c_int64), ('idx_off', C.c_int64), ('lin_col', C.c_int32),
                ('lin_unit', C.c_int32)]


class Plan(C.Structure):
    _fields_ = [('total_floats', C.c_int64), ('acat_off', C.c_int64), ('gcat_off', C.c_int64),
                ('Dtot', C.c_int32), ('Mtot', C.c_int32), ('n_linear', C.c_int32),
                ('first_linear', C.c_int32), ('P', C.c_int64), ('K', C.c_int32),
                ('reserved', C.c_int32), ('conv_scratch_off', C.c_int64), ('op', OpPlan * MAX_OPS)]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        'libbnnp_b200.so not found at %s -- build it with `python bnn-predictions_b200/build.py` '
        '(there is no CPU or PyTorch fallback for the Laplace-GGN path)' % LIB_PATH)
lib = C.CDLL(LIB_PATH)

_P, _I, _L, _F, _V = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_void_p
_OPS, _PLAN = C.POINTER(Op), C.POINTER(Plan)

_SIGS = {
    'bnnp_strerror': (C.c_char_p, [_I]),
    'bnnp_version': (_I, []),
    'bnnp_device_ok': (_I, []),
    'bnnp_net_plan': (_I, [_OPS, _I, _L, _I, _PLAN]),
    'bnnp_net_forward': (_I, [_OPS, _I, _PLAN, _P, _L, _P, _P, _V]),
    'bnnp_net_backward': (_I, [_OPS, _I, _PLAN, _P, _L, _I, _P, _V]),
    'bnnp_net_backward_ex': (_I, [_OPS, _I, _PLAN, _P, _L, _I, _P, _I, _V]),
    'bnnp_net_first_conv_possum_ok': (_I, [_OPS, _I]),
    'bnnp_lik_hessian': (_I, [_I, _P, _L, _I, _F, _P, _V]),
    'bnnp_lik_sqrt_seed': (_I, [_I, _P, _L, _I, _P, _V]),
    'bnnp_identity_seed': (_I, [_L, _I, _P, _V]),
    'bnnp_lik_residual': (_I, [_I, _P, _P, _L, _I, _F, _P, _V]),
    'bnnp_lik_inv_link': (_I, [_I, _P, _L, _I, _P, _V]),
    'bnnp_jacobians': (_I, [_OPS, _I, _PLAN, _L, _P, _P, _P, _V]),
    'bnnp_jacobians_cols': (_I, [_OPS, _I, _PLAN, _L, _I, _P, _P, _P, _V]),
    'bnnp_ggn_full_scratch_floats': (_L, [_PLAN, _L, _I]),
    'bnnp_ggn_full_accum': (_I, [_OPS, _I, _PLAN, _L, _P, _P, _P, _P, _I, _V]),
    'bnnp_ggn_full_row_chunks': (_I, [_OPS, _I, _PLAN, _L, _I, _I, C.POINTER(C.c_int64)]),
    'bnnp_ggn_full_accum_rows': (_I, [_OPS, _I, _PLAN, _L, _P, _P, _P, _P, _I, _L, _L, _I, _V]),
    'bnnp_zero_lower': (_I, [_P, _L, _V]),
    'bnnp_symmetrize_lower': (_I, [_P, _L, _V]),
    'bnnp_symmetrize_lower_rows': (_I, [_P, _L, _L, _L, _V]),
    'bnnp_exchange_lower_allreduce': (_I, [C.POINTER(_P), _P, _I, _I, _L, _L, _L, _P, _F, _I, _V]),
    'bnnp_packed_lower_floats': (_L, [_L, _L]),
    'bnnp_pack_lower': (_I, [_P, _L, _L, _L, _P, _V]),
    'bnnp_unpack_lower': (_I, [_P, _L, _L, _L, _P, _P, _F, _V]),
    'bnnp_ggn_diag_accum': (_I, [_OPS, _I, _PLAN, _L, _I, _P, _P, _F, _P, _P, _V]),
    'bnnp_ggn_diag_accum_ex': (_I, [_OPS, _I, _PLAN, _L, _I, _P, _P, _F, _P, _P, _I, _V]),
    'bnnp_net_first_conv_planes_ok': (_I, [_OPS, _I, _L, _I]),
    'bnnp_ggn_diag_scratch_floats': (_L, [_OPS, _I, _PLAN, _L, _I]),
    'bnnp_kfac_accum': (_I, [_OPS, _I, _PLAN, _L, _I, _P, _P, _F, _F, C.POINTER(_P), _P, _V]),
    'bnnp_kfac_accum_ex': (_I, [_OPS, _I, _PLAN, _L, _I, _P, _P, _F, _F, C.POINTER(_P), _P, _I, _V]),
    'bnnp_kfac_scratch_floats': (_L, [_OPS, _I, _PLAN, _L, _I]),
    'bnnp_glm_fmu': (_I, [_OPS, _I, _PLAN, _L, _P, _P, _P, _P, _P, _V]),
    'bnnp_glm_fvar_full_scratch_floats': (_L, [_PLAN, _L, _I]),
    'bnnp_glm_fvar_full': (_I, [_OPS, _I, _PLAN, _L, _P, _P, _P, _P, _I, _V]),
    'bnnp_split_bf16': (_I, [_P, _L, _L, _L, _P, _P, _V]),
    'bnnp_glm_fvar_full_tma_scratch_floats': (_L, [_OPS, _I, _PLAN, _L]),
    'bnnp_glm_fvar_full_tma': (_I, [_OPS, _I, _PLAN, _L, _P, _P, _P, _P, _P, _P, _V]),
    'bnnp_glm_fvar_diag': (_I, [_OPS, _I, _PLAN, _L, _P, _P, _P, _P, _P, _V]),
    'bnnp_glm_fvar_diag_scratch_floats': (_L, [_OPS, _I, _PLAN, _L]),
    'bnnp_glm_fvar_kron': (_I, [_OPS, _I, _PLAN, _L, _P, _P, C.POINTER(_P), C.POINTER(_F),
                                C.POINTER(_F), _I, _P, _P, _V]),
    'bnnp_glm_fvar_kron_scratch_floats': (_L, [_OPS, _I, _PLAN, _L]),
    'bnnp_kron_fvar_block': (_I, [_P, _L, _I, _I, _I, _P, _P, _P, _P, _F, _I, _P, _P, _L, _V]),
    'bnnp_glm_sample': (_I, [_I, _P, _P, _P, C.c_uint64, _L, _L, _I, _P, _P, _V]),
    'bnnp_glm_sample_diag_fallback': (_I, [_I, _P, _P, _P, C.c_uint64, _L, _L, _I, _P, _V]),
    'bnnp_sample_theta_full': (_I, [_P, _P, _P, _L, _L, _P, _V]),
    'bnnp_sample_theta_diag': (_I, [_P, _P, _P, _L, _L, _P, _V]),
    'bnnp_kron_sample_group': (_I, [_P, _P, _P, _P, _I, _I, _F, _I, _P, _L, _P, _P, _L, _L, _P, _V]),
    'bnnp_bnn_forward_scratch_floats': (_L, [_OPS, _I, _L, _L]),
    'bnnp_bnn_forward': (_I, [_OPS, _I, _I, _P, _L, _P, _L, _L, _P, _P, _V]),
    'bnnp_jtr_scratch_floats': (_L, [_PLAN, _L]),
    'bnnp_jtr_accum': (_I, [_OPS, _I, _PLAN, _L, _P, _P, _P, _P, _V]),
    'bnnp_ggn_diag_lambda': (_I, [_OPS, _I, _PLAN, _L, _P, _P, _P, _P, _V]),
    'bnnp_glm_linear_samples_scratch_floats': (_L, [_PLAN, _L, _L]),
    'bnnp_glm_linear_samples': (_I, [_OPS, _I, _PLAN, _L, _P, _P, _P, _L, _P, _P, _V]),
    'bnnp_glm_sample_mean': (_I, [_I, _P, _P, _P, C.c_uint64, _L, _L, _I, _P, _P, _V]),
    'bnnp_lik_link_moments': (_I, [_I, _P, _P, _P, _F, _L, _I, _P, _P, _P, _V]),
    'bnnp_metric_cls': (_I, [_I, _P, _P, _L, _I, _P, _I, _P, _V]),
    'bnnp_sgemm': (_I, [_I, _I, _L, _L, _L, _F, _P, _L, _P, _L, _F, _P, _L, _V]),
    'bnnp_axpy': (_I, [_L, _F, _P, _P, _V]),
    'bnnp_gemm_nt_scratch_floats': (_L, [_L, _L, _L]),
    'bnnp_gemm_nt_split': (_I, [_L, _L, _L, _F, _P, _P, _L, _P, _P, _L, _F, _P, _L, _P, _V]),
    'bnnp_split_bf16_t': (_I, [_P, _L, _L, _L, _L, _P, _P, _V]),
    'bnnp_split_bf16_rows': (_I, [_P, _L, _L, _L, _L, _P, _P, _V]),
    'bnnp_gemm_nt_split_had': (_I, [_L, _L, _L, _F, _P, _P, _L, _P, _P, _L, _F, _P, _L, _P, _L, _V]),
    'bnnp_split3_bf16': (_I, [_P, _L, _L, _L, _P, _L, _L, _I, _F, _V]),
    'bnnp_gemm3_nt': (_I, [_L, _L, _L, _F, _P, _L, _L, _P, _L, _L, _F, _P, _L, _I, _I, _L, _V]),
    'bnnp_gemm3_nt_owned': (_I, [_L, _L, _L, _F, _P, _L, _L, _P, _L, _L, _F, _P, _L, _I, _I, _L, _I, _I, _I, _L, _V]),
    'bnnp_sgemm_tc_scratch_floats': (_L, [_L, _L, _L]),
    'bnnp_sgemm_tc': (_I, [_I, _I, _L, _L, _L, _F, _P, _L, _P, _L, _F, _P, _L, _P, _V]),
    'bnnp_launch_count': (C.c_uint64, []),
    'bnnp_set_option': (_I, [C.c_char_p, _I]),
    'bnnp_get_option': (_I, [C.c_char_p]),
    'bnnp_prof_enable': (_I, [_I]),
    'bnnp_prof_read': (_I, [C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_double), _I]),
    'bnnp_glm_prof_read': (_I, [C.POINTER(C.c_float), C.POINTER(C.c_double), _I]),
}
EXPORTED = sorted(_SIGS)
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)          # AttributeError here = header/library mismatch
    _fn.restype, _fn.argtypes = _res, _args


class BnnpError(RuntimeError):
    pass


def check(code, what=''):
    if code == 0:
        return
    msg = '%s: %s' % (what or 'libbnnp_b200', lib.bnnp_strerror(code).decode())
    if code == ERR_INVALID:
        raise ValueError(msg)
    raise BnnpError(msg)


def require_device(t=None):
    if not torch.cuda.is_available():
        raise BnnpError('preds_b200 needs a CUDA device: the Laplace-GGN path has no CPU fallback')
    if t is not None and not t.is_cuda:
        raise BnnpError('preds_b200 expects the model / tensors on a CUDA device (got %s)' % t.device)


def ptr(t):
    if t is None:
        return None
    assert t.is_contiguous(), 'non-contiguous tensor handed to the C ABI'
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def f32(t, device):
    return t.to(device=device, dtype=torch.float32).contiguous()
