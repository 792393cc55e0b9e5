"""ctypes binding of libflowril_b200.so (the C ABI in include/flowril_b200.h).

There is NO fallback: if the library is missing or a call fails, this raises.  The library is built in-tree by
``python -m modril_b200.build`` (``__graft_entry__.build()``), never JIT-compiled into a cache directory.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libflowril_b200.so")

ABI_VERSION = 7
PREC_FP32, PREC_BF16, PREC_FP16X3, PREC_FP16 = 0, 1, 2, 3
PREC_BF16X3 = PREC_FP16X3   # former (reserved) name
HAS_TC = True  # tensor-core (tcgen05) log-density kernels are built in
PROBE_SHARED, PROBE_PER_STEP, PROBE_EXACT = 0, 1, 2
REWARD_TYPES = {"raw": 0, "norm": 1, "exp": 2, "clip": 3, "smooth": 4}
# "fp16x3": 3-term fp16 operand split on the tensor cores = fp32-grade products (the parity mode, and the default)
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "fp16": PREC_FP16, "fp16x3": PREC_FP16X3}

# every symbol include/flowril_b200.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = (
    "frl_abi_version", "frl_last_error", "frl_net_create", "frl_net_destroy", "frl_net_num_params",
    "frl_net_set_params", "frl_vfield", "frl_logprob", "frl_score", "frl_score_host", "frl_cfm_loss_grad",
    "frl_clip_adam", "frl_reg_loss_grad", "frl_grad_axpy", "frl_tensor_norms", "frl_normalise_rewards", "frl_cfm_loss_grad_p", "frl_clip_adam_dev",
    "frl_vnet_create", "frl_vnet_destroy", "frl_vnet_num_params", "frl_vnet_set_params", "frl_vnet_forward",
    "frl_vnet_sample", "frl_bcf_loss_grad", "frl_compute_returns", "frl_normalise_advantages", "frl_cfm_loss_grad_pair",
    "frl_acnet_create", "frl_acnet_destroy", "frl_acnet_num_params", "frl_acnet_set_params", "frl_acnet_forward",
    "frl_acnet_act", "frl_acnet_evaluate", "frl_ppo_loss_grad", "frl_column_moments", "frl_norm_vec",
    "frl_comm_create", "frl_comm_handle", "frl_comm_connect", "frl_comm_connect_local", "frl_comm_publish",
    "frl_comm_reduce", "frl_comm_allreduce", "frl_comm_timeouts", "frl_comm_destroy", "frl_reward", "frl_rademacher_calls",
    "frl_train_timeouts",
)


class NetConfig(C.Structure):
    _fields_ = [("s_dim", C.c_int), ("a_dim", C.c_int), ("hidden", C.c_int), ("depth", C.c_int),
                ("n_heads", C.c_int), ("device", C.c_int)]


class CfmTerm(C.Structure):
    _fields_ = [("role_sign", C.c_float), ("s", C.c_void_p), ("a0", C.c_void_p), ("t", C.c_void_p), ("noise", C.c_void_p),
                ("B", C.c_longlong), ("mean_divisor", C.c_longlong), ("grad_scale", C.c_float), ("loss", C.c_void_p)]


class VNetConfig(C.Structure):
    _fields_ = [("s_dim", C.c_int), ("a_dim", C.c_int), ("hidden", C.c_int), ("time_dim", C.c_int), ("device", C.c_int)]


class ACNetConfig(C.Structure):
    _fields_ = [("obs_dim", C.c_int), ("a_dim", C.c_int), ("hidden", C.c_int), ("layers", C.c_int), ("device", C.c_int)]


class AdamSegment(C.Structure):
    _fields_ = [("params", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_longlong)]


class NativeError(RuntimeError):
    pass


_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} not found: build it with `python -m modril_b200.build` (there is no CPU/PyTorch fallback)")
    L = C.CDLL(LIB_PATH)
    vp, ll, i32, f32, f64 = C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_double
    L.frl_abi_version.restype = i32
    L.frl_last_error.restype = C.c_char_p
    L.frl_net_create.argtypes = [C.POINTER(NetConfig), C.POINTER(vp)]
    L.frl_net_destroy.argtypes = [vp]
    L.frl_net_num_params.argtypes = [vp]
    L.frl_net_num_params.restype = ll
    L.frl_net_set_params.argtypes = [vp, vp, vp]
    L.frl_vfield.argtypes = [vp, vp, vp, vp, vp, vp, ll, i32, vp]
    L.frl_logprob.argtypes = [vp, f32, vp, vp, vp, i32, C.POINTER(f32), i32, vp, ll, i32, vp]
    L.frl_score.argtypes = [vp, vp, vp, vp, vp, i32, C.POINTER(f32), i32, i32, vp, vp, vp, ll, i32, vp]
    L.frl_score_host.argtypes = [vp, vp, vp, vp, vp, i32, C.POINTER(f32), i32, i32, vp, vp, vp, ll, i32]
    L.frl_cfm_loss_grad.argtypes = [vp, f32, vp, vp, vp, vp, ll, ll, f32, i32, vp, vp, vp]
    L.frl_cfm_loss_grad_p.argtypes = [vp, f32, vp, vp, vp, vp, ll, ll, f32, i32, vp, vp, i32, vp]
    L.frl_reg_loss_grad.argtypes = [vp, vp, vp, vp, vp, ll, ll, i32, vp, vp, vp, vp, vp]
    L.frl_grad_axpy.argtypes = [vp, vp, ll, f32, vp, vp, f32, vp, vp, i32, vp]
    L.frl_normalise_rewards.argtypes = [vp, vp, i32, i32, f32, vp, i32, vp, vp, vp, i32, vp]
    L.frl_clip_adam_dev.argtypes = [C.POINTER(AdamSegment), i32, vp, f64, f64, f64, f64, f64, vp, i32, vp]
    L.frl_tensor_norms.argtypes = [vp, C.POINTER(ll), i32, vp, i32, vp]
    L.frl_clip_adam.argtypes = [C.POINTER(AdamSegment), i32, ll, f64, f64, f64, f64, f64, vp, i32, vp]
    L.frl_vnet_create.argtypes = [C.POINTER(VNetConfig), C.POINTER(vp)]
    L.frl_vnet_destroy.argtypes = [vp]
    L.frl_vnet_num_params.argtypes = [vp]
    L.frl_vnet_num_params.restype = ll
    L.frl_vnet_set_params.argtypes = [vp, vp, vp]
    L.frl_vnet_forward.argtypes = [vp, vp, vp, vp, vp, ll, vp]
    L.frl_vnet_sample.argtypes = [vp, vp, vp, i32, vp, ll, vp]
    L.frl_bcf_loss_grad.argtypes = [vp, vp, vp, vp, vp, ll, ll, f32, f32, i32, vp, vp, vp]
    L.frl_cfm_loss_grad_pair.argtypes = [vp, C.POINTER(CfmTerm), i32, vp, i32, vp]
    L.frl_compute_returns.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, f64, f64, i32, vp, i32, vp]
    L.frl_normalise_advantages.argtypes = [vp, vp, ll, f32, vp, vp, vp, i32, vp]
    L.frl_acnet_create.argtypes = [C.POINTER(ACNetConfig), C.POINTER(vp)]
    L.frl_acnet_destroy.argtypes = [vp]
    L.frl_acnet_num_params.argtypes = [vp]
    L.frl_acnet_num_params.restype = ll
    L.frl_acnet_set_params.argtypes = [vp, vp, vp]
    L.frl_acnet_forward.argtypes = [vp, vp, ll, vp, vp, vp]
    L.frl_acnet_act.argtypes = [vp, vp, vp, ll, vp, vp, vp, vp, vp]
    L.frl_acnet_evaluate.argtypes = [vp, vp, vp, ll, vp, vp, vp, vp]
    L.frl_column_moments.argtypes = [vp, ll, i32, vp, vp, i32, vp]
    L.frl_norm_vec.argtypes = [vp, vp, vp, ll, i32, vp, i32, vp]
    L.frl_ppo_loss_grad.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ll, ll, f32, f32, f32, f32, i32, vp, vp, vp]
    L.frl_reward.argtypes = [vp, vp, vp, ll, i32, i32, vp]
    L.frl_rademacher_calls.argtypes = [C.c_ulonglong, C.c_ulonglong, i32, i32, i32, i32, vp, vp, i32, vp]
    L.frl_comm_create.argtypes = [i32, i32, i32, ll, C.POINTER(vp)]
    L.frl_comm_handle.argtypes = [vp, vp]
    L.frl_comm_connect.argtypes = [vp, vp]
    L.frl_comm_connect_local.argtypes = [vp, C.POINTER(vp)]
    for name in ("frl_comm_publish", "frl_comm_reduce", "frl_comm_allreduce"):
        getattr(L, name).argtypes = [vp, C.POINTER(vp), C.POINTER(ll), i32, vp]
    L.frl_comm_timeouts.argtypes = [vp]
    L.frl_comm_timeouts.restype = ll
    L.frl_train_timeouts.argtypes = [vp]
    L.frl_train_timeouts.restype = ll
    L.frl_comm_destroy.argtypes = [vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("frl_last_error", "frl_net_num_params", "frl_vnet_num_params", "frl_acnet_num_params",
                        "frl_comm_timeouts", "frl_train_timeouts"):
            fn.restype = i32
    if L.frl_abi_version() != ABI_VERSION:
        raise NativeError(f"ABI mismatch: library {L.frl_abi_version()} != binding {ABI_VERSION}")
    _lib = L
    return L


def check(rc: int, what: str):
    if rc == 0:
        return
    msg = lib().frl_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(f"{what}: {msg}")
    raise NativeError(f"{what} failed (status {rc}): {msg}")


def float_array(values):
    arr = (C.c_float * len(values))(*[float(v) for v in values])
    return arr
