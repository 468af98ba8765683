"""ctypes binding of libb200rec.so -- exactly the symbols include/b200rec.h declares, i.e. what a
JNA ``Native.load("b200rec")`` interface on the Scala side would bind.

There is no fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200REC_LIB") or os.path.join(HERE, "libb200rec.so")

OK, ERR_ARG, ERR_SHAPE, ERR_INDEX, ERR_CUDA, ERR_STATE, ERR_NOMEM = 0, -1, -2, -3, -4, -5, -6
OPTIMIZERS = {"sgd": 0, "momentum": 1, "adagrad": 2, "adam": 3}
KINDS = {"lr": 0, "fm": 1, "deepfm": 2, "xdeepfm": 3, "dcn": 4, "pnn": 5}

c_int_p = C.POINTER(C.c_int)
c_float_p = C.POINTER(C.c_float)
c_i64_p = C.POINTER(C.c_int64)
vp = C.c_void_p

# name -> argtypes; every function returns int except the two noted below
SIGNATURES = {
    "b200rec_abi_version": [],
    "b200rec_last_error": [],
    "b200rec_device_count": [c_int_p],
    "b200rec_launch_count": [c_i64_p],
    "b200rec_profile_begin": [],
    "b200rec_profile_end": [C.c_char_p, C.c_int64, c_i64_p],
    "b200rec_profile_overhead_us": [C.c_int, c_float_p],
    "b200rec_model_create": [C.c_int, C.c_int, C.c_int, c_int_p, C.c_int, c_int_p, C.c_int, C.c_int,
                             C.c_int, C.POINTER(vp)],
    "b200rec_model_destroy": [vp],
    "b200rec_model_mats_size": [vp, c_int_p, C.c_int, c_int_p],
    "b200rec_model_mats_len": [vp, c_i64_p],
    "b200rec_model_stream": [vp, C.POINTER(vp)],
    "b200rec_model_set_gemm_mode": [vp, C.c_int],
    "b200rec_set_default_gemm_mode": [C.c_int],
    "b200rec_model_sync": [vp],
    "b200rec_forward": [vp, C.c_int, C.c_int64, vp, vp, vp, vp, vp, vp],
    "b200rec_backward": [vp, C.c_int, C.c_int64, vp, vp, vp, vp, vp, vp, c_float_p],
    "b200rec_forward_dev": [vp, C.c_int, C.c_int64, vp, vp, vp, vp, vp, vp, vp],
    "b200rec_backward_dev": [vp, C.c_int, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp],
    "b200rec_table_create": [C.c_int64, C.c_int, C.c_int, C.POINTER(vp)],
    "b200rec_table_destroy": [vp],
    "b200rec_table_init_uniform": [vp, C.c_uint64, C.c_float, C.c_float, C.c_int64, C.c_int64],
    "b200rec_table_write": [vp, C.c_int64, C.c_int64, vp, vp],
    "b200rec_table_read": [vp, C.c_int64, C.c_int64, vp, vp],
    "b200rec_table_ptrs": [vp, C.POINTER(vp), C.POINTER(vp)],
    "b200rec_table_lookup": [vp, C.c_int64, vp, vp, vp],
    "b200rec_table_lookup_dev": [vp, C.c_int64, vp, vp, vp, vp],
    "b200rec_distinct": [C.c_int, C.c_int64, vp, vp, c_i64_p],
    "b200rec_scatter_add": [C.c_int, C.c_int, C.c_int64, vp, vp, vp, vp, vp, vp, c_i64_p],
    "b200rec_model_set_params": [vp, vp, vp],
    "b200rec_model_get_params": [vp, vp, vp],
    "b200rec_model_param_ptrs": [vp, C.POINTER(vp), C.POINTER(vp)],
    "b200rec_model_set_graph": [vp, C.c_int],
    "b200rec_step": [vp, vp, C.c_int, vp, vp, c_float_p],
    "b200rec_step_dev": [vp, vp, C.c_int, vp, vp, vp],
    "b200rec_predict": [vp, vp, C.c_int, vp, vp],
    "b200rec_predict_dev": [vp, vp, C.c_int, vp, vp, vp],
    "b200rec_step_results": [vp, c_float_p, c_i64_p, vp, vp, vp, c_float_p, vp],
    "b200rec_step_result_ptrs": [vp] + [C.POINTER(vp)] * 7,
    "b200rec_step_nnz_grad_ptrs": [vp, C.POINTER(vp), C.POINTER(vp)],
    "b200rec_model_set_fused_scatter": [vp, C.c_int, C.c_int],
    "b200rec_step_gathered_dev": [vp, C.c_int, vp, vp, vp, vp],
    "b200rec_segsum_dev": [vp, C.c_int, C.c_int64, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp],
    "b200rec_segsum_sort_dev": [vp, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, vp, vp, vp, vp],
    "b200rec_segsum_reduce_dev": [vp, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp],
    "b200rec_segsum_join_dev": [vp, C.c_int, vp],
    "b200rec_segsum_inverse_dev": [vp, C.c_int, C.c_int64, vp, vp],
    "b200rec_p2p_begin_step_dev": [vp, vp, C.c_int64, vp],
    "b200rec_p2p_allreduce_dev": [vp, C.c_int64, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp],
    "b200rec_model_side_stream": [vp, C.c_int, C.POINTER(vp)],
    "b200rec_side_fork_dev": [vp, C.c_int, vp],
    "b200rec_side_rejoin_dev": [vp, C.c_int],
    "b200rec_parse_samples": [C.c_int, C.c_char_p, C.c_int64, C.c_int64, C.c_int64, vp, vp, vp, vp, vp, c_i64_p,
                              c_i64_p],
    "b200rec_stage_batch": [vp, C.c_int, vp, vp],
    "b200rec_step_staged": [vp, vp, C.POINTER(C.c_float)],
    "b200rec_step_staged_async": [vp, vp],
    "b200rec_step_wait": [vp, vp, C.POINTER(C.c_float)],
    "b200rec_capture_begin": [vp, vp],
    "b200rec_capture_end": [vp, C.POINTER(C.c_int), vp],
    "b200rec_graph_launch": [vp, C.c_int, vp],
    "b200rec_p2p_compose_dst_dev": [vp, C.c_int, C.c_int64, vp, vp, vp],
    "b200rec_table_init_uniform_sharded": [vp, C.c_uint64, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int64],
    "b200rec_shard_plan_dev": [vp, C.c_int64, C.c_int, C.c_int64, C.c_int, vp, vp, vp, vp, vp],
    "b200rec_table_lookup_padded_dev": [vp, C.c_int64, vp, vp, vp, vp],
    "b200rec_step_rows_dev": [vp, C.c_int, vp, vp, vp, C.c_int64, vp, vp, vp, C.c_int, vp],
    "b200rec_p2p_dispatch_ids_dev": [vp, C.c_int64, vp, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, vp, vp, vp, vp, vp,
                                     vp],
    "b200rec_p2p_wait_dev": [vp, vp, C.c_int, C.c_int, C.c_int, vp],
    "b200rec_p2p_gather_dev": [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp],
    "b200rec_p2p_push_grads_dev": [vp, C.c_int64, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp],
    "b200rec_p2p_reduce_push_dev": [vp, C.c_int, C.c_int64, C.c_int, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                    vp, vp, vp, vp, vp],
    "b200rec_table_apply_sgd_dev": [vp, C.c_int64, vp, vp, vp, vp, C.c_float, vp],
    "b200rec_table_apply_optimizer_dev": [vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int64, C.c_int64, vp, vp, vp,
                                          vp, vp],
    "b200rec_model_apply_optimizer_dev": [vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int64, vp],
    "b200rec_table_apply_optimizer_stepdev_dev": [vp, C.c_int, C.c_float, C.c_float, C.c_float, vp, C.c_int64, vp, vp,
                                                  vp, vp, vp],
    "b200rec_model_apply_optimizer_stepdev_dev": [vp, C.c_int, C.c_float, C.c_float, C.c_float, vp, vp],
    "b200rec_model_step_counter": [vp, C.POINTER(vp)],
    "b200rec_table_status": [vp, C.c_int, vp],
    "b200rec_alloc_epoch": [c_i64_p],
    "b200rec_encoder_mats_len": [vp, c_i64_p],
    "b200rec_duplicate_table_update_grad_input": [C.c_int, C.c_int, C.c_int64, vp, vp],
    "b200rec_higher_order_update_output": [vp, C.c_int, vp, vp, vp],
    "b200rec_higher_order_update_grad_input": [vp, C.c_int, vp, vp, vp, vp],
    "b200rec_higher_order_acc_grad_parameters": [vp, C.c_int, vp, vp, vp, C.c_float, vp],
    "b200rec_higher_order_backward": [vp, C.c_int, vp, vp, vp, vp],
    "b200rec_cin_update_output": [vp, C.c_int, vp, vp, vp],
    "b200rec_cin_update_grad_input": [vp, C.c_int, vp, vp, vp, vp],
    "b200rec_cin_acc_grad_parameters": [vp, C.c_int, vp, vp, vp, C.c_float, vp],
    "b200rec_cin_backward": [vp, C.c_int, vp, vp, vp, vp],
    "b200rec_cross_update_output": [vp, C.c_int, vp, vp, vp],
    "b200rec_cross_update_grad_input": [vp, C.c_int, vp, vp, vp, vp],
    "b200rec_cross_acc_grad_parameters": [vp, C.c_int, vp, vp, vp, C.c_float, vp],
    "b200rec_cross_backward": [vp, C.c_int, vp, vp, vp, vp],
    "b200rec_product_update_output": [vp, C.c_int, vp, vp, vp],
    "b200rec_product_update_grad_input": [vp, C.c_int, vp, vp, vp, vp],
    "b200rec_product_acc_grad_parameters": [vp, C.c_int, vp, vp, vp, C.c_float, vp],
    "b200rec_product_backward": [vp, C.c_int, vp, vp, vp, vp],
    "b200rec_scatter_update_output": [C.c_int, C.c_int, C.c_int, C.c_int64, vp, vp, vp],
    "b200rec_scatter_update_grad_input": [C.c_int, C.c_int, C.c_int, C.c_int64, vp, vp, vp],
    "b200rec_gather_update_output": [C.c_int] * 5 + [vp] * 5,
    "b200rec_gather_update_grad_input": [C.c_int] * 5 + [vp] * 5,
    "b200rec_dotproduct2_update_output": [C.c_int, C.c_int64, C.c_int, vp, vp, vp],
    "b200rec_dotproduct2_update_grad_input": [C.c_int, C.c_int64, C.c_int, vp, vp, vp, vp, vp],
    "b200rec_second_order_update_output": [C.c_int] * 4 + [vp, vp],
    "b200rec_second_order_update_grad_input": [C.c_int] * 4 + [vp, vp, vp],
    "b200rec_linear_update_output": [C.c_int] * 4 + [vp, vp, vp, C.c_int, vp],
    "b200rec_linear_update_grad_input": [C.c_int] * 4 + [vp, vp, vp],
    "b200rec_linear_acc_grad_parameters": [C.c_int] * 4 + [vp, vp, C.c_float, vp, vp],
}

_lib = None


class B200RecError(RuntimeError):
    """CUDA / state / memory failures (status -4, -5, -6)."""


def lib():
    """Load libb200rec.so (built in-tree by build.py).  Fails loudly when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200RecError(
                f"{LIB_PATH} is missing: run `python recommendation-models_b200/build.py` "
                "(there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = C.c_char_p if name == "b200rec_last_error" else C.c_int
        _lib = l
    return _lib


def last_error() -> str:
    return (lib().b200rec_last_error() or b"").decode("utf-8", "replace")


def check(status: int):
    """Non-zero status -> exception.  Argument / shape / index problems raise ValueError, the
    Python stand-in for the IllegalArgumentException of the reference's `require`s."""
    if status == OK:
        return
    msg = f"b200rec status {status}: {last_error()}"
    if status in (ERR_ARG, ERR_SHAPE, ERR_INDEX):
        raise ValueError(msg)
    raise B200RecError(msg)


def ptr(a):
    """void* of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data_as(vp)


def f32(a, copy=False):
    if a is None:
        return None
    return np.array(a, dtype=np.float32, order="C", copy=True) if copy else np.ascontiguousarray(a, dtype=np.float32)


def i32(a):
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=np.int32)


def device_count() -> int:
    n = C.c_int(0)
    check(lib().b200rec_device_count(C.byref(n)))
    return n.value


def launch_count() -> int:
    n = C.c_int64(0)
    check(lib().b200rec_launch_count(C.byref(n)))
    return n.value


def profile_begin():
    check(lib().b200rec_profile_begin())


def profile_end():
    """-> list of (phase, kernel, launches, total_ms) for the launches since profile_begin()."""
    buf = C.create_string_buffer(1 << 16)
    need = C.c_int64(0)
    check(lib().b200rec_profile_end(buf, len(buf), C.byref(need)))
    rows = []
    for line in buf.value.decode().splitlines():
        tag, name, cnt, ms = line.split("|")
        rows.append((tag, name, int(cnt), float(ms)))
    return rows


def profile_overhead_us(device=0) -> float:
    """What the event pair of a profile pass adds to each bracketed kernel (microseconds)."""
    us = C.c_float(0)
    check(lib().b200rec_profile_overhead_us(device, C.byref(us)))
    return float(us.value)


def set_default_gemm_mode(mode: int):
    """0 = fp32 FFMA, 1 = 3xTF32 tcgen05 (default), 2 = 1xTF32 tcgen05 (not parity grade)."""
    check(lib().b200rec_set_default_gemm_mode(int(mode)))
