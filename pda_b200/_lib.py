"""ctypes binding of libpda_b200.so (the C ABI declared in include/pda_b200.h).

The shared library is the product; this module only loads it and declares prototypes.
There is no fallback: if the library is missing or no B200 is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libpda_b200.so")

c_f = C.POINTER(C.c_float)
c_i32 = C.POINTER(C.c_int32)
c_i64 = C.POINTER(C.c_int64)
c_u8 = C.POINTER(C.c_uint8)
c_d = C.POINTER(C.c_double)
c_vp = C.c_void_p


class PdaError(RuntimeError):
    pass


class PdaConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_users", C.c_int64), ("n_items", C.c_int64), ("embed_size", C.c_int32),
                ("train_mode", C.c_int32), ("batch_size", C.c_int32), ("lr", C.c_float), ("regs", C.c_float),
                ("max_batch", C.c_int64), ("temp_num", C.c_int32)]


# name -> (restype, argtypes); mirrors include/pda_b200.h one to one
PROTOTYPES = {
    "pda_last_error": (C.c_char_p, []),
    "pda_version": (C.c_int, []),
    "pda_device_count": (C.c_int, []),
    "pda_create": (C.c_int, [C.POINTER(PdaConfig), C.POINTER(c_vp)]),
    "pda_destroy": (None, [c_vp]),
    "pda_init_tables": (C.c_int, [c_vp, C.c_uint32]),
    "pda_set_table": (C.c_int, [c_vp, C.c_int, c_vp]),
    "pda_get_table": (C.c_int, [c_vp, C.c_int, c_vp]),
    "pda_table_ptr": (c_vp, [c_vp, C.c_int]),
    "pda_get_adam_powers": (C.c_int, [c_vp, c_vp]),
    "pda_set_adam_powers": (C.c_int, [c_vp, c_vp]),
    "pda_synchronize": (C.c_int, [c_vp]),
    "pda_set_adam_mode": (C.c_int, [c_vp, C.c_int]),
    "pda_set_deterministic": (C.c_int, [c_vp, C.c_int]),
    "pda_set_hot_items": (C.c_int, [c_vp, c_vp, C.c_int32]),
    "pda_adam_stats": (C.c_int, [c_vp, c_vp, C.c_int]),
    "pda_profile_enable": (C.c_int, [c_vp, C.c_int]),
    "pda_profile_read": (C.c_int, [c_vp, c_vp, c_vp]),
    "pda_host_alloc": (c_vp, [C.c_int64]),
    "pda_host_free": (None, [c_vp]),
    "pda_set_train_csr": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int64, c_vp, C.c_int32]),
    "pda_set_train_csr_device": (C.c_int, [c_vp, c_vp, c_vp, c_vp, C.c_int64, c_vp, C.c_int64, c_vp, C.c_int32]),
    "pda_set_train_pop": (C.c_int, [c_vp, c_vp, C.c_int32]),
    "pda_sample_batch": (C.c_int, [c_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int64, c_vp]),
    "pda_get_batch": (C.c_int, [c_vp, C.c_int64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pda_train_step_host": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int64, c_vp]),
    "pda_train_steps_host": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int32, C.c_int64, c_vp]),
    "pda_train_step_device": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int64, c_vp]),
    "pda_train_steps_sampled": (C.c_int, [c_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_int64, c_vp]),
    "pda_set_global_batch": (C.c_int, [c_vp, C.c_int64]),
    "pda_grad_ptr": (c_vp, [c_vp, C.c_int]),
    "pda_loss_acc_ptr": (c_vp, [c_vp]),
    "pda_forward_backward_device": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int64, c_vp]),
    "pda_adam_apply": (C.c_int, [c_vp, c_vp]),
    "pda_adam_apply_part": (C.c_int, [c_vp, C.c_int, c_vp]),
    "pda_adam_dense_rows": (C.c_int, [c_vp, C.c_int, C.c_int64, C.c_int64, c_vp]),
    "pda_adopt_item_buffers": (C.c_int, [c_vp, c_vp, c_vp]),
    "pda_dp_exchange_adam_p2p": (C.c_int, [c_vp, c_vp, c_vp, C.c_int32, C.c_int32, C.c_int64, C.c_int64, c_vp]),
    "pda_dp_set_barrier": (C.c_int, [c_vp, c_vp, c_vp, C.c_int32, C.c_int32]),
    "pda_set_item_grad_buffer": (C.c_int, [c_vp, c_vp]),
    "pda_dp_exchange_adam": (C.c_int, [c_vp, c_vp, c_vp, C.c_int64, C.c_int64, c_vp]),
    "pda_stage_batch_host": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int64, c_vp]),
    "pda_stage_batch_host_async": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int64, c_vp]),
    "pda_staged_batch_wait": (C.c_int, [c_vp, c_vp]),
    "pda_read_loss_async": (C.c_int, [c_vp, c_vp, c_vp]),
    "pda_adam_dense_rows_ext": (C.c_int, [c_vp, C.c_int, C.c_int64, C.c_int64, c_vp, c_vp]),
    "pda_read_loss": (C.c_int, [c_vp, c_vp, c_vp]),
    "pda_read_loss_sums": (C.c_int, [c_vp, c_vp, C.c_int, c_vp]),
    "pda_gradients_host": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int64, c_vp, c_vp, c_vp]),
    "pda_gradients_temp_host": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pda_temp_item_bias_host": (C.c_int, [c_vp, C.c_int32, c_vp]),
    "pda_recommend_host": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int, c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
    "pda_recommend_device": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int, c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp, c_vp,
                                       c_vp]),
    "pda_tc_last_stats": (C.c_int, [c_vp, c_vp]),
    "pda_tc_plan_host": (C.c_int, [C.c_int64, C.c_int64, C.c_int32, C.c_int32, c_vp]),
    "pda_tc_debug_dense_host": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int, c_vp, c_vp, c_vp, c_vp]),
    "pda_scores_host": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int, c_vp, c_vp]),
    "pda_arg_top_k_2d_host": (C.c_int, [c_vp, C.c_int32, C.c_int32, C.c_int32, c_vp]),
    "pda_evaluate_matrix_host": (C.c_int, [c_vp, C.c_int32, C.c_int32, c_vp, c_vp, c_vp, C.c_int32, C.c_int32, c_vp]),
    "pda_debug_numerics": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, c_vp]),
    "pda_metrics_host": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int, c_vp, c_vp, c_vp, C.c_int64, c_vp, C.c_int, c_vp]),
}

_LIB = None


def load(build_if_missing: bool = True):
    """dlopen the library (building it with nvcc first if the .so is absent or stale)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if build_if_missing:
        from . import build as _build
        try:
            _build.build()
        except RuntimeError:
            if not os.path.exists(SO_PATH):
                raise
    if not os.path.exists(SO_PATH):
        raise PdaError(f"{SO_PATH} is missing: build it with `python -m pda_b200.build` (no CPU fallback exists)")
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().pda_last_error()
        raise PdaError(f"pda_b200 error {rc}: {msg.decode() if msg else '?'}")


def ptr(a):
    """numpy array / int / None -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)
