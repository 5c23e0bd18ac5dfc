"""ctypes binding of libflk.so (include/flk.h).  Fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLK_LIB") or os.path.join(HERE, "libflk.so")   # FLK_LIB: tuning builds only

OK, E_INVALID, E_CUDA, E_CAPACITY, E_UNSUPPORTED, E_NCCL, E_STATE = 0, -1, -2, -3, -4, -5, -6


class FlkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libflk error {code}: {msg}")
        self.code = code


_vp, _f32, _i32, _u32, _u64 = C.c_void_p, C.c_float, C.c_int32, C.c_uint32, C.c_uint64
_P = C.POINTER

# name -> argtypes (every function returns int unless listed in _RESTYPE)
SIGNATURES = {
    "flk_field_create": [_f32, _f32, _f32, C.c_int, _u64, C.c_int, _P(_vp)],
    "flk_field_create_ex": [_f32, _f32, _f32, C.c_int, _u64, C.c_int, _u32, _P(_vp)],
    "flk_field_destroy": [_vp],
    "flk_field_clear": [_vp],
    "flk_set_params": [_vp, _f32, _f32, _f32, _f32, _f32, _f32, _f32, C.c_int],
    "flk_set_math_mode": [_vp, C.c_int],
    "flk_set_engine_semantics": [_vp, C.c_int],
    "flk_field_dims": [_vp, _P(_i32), _P(_i32), _P(_i32), _P(_i32)],
    "flk_set_layout": [_vp, C.c_int],
    "flk_get_layout": [_vp, _P(_i32), _P(_i32), _P(_u32)],
    "flk_set_object_location": [_vp, _u32, _f32, _f32, _f32, _f32],
    "flk_set_objects": [_vp, _u64, _vp, _vp, _vp],
    "flk_set_objects_ex": [_vp, _u64, _vp, _vp, _vp, _vp],
    "flk_snapshot_ex": [_vp, _vp, _vp, _vp, _vp, _P(_u64)],
    "flk_lazy_update": [_vp],
    "flk_step": [_vp, _u64, _u64, _vp, _u64],
    "flk_run": [_vp, _u64, _u64, _u64],
    "flk_get_neighbors_within_relax_distance": [_vp, _f32, _f32, _f32, _vp, _u32, _P(_u32)],
    "flk_get_neighbors_within_distance": [_vp, _f32, _f32, _f32, _vp, _u32, _P(_u32)],
    "flk_query_batch": [_vp, _u64, _vp, _f32, C.c_int, _vp, _vp, _u64],
    "flk_query_batch_objects": [_vp, _u64, _vp, _f32, C.c_int, _vp, _vp, _u64],
    "flk_get_neighbors_within_relax_distance_objects": [_vp, _f32, _f32, _f32, _vp, _u32, _P(_u32)],
    "flk_get_neighbors_within_distance_objects": [_vp, _f32, _f32, _f32, _vp, _u32, _P(_u32)],
    "flk_num_objects": [_vp, _P(_u64)],
    "flk_snapshot": [_vp, _vp, _vp, _vp, _P(_u64)],
    "flk_snapshot_async": [_vp, _vp, _vp, _vp, _P(_u64)],
    "flk_set_objects_dense": [_vp, _u64, _u32, _vp],
    "flk_snapshot_dense": [_vp, _u32, _u64, _vp],
    "flk_snapshot_dense_async": [_vp, _u32, _u64, _vp],
    "flk_frame_begin": [_vp, _vp, _vp, _P(_u64)],
    "flk_frame_wait": [_vp],
    "flk_get_binning": [_vp, _vp, _vp],
    "flk_get_object": [_vp, _u32, _P(_f32), _P(_i32)],
    "flk_synchronize": [_vp],
    "flk_elapsed_ms": [_vp, _P(_f32)],
    "flk_stream": [_vp, _P(_vp)],
    "flk_launch_count": [_vp, _P(_u64)],
    "flk_profile_enable": [_vp, C.c_int],
    "flk_use_graph": [_vp, C.c_int],
    "flk_profile_step_kernel": [_vp, _P(C.c_double), _P(_u64)],
    "flk_profile_rebin": [_vp, _P(C.c_double), _P(_u64)],
    "flk_host_alloc": [_P(_vp), _u64],
    "flk_host_free": [_vp],
    "flk_selftest_div2": [C.c_int, _u64, _u64, _P(_u64)],
    "flk_last_error": [],
    "flk_version": [],
    "flk_dist_unique_id": [_vp],
    "flk_dist_create": [_f32, _f32, _f32, _u64, C.c_int, C.c_int, C.c_int, _vp, _P(_vp)],
    "flk_dist_create_cols": [_f32, _f32, _f32, _u64, C.c_int, C.c_int, C.c_int, _vp, _vp, _P(_vp)],
    "flk_dist_owner": [_vp, _f32, _f32, _P(_i32)],
    "flk_dist_strip": [_vp, _P(_i32), _P(_i32)],
    "flk_dist_step": [_vp, _u64, _u64, _vp, _u64],
    "flk_dist_num_owned": [_vp, _P(_u64)],
    "flk_dist_commit": [_vp],
    "flk_dist_snapshot_owned": [_vp, _vp, _vp, _vp, _P(_u64)],
    "flk_dist_snapshot_owned_async": [_vp, _vp, _vp, _vp, _P(_u64)],
    "flk_mgpu_create": [_f32, _f32, _f32, _u64, C.c_int, _P(C.c_int), _P(_vp)],
    "flk_mgpu_create_cols": [_f32, _f32, _f32, _u64, C.c_int, _P(C.c_int), _vp, _P(_vp)],
    "flk_dist_create_tiles": [_f32, _f32, _f32, _u64, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _P(_vp)],
    "flk_mgpu_create_tiles": [_f32, _f32, _f32, _u64, C.c_int, _P(C.c_int), C.c_int, _vp, _vp, _P(_vp)],
    "flk_dist_tile": [_vp, _P(_i32), _P(_i32), _P(_i32), _P(_i32)],
    "flk_dist_create_ex": [_f32, _f32, _f32, _u64, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _u32, _P(_vp)],
    "flk_mgpu_create_ex": [_f32, _f32, _f32, _u64, C.c_int, _P(C.c_int), C.c_int, _vp, _vp, _u32, _P(_vp)],
    "flk_dist_snapshot_owned_ex": [_vp, _vp, _vp, _vp, _vp, _P(_u64)],
    "flk_balance_columns": [_vp, _u64, _u64, _f32, _f32, C.c_int, _vp],
    "flk_mgpu_commit": [_P(_vp), C.c_int],
    "flk_mgpu_step": [_P(_vp), C.c_int, _u64, _u64, _vp, _u64],
    "flk_mgpu_run": [_P(_vp), C.c_int, _u64, _u64, _u64],
}
_RESTYPE = {"flk_last_error": C.c_char_p, "flk_version": C.c_char_p}

_lib = None


def load() -> C.CDLL:
    """Load libflk.so; raises if it has not been built (run `python -m examples_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FlkError(E_CUDA, f"{LIB_PATH} is missing: build it with `python examples_b200/build.py` "
                                   "(the CUDA extension is the only implementation; there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, C.c_int)
        _lib = L
    return _lib


def check(rc: int) -> int:
    if rc != OK:
        raise FlkError(rc, load().flk_last_error().decode())
    return rc
