"""ctypes binding of libmunit_b200.so (the C-ABI declared in include/munit_b200.h).

The product path has no CPU or library fallback: if the shared library is missing the import of
this module raises, and if no sm_100 device is present `init()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MUNIT_LIB") or os.path.join(_HERE, "libmunit_b200.so")  # MUNIT_LIB: A/B builds (dev aid)

MAX_TAPS = 49
MAX_PHASES = 4
ACT = {"none": 0, "relu": 1, "lrelu": 2, "tanh": 3}
NORM = {"in": 0, "adain": 1, "ln": 2}


class TapGemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_rank", C.c_int32), ("a_dim", C.c_uint64 * 5), ("a_stride", C.c_uint64 * 5),
        ("a_box", C.c_uint32 * 5),
        ("b", C.c_void_p), ("b_rows", C.c_uint64), ("b_k", C.c_uint64), ("bn", C.c_int32),
        ("tw", C.c_int32), ("th", C.c_int32), ("tn", C.c_int32),
        ("out_w", C.c_int32), ("out_h", C.c_int32), ("n_img", C.c_int32),
        ("mx", C.c_int32 * 5), ("my", C.c_int32 * 5), ("mn", C.c_int32 * 5),
        ("num_taps", C.c_int32), ("chunks", C.c_int32), ("tap_off", (C.c_int32 * 5) * MAX_TAPS),
        ("phases", C.c_int32), ("b_k0", C.c_int32 * MAX_PHASES),
        ("o_yoff", C.c_int32 * MAX_PHASES), ("o_xoff", C.c_int32 * MAX_PHASES),
        ("out", C.c_void_p), ("o_sn", C.c_int64), ("o_sy", C.c_int64), ("o_sx", C.c_int64),
        ("o_ymul", C.c_int32), ("o_xmul", C.c_int32), ("n_store", C.c_int32),
        ("bias", C.c_void_p), ("act", C.c_int32), ("stages", C.c_int32), ("out_f16", C.c_int32), ("halo", C.c_int32),
        ("stats", C.c_void_p), ("stats_kind", C.c_int32), ("ksplit", C.c_int32), ("scratch", C.c_void_p),
        ("pair", C.c_int32),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_rank", C.c_int32), ("a_dim", C.c_uint64 * 5), ("a_stride", C.c_uint64 * 5),
        ("a_box", C.c_uint32 * 5), ("a_mx", C.c_int32 * 5), ("a_my", C.c_int32 * 5), ("a_mn", C.c_int32 * 5),
        ("b", C.c_void_p), ("b_rank", C.c_int32), ("b_dim", C.c_uint64 * 5), ("b_stride", C.c_uint64 * 5),
        ("b_box", C.c_uint32 * 5), ("b_mx", C.c_int32 * 5), ("b_my", C.c_int32 * 5), ("b_mn", C.c_int32 * 5),
        ("pw", C.c_int32), ("ph", C.c_int32), ("pn", C.c_int32),
        ("out_w", C.c_int32), ("out_h", C.c_int32), ("n_img", C.c_int32),
        ("m_total", C.c_int32), ("n_total", C.c_int32), ("bn", C.c_int32), ("num_taps", C.c_int32),
        ("tap_off", (C.c_int32 * 5) * MAX_TAPS),
        ("dw", C.c_void_p), ("s_m", C.c_int64), ("s_t", C.c_int64), ("s_n", C.c_int64),
        ("ksplit", C.c_int32), ("stages", C.c_int32), ("tap_on_a", C.c_int32),
    ]


if not os.path.isfile(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with munit_b200/csrc/build.sh (or __graft_entry__.build()); "
        "munit_b200 has no CPU fallback")

lib = C.CDLL(LIB_PATH)

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_SIGS = {
    "munit_version": ([], C.c_int),
    "munit_last_error": ([], C.c_char_p),
    "munit_init": ([], C.c_int),
    "munit_error_flag_ptr": ([C.POINTER(C.c_void_p)], C.c_int),
    "munit_tapgemm": ([C.POINTER(TapGemmDesc), _vp], C.c_int),
    "munit_splitk_finish": ([_vp, _vp, _i, _vp, _i, _i64, _i, _vp], C.c_int),
    "munit_wgrad": ([C.POINTER(WgradDesc), _vp], C.c_int),
    "munit_image_to_act": ([_vp, _vp, _i, _i, _i, _i, _i, _i, _vp], C.c_int),
    "munit_image_to_kwexp": ([_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp], C.c_int),
    "munit_kwexp_to_image_grad": ([_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp], C.c_int),
    "munit_act_to_nchw": ([_vp, _vp, _i, _i, _i, _i, _i, _i, _vp], C.c_int),
    "munit_nchw_to_act": ([_vp, _vp, _i, _i, _i, _i, _i, _i, _vp], C.c_int),
    "munit_u8_crop_normalize": ([_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp], C.c_int),
    "munit_halo_fill": ([_vp, _i, _i, _i, _i, _i, _vp], C.c_int),
    "munit_norm_splits": ([_i, _i], C.c_int),
    "munit_norm_stats": ([_vp, _i, _vp, _vp, _i, _i, _i, _vp], C.c_int),
    "munit_norm_finalize": ([_vp, _vp, _i, _vp, _vp, _i64, _f, _vp, _vp, _vp, _vp, _i, _i, _i, _vp], C.c_int),
    "munit_norm_stats_finalize": ([_vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _i64, _f, _vp, _vp, _vp, _vp, _i, _i, _i, _vp], C.c_int),
    "munit_norm_bwd_reduce_finalize": ([_vp, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _i64, _f, _vp, _vp, _vp,
                                        _vp, _vp, _i64, _i, _i, _i, _i, _vp], C.c_int),
    "munit_norm_finalize_parts": ([_vp, _i, _i, _i, _vp, _vp, _i64, _f, _vp, _vp, _vp, _vp, _i, _i, _i, _vp], C.c_int),
    "munit_norm_apply": ([_vp, _i, _vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp], C.c_int),
    "munit_norm_bwd_reduce": ([_vp, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp], C.c_int),
    "munit_norm_bwd_finalize": ([_vp, _i, _vp, _i64, _vp, _f, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _vp], C.c_int),
    "munit_norm_bwd_apply": ([_vp, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp], C.c_int),
    "munit_act_bwd": ([_vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _vp, _i, _vp], C.c_int),
    "munit_colsum": ([_vp, _vp, _i64, _i, _i, _vp], C.c_int),
    "munit_rspace_combine": ([_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp], C.c_int),
    "munit_rspace_expand": ([_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp], C.c_int),
    "munit_gather_cast": ([_vp, _vp, _vp, _i64, _vp], C.c_int),
    "munit_gather_cast_multi": ([_vp, _i, _i64, _vp], C.c_int),
    "munit_gather_add": ([_vp, _vp, _vp, _i64, _vp], C.c_int),
    "munit_cast_bf16": ([_vp, _vp, _i64, _vp], C.c_int),
    "munit_linear_fwd": ([_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp], C.c_int),
    "munit_mlp3_fwd": ([_vp] * 10 + [_i, _i, _i, _i, _vp], C.c_int),
    "munit_linear_bwd": ([_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp], C.c_int),
    "munit_gap_fwd": ([_vp, _vp, _i, _i, _i, _vp], C.c_int),
    "munit_gap_bwd": ([_vp, _vp, _i, _i, _i, _vp], C.c_int),
    "munit_dis_head_fwd": ([_vp, _vp, _vp, _f, _vp, _vp, _f, _i64, _i, _vp], C.c_int),
    "munit_dis_head_bwd": ([_vp, _vp, _vp, _f, _vp, _f, _vp, _vp, _vp, _i64, _i, _vp], C.c_int),
    "munit_avgpool3s2_fwd": ([_vp, _vp, _i, _i, _i, _vp], C.c_int),
    "munit_avgpool3s2_bwd": ([_vp, _vp, _i, _i, _i, _vp], C.c_int),
    "munit_l1_fwd": ([_vp, _vp, _vp, _f, _i64, _vp], C.c_int),
    "munit_l1_bwd": ([_vp, _vp, _vp, _f, _vp, _vp, _i64, _vp], C.c_int),
    "munit_l1_masked_fwd": ([_vp, _vp, _vp, _vp, _f, _i, _i, _i, _vp], C.c_int),
    "munit_l1_masked_bwd": ([_vp, _vp, _vp, _vp, _f, _vp, _vp, _i, _i, _i, _vp], C.c_int),
    "munit_l1_bf16_fwd": ([_vp, _vp, _vp, _f, _i64, _vp], C.c_int),
    "munit_l1_bf16_bwd": ([_vp, _vp, _vp, _f, _vp, _vp, _i64, _vp], C.c_int),
    "munit_adam": ([_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _f, _f, _f, _f, _f, _i, _f, _vp, _vp], C.c_int),
    "munit_fill_f32": ([_vp, _f, _i64, _vp], C.c_int),
    "munit_add_bf16": ([_vp, _vp, _i64, _vp], C.c_int),
    "munit_maxpool2_fwd": ([_vp, _i, _vp, _i, _i, _i, _i, _vp], C.c_int),
    "munit_maxpool2_bwd": ([_vp, _vp, _i, _vp, _i, _i, _i, _i, _vp], C.c_int),
    "munit_bn_finalize": ([_vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _f, _f, _i, _vp, _vp, _vp, _vp, _i, _i, _vp], C.c_int),
    "munit_bn_bwd_finalize": ([_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp], C.c_int),
    "munit_add_relu": ([_vp, _vp, _vp, _i64, _vp], C.c_int),
    "munit_mse_const_fwd": ([_vp, _f, _vp, _f, _i, _vp], C.c_int),
    "munit_mse_const_bwd": ([_vp, _f, _vp, _f, _vp, _i, _vp], C.c_int),
}
for _name, (_args, _res) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here == header/library mismatch: fail loudly
    _fn.argtypes = _args
    _fn.restype = _res

EXPORTS = tuple(_SIGS.keys())


class MunitError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        raise MunitError(f"{what} failed (status {rc}): {lib.munit_last_error().decode()}")


_initialised = False
launches = 0  # kernels launched through this binding (bench.py reports it as gpu_launches)


def init():
    """Bind to the current CUDA device; raises if there is no sm_100 GPU (no fallback)."""
    global _initialised
    if not _initialised:
        check(lib.munit_init(), "munit_init")
        _initialised = True


def error_flag_ptr() -> int:
    p = C.c_void_p()
    check(lib.munit_error_flag_ptr(C.byref(p)), "munit_error_flag_ptr")
    return p.value or 0
