"""ctypes binding of libbuddy_b200.so (the C ABI declared in include/buddy_b200.h).

There is no CPU fallback: if the library is missing, or a call fails, this raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbuddy_b200.so")
_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int32
c_i64 = ctypes.c_int64
c_float = ctypes.c_float


class BuddyError(RuntimeError):
    pass


class GemmDesc(ctypes.Structure):
    """Mirror of `buddy_gemm_desc` (include/buddy_b200.h)."""
    _fields_ = [
        ("a", c_void_p), ("a_c", c_int), ("a_stride_w", c_i64), ("a_stride_h", c_i64), ("a_stride_b", c_i64),
        ("a2", c_void_p), ("a2_c", c_int), ("a2_stride_w", c_i64), ("a2_stride_h", c_i64), ("a2_stride_b", c_i64),
        ("b", c_void_p), ("b_rows", c_int), ("b_t", c_int), ("b_stride_n", c_i64), ("b_stride_t", c_i64),
        ("b2", c_void_p), ("b2_rows", c_int), ("b2_stride_n", c_i64),
        ("batch", c_int), ("H", c_int), ("W", c_int),
        ("taps", c_int), ("b_batched", c_int), ("n_total", c_int), ("n_tile", c_int),
        ("out", c_void_p), ("out_fp16", c_int), ("ldc", c_i64), ("col_off", c_int),
        ("bias", c_void_p), ("bias_b", c_void_p), ("resid", c_void_p), ("ld_res", c_i64),
        ("scale", c_float), ("stats", c_void_p), ("max_ctas", c_int), ("k_total", c_int), ("k2_total", c_int),
        ("a8", c_void_p), ("a8_c", c_int), ("a8_stride_w", c_i64), ("a8_stride_h", c_i64), ("a8_stride_b", c_i64),
        ("b8", c_void_p),
        ("a8_2", c_void_p), ("a8_2_c", c_int), ("a8_2_stride_w", c_i64), ("a8_2_stride_h", c_i64),
        ("a8_2_stride_b", c_i64), ("b8_2", c_void_p), ("no_staged_epilogue", c_int), ("no_cta_pairs", c_int), ("debug_flags", c_int), ("one_tap_per_stage", c_int),
        ("single_tile_per_cta", c_int),
        ("gnb_x", c_void_p), ("gnb_stats", c_void_p), ("gnb_gamma", c_void_p), ("gnb_beta", c_void_p),
        ("gnb_gsum", c_void_p), ("gnb_groups", c_int), ("gnb_eps", c_float), ("gnb_silu", c_int),
    ]


def lib():
    """Load the shared library once.  Raises BuddyError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise BuddyError(
                f"{_LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(buddy_b200 has no CPU fallback)")
        l = ctypes.CDLL(_LIB_PATH)
        l.buddy_last_error.restype = ctypes.c_char_p
        l.buddy_launch_count.restype = c_i64
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().buddy_last_error().decode("utf-8", "replace")
        raise BuddyError(f"{what} failed (rc={rc}): {msg}")


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda, "buddy_b200 kernels take CUDA tensors only (no CPU fallback)"
    return c_void_p(t.data_ptr())


def launch_count():
    return int(lib().buddy_launch_count())


def reset_launch_count():
    lib().buddy_reset_launch_count()


c_double_p = ctypes.c_void_p


class PackDesc(ctypes.Structure):
    """Mirror of `buddy_pack_desc`."""
    _fields_ = [
        ("src", c_void_p), ("off0", c_i64), ("st", c_i64), ("ndiv", c_int), ("sn_outer", c_i64), ("sn_inner", c_i64),
        ("kdiv", c_int), ("sk_outer", c_i64), ("sk_inner", c_i64),
        ("T", c_int), ("N", c_int), ("K", c_int), ("n_valid", c_int), ("k_valid", c_int), ("passes", c_int),
        ("w16", c_void_p), ("w8", c_void_p),
    ]


class GnDesc(ctypes.Structure):
    """Mirror of `buddy_gn_desc`."""
    _fields_ = [
        ("xa", c_void_p), ("xb", c_void_p), ("Ca", c_int), ("Cb", c_int),
        ("stats_a", c_void_p), ("stats_b", c_void_p), ("gamma", c_void_p), ("beta", c_void_p),
        ("batch", c_int), ("H", c_int), ("W", c_int), ("groups", c_int), ("eps", c_float),
        ("silu", c_int), ("mode", c_int), ("out", c_void_p), ("out_raw", c_void_p), ("split", c_int),
        ("out8", c_void_p), ("out_raw8", c_void_p),
    ]


class GnBwdDesc(ctypes.Structure):
    """Mirror of `buddy_gn_bwd_desc`."""
    _fields_ = [
        ("da", c_void_p), ("dskip", c_void_p), ("skip_scale", c_float),
        ("extra_a", c_void_p), ("extra_b", c_void_p), ("gsum", c_void_p),
        ("dxa", c_void_p), ("dxb", c_void_p), ("g16a", c_void_p), ("g16b", c_void_p), ("g16_scale", c_float),
        ("g8a", c_void_p), ("g8b", c_void_p), ("pass0_done", c_int),
    ]
