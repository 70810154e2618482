"""Helpers shared by the -m gpu tests."""
import ctypes

import torch


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-20))


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def error_flag():
    """Value of the library's device-side 'bounded wait timed out' flag (0 = healthy)."""
    from munit_b200 import _lib

    _lib.init()
    ptr = _lib.error_flag_ptr()
    torch.cuda.synchronize()
    host = ctypes.c_int(0)
    cudart = ctypes.CDLL("libcudart.so.12")
    cudart.cudaMemcpy(ctypes.byref(host), ctypes.c_void_p(ptr), 4, 2)
    return host.value
