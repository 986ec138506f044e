"""ctypes binding for the CPU ORACLE (oracle/libggq_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under gguf_b200/ may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libggq_oracle.so")

F32, F16, Q4_0, Q4_1, Q5_0, Q5_1, Q8_0, Q8_1 = 0, 1, 2, 3, 6, 7, 8, 9
Q2K, Q3K, Q4K, Q5K, Q6K, Q8K, BF16 = 10, 11, 12, 13, 14, 15, 30
OK, INDIVISIBLE, LENGTH_MISMATCH = 0, 1, 2

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.ggo_block_info.argtypes = [ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
        L.ggo_quantize_slice.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t,
                                         ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        L.ggo_dequantize_slice.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t,
                                           ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        L.ggo_f32_to_f16.argtypes = [ctypes.c_float]
        L.ggo_f32_to_f16.restype = ctypes.c_uint16
        L.ggo_f32_to_bf16.argtypes = [ctypes.c_float]
        L.ggo_f32_to_bf16.restype = ctypes.c_uint16
        L.ggo_f16_to_f32.argtypes = [ctypes.c_uint16]
        L.ggo_f16_to_f32.restype = ctypes.c_float
        L.ggo_bf16_to_f32.argtypes = [ctypes.c_uint16]
        L.ggo_bf16_to_f32.restype = ctypes.c_float
        _lib = L
    return _lib


def block_info(ty):
    e, b = ctypes.c_uint32(), ctypes.c_uint32()
    if lib().ggo_block_info(ty, ctypes.byref(e), ctypes.byref(b)) != 0:
        raise ValueError(f"unsupported type {ty}")
    return e.value, b.value


_NP = {F32: np.float32, F16: np.uint16, BF16: np.uint16}


def float_view(arr, fdt):
    """Bit view of a float-side array: f32 stays f32, f16/bf16 are carried as uint16 bit patterns."""
    a = np.ascontiguousarray(arr)
    if fdt == F32:
        assert a.dtype == np.float32
    else:
        if a.dtype == np.float16:
            a = a.view(np.uint16)
        assert a.dtype == np.uint16
    return a


def quantize(ty, fdt, src, threads=1):
    """src: float-side 1-D array (float32, or uint16/float16 bits). Returns uint8 array of blocks."""
    src = float_view(src, fdt).reshape(-1)
    e, b = block_info(ty)
    nb = src.size // e
    dst = np.empty(nb * b, dtype=np.uint8)
    rc = lib().ggo_quantize_slice(ty, fdt, dst.ctypes.data, nb, src.ctypes.data, src.size, threads)
    if rc != 0:
        raise RuntimeError(f"oracle quantize rc={rc}")
    return dst


def dequantize(ty, fdt, blocks, threads=1):
    """blocks: uint8 array. Returns float32 array (fdt F32) or uint16 bit patterns (F16/BF16)."""
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1)
    e, b = block_info(ty)
    nb = blocks.size // b
    dst = np.empty(nb * e, dtype=_NP[fdt])
    rc = lib().ggo_dequantize_slice(ty, fdt, dst.ctypes.data, dst.size, blocks.ctypes.data, nb, threads)
    if rc != 0:
        raise RuntimeError(f"oracle dequantize rc={rc}")
    return dst


def quantize_rc(ty, fdt, dst_blocks, src_elems):
    """Error-path probe: lengths only (buffers are sized generously)."""
    e, b = block_info(ty)
    src = np.zeros(max(src_elems, 1), dtype=_NP[fdt])
    dst = np.zeros(max(dst_blocks, 1) * b, dtype=np.uint8)
    return lib().ggo_quantize_slice(ty, fdt, dst.ctypes.data, dst_blocks, src.ctypes.data, src_elems, 1)


def dequantize_rc(ty, fdt, dst_elems, src_blocks):
    e, b = block_info(ty)
    dst = np.zeros(max(dst_elems, 1), dtype=_NP[fdt])
    src = np.zeros(max(src_blocks, 1) * b, dtype=np.uint8)
    return lib().ggo_dequantize_slice(ty, fdt, dst.ctypes.data, dst_elems, src.ctypes.data, src_blocks, 1)
