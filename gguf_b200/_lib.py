"""ctypes loader for libggq.so (the C ABI declared in include/ggq.h).

There is deliberately no fallback: if the shared library has not been built (run
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C gguf_b200/csrc``) importing a
compute entry point raises, and every compute call fails with GGQ_ERR_CUDA when no GPU is present.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("GGQ_SO") or os.path.join(_HERE, "libggq.so")  # GGQ_SO: A/B builds of the same sources (tools/)

# every symbol include/ggq.h declares: (name, restype, argtypes)
_c = ctypes
SYMBOLS = [
    ("ggq_block_info", _c.c_int, [_c.c_uint32, _c.POINTER(_c.c_uint32), _c.POINTER(_c.c_uint32)]),
    ("ggq_last_error", _c.c_char_p, []),
    ("ggq_device_count", _c.c_int, []),
    ("ggq_set_device", _c.c_int, [_c.c_int]),
    ("ggq_set_shard_devices", _c.c_int, [_c.c_int]),
    ("ggq_quantize_slice", _c.c_int, [_c.c_uint32, _c.c_uint32, _c.c_void_p, _c.c_size_t, _c.c_void_p, _c.c_size_t]),
    ("ggq_dequantize_slice", _c.c_int, [_c.c_uint32, _c.c_uint32, _c.c_void_p, _c.c_size_t, _c.c_void_p, _c.c_size_t]),
    ("ggq_quantize_slice_device", _c.c_int,
     [_c.c_uint32, _c.c_uint32, _c.c_void_p, _c.c_size_t, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    ("ggq_dequantize_slice_device", _c.c_int,
     [_c.c_uint32, _c.c_uint32, _c.c_void_p, _c.c_size_t, _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    ("ggq_cast", _c.c_int, [_c.POINTER(_c.c_uint32), _c.c_int, _c.c_void_p, _c.c_void_p, _c.c_size_t]),
    ("ggq_type_nbytes", _c.c_size_t, [_c.c_uint32, _c.c_size_t]),
    ("ggq_host_alloc", _c.c_void_p, [_c.c_size_t]),
    ("ggq_host_free", None, [_c.c_void_p]),
    ("ggq_shutdown", None, []),
    ("ggq_launch_count", _c.c_uint64, []),
    ("ggq_version", _c.c_char_p, []),
]



class SliceJob(_c.Structure):
    _fields_ = [("type", _c.c_uint32), ("fdt", _c.c_uint32), ("quantize", _c.c_int), ("dst", _c.c_void_p), ("dst_len", _c.c_size_t),
                ("src", _c.c_void_p), ("src_len", _c.c_size_t)]


class ShardPiece(_c.Structure):
    _fields_ = [("job", _c.c_uint32), ("device", _c.c_int), ("elem_begin", _c.c_size_t), ("elem_end", _c.c_size_t)]


class ConvertStats(_c.Structure):
    _fields_ = [("n_tensors", _c.c_uint64), ("n_cast_tensors", _c.c_uint64), ("cast_elems", _c.c_uint64), ("bytes_in", _c.c_uint64),
                ("bytes_out", _c.c_uint64), ("seconds_plan", _c.c_double), ("seconds_convert", _c.c_double),
                ("seconds_sync", _c.c_double), ("n_devices", _c.c_int), ("n_out_files", _c.c_int), ("n_rearranged_tensors", _c.c_uint64),
                ("n_workers", _c.c_int), ("worker_seconds_read", _c.c_double), ("worker_seconds_write", _c.c_double),
                ("worker_seconds_gpu_wait", _c.c_double), ("h2d_bytes", _c.c_uint64), ("d2h_bytes", _c.c_uint64),
                ("n_direct_inputs", _c.c_int)]


class ConvertOptions(_c.Structure):
    _fields_ = [("n_devices", _c.c_int), ("max_tensors", _c.c_uint64), ("max_bytes", _c.c_uint64), ("no_tensor_first", _c.c_int),
                ("no_data", _c.c_int), ("direct_io", _c.c_int)]


class Layout(_c.Structure):
    """struct ggq_layout: ndarray-layout's ArrayLayout<4> (shape in units, strides / offset in bytes)."""
    _fields_ = [("ndim", _c.c_uint32), ("shape", _c.c_uint64 * 4), ("strides", _c.c_int64 * 4), ("offset", _c.c_int64)]


SYMBOLS += [
    ("ggq_rearrange_device", _c.c_int, [_c.c_void_p, _c.POINTER(Layout), _c.c_void_p, _c.POINTER(Layout), _c.c_size_t, _c.c_void_p]),
    ("ggq_rearrange", _c.c_int, [_c.c_void_p, _c.POINTER(Layout), _c.c_void_p, _c.POINTER(Layout), _c.c_size_t]),
    ("ggq_slices", _c.c_int, [_c.POINTER(SliceJob), _c.c_size_t]),
    ("ggq_slices_device", _c.c_int, [_c.POINTER(SliceJob), _c.c_size_t, _c.c_void_p]),
    ("ggq_plan_shards", _c.c_size_t, [_c.POINTER(SliceJob), _c.c_size_t, _c.c_int, _c.POINTER(ShardPiece), _c.c_size_t]),
    ("ggq_convert_gguf", _c.c_int, [_c.c_char_p, _c.c_char_p, _c.c_char_p, _c.c_int, _c.POINTER(ConvertStats)]),
    ("ggq_convert_gguf_ex", _c.c_int, [_c.POINTER(_c.c_char_p), _c.c_size_t, _c.c_char_p, _c.c_char_p, _c.POINTER(ConvertOptions),
                                       _c.POINTER(ConvertStats)]),
    ("ggq_convert_last_error", _c.c_char_p, []),
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
                "gguf_b200 has no CPU fallback.")
        L = ctypes.CDLL(SO_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
