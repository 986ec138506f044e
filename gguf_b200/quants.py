"""Host-side mirror of the reference's `ggml-quants` trait surface over libggq.so.

Reference interface (paths under /root/reference/):
  * ``DataBlock::{COUNT, ZEROS}``               ggml-quants/src/lib.rs:11-21
  * ``Quantize<T, N>::{quantize, dequantize}``  ggml-quants/src/lib.rs:53-90
  * ``QuantExt<T, N>::{quantize_slice, dequantize_slice}`` + ``QuantizeError``  lib.rs:98-148
  * caller: ``xtask/src/utils/operator/cast.rs:140-156``

Slices are numpy arrays over host memory (the reference's `&[T]` / `&mut [Blk]`).  `[Blk]` slices are
``uint8`` arrays whose byte length is a whole number of blocks; `[T]` slices are ``float32``,
``float16`` or — for bf16, which numpy lacks — ``uint16`` bit patterns with ``fdt=BF16``.
All arithmetic happens on the GPU; this module only validates shapes and forwards pointers.
"""
import ctypes

import numpy as np

from ._lib import lib

# GGmlType discriminants (ggus/src/tensor.rs:15-50)
F32, F16, Q4_0, Q4_1, Q5_0, Q5_1, Q8_0, Q8_1 = 0, 1, 2, 3, 6, 7, 8, 9
Q2K, Q3K, Q4K, Q5K, Q6K, Q8K, BF16 = 10, 11, 12, 13, 14, 15, 30

TYPE_NAMES = {F32: "F32", F16: "F16", Q4_0: "Q4_0", Q4_1: "Q4_1", Q5_0: "Q5_0", Q5_1: "Q5_1", Q8_0: "Q8_0",
              Q8_1: "Q8_1", Q2K: "Q2K", Q3K: "Q3K", Q4K: "Q4K", Q5K: "Q5K", Q6K: "Q6K", Q8K: "Q8K", BF16: "BF16"}
BLOCK_TYPES = [F16, BF16, Q4_0, Q4_1, Q5_0, Q5_1, Q8_0, Q8_1, Q2K, Q3K, Q4K, Q5K, Q6K, Q8K]
FLOAT_TYPES = [F32, F16, BF16]
FLOAT_SIZE = {F32: 4, F16: 2, BF16: 2}

GGQ_OK, GGQ_ERR_INDIVISIBLE, GGQ_ERR_LENGTH_MISMATCH = 0, 1, 2


class QuantizeError(Exception):
    """`QuantizeError::{Indivisible, LengthMismatch}` (lib.rs:107-113)."""

    def __init__(self, kind):
        super().__init__(kind)
        self.kind = kind


class GgqError(RuntimeError):
    """Any other non-zero status from libggq (CUDA failure, unsupported type, ...)."""

    def __init__(self, code, msg):
        super().__init__(f"libggq status {code}: {msg}")
        self.code = code


def _check(rc):
    if rc == GGQ_OK:
        return
    if rc == GGQ_ERR_INDIVISIBLE:
        raise QuantizeError("Indivisible")
    if rc == GGQ_ERR_LENGTH_MISMATCH:
        raise QuantizeError("LengthMismatch")
    raise GgqError(rc, lib().ggq_last_error().decode())


def block_info(ty):
    """(COUNT, size_of::<Blk>()) of a block type."""
    e, b = ctypes.c_uint32(), ctypes.c_uint32()
    _check(lib().ggq_block_info(ty, ctypes.byref(e), ctypes.byref(b)))
    return e.value, b.value


def _fdt_of(arr, fdt):
    if fdt is not None:
        return fdt
    if arr.dtype == np.float32:
        return F32
    if arr.dtype == np.float16:
        return F16
    raise TypeError("pass fdt=BF16 (or F16) explicitly for uint16 bit-pattern arrays")


def _blocks_len(arr, ty):
    _, size = block_info(ty)
    if arr.dtype != np.uint8 or not arr.flags.c_contiguous:
        raise TypeError("[Blk] slices are C-contiguous uint8 arrays")
    if arr.nbytes % size:
        raise ValueError(f"byte length {arr.nbytes} is not a whole number of {TYPE_NAMES[ty]} blocks ({size} B)")
    return arr.nbytes // size


def _float_len(arr, fdt):
    if not arr.flags.c_contiguous or arr.dtype.itemsize != FLOAT_SIZE[fdt]:
        raise TypeError("[T] slices are C-contiguous arrays of the float-side element size")
    return arr.size


def quantize_slice(ty, dst, src, fdt=None):
    """`<Blk as QuantExt<T, N>>::quantize_slice(dst, src)` (lib.rs:121-133). Raises QuantizeError."""
    fdt = _fdt_of(src, fdt)
    _check(lib().ggq_quantize_slice(ty, fdt, dst.ctypes.data, _blocks_len(dst, ty), src.ctypes.data, _float_len(src, fdt)))


def dequantize_slice(ty, dst, src, fdt=None):
    """`<Blk as QuantExt<T, N>>::dequantize_slice(dst, src)` (lib.rs:135-147). Raises QuantizeError."""
    fdt = _fdt_of(dst, fdt)
    _check(lib().ggq_dequantize_slice(ty, fdt, dst.ctypes.data, _float_len(dst, fdt), src.ctypes.data, _blocks_len(src, ty)))


def quantize(ty, src, fdt=None):
    """cast.rs:140-148 `quantize::<Blk, T, N>`: allocate the destination and quantize into it."""
    fdt = _fdt_of(src, fdt)
    count, size = block_info(ty)
    src = np.ascontiguousarray(src).reshape(-1)
    if src.size % count:
        raise QuantizeError("Indivisible")
    dst = np.empty(src.size // count * size, dtype=np.uint8)
    quantize_slice(ty, dst, src, fdt)
    return dst


def dequantize(ty, src, fdt=F32):
    """cast.rs:150-156 `dequantize::<Blk, T, N>`."""
    count, _ = block_info(ty)
    src = np.ascontiguousarray(src, dtype=np.uint8).reshape(-1)
    n = _blocks_len(src, ty) * count
    dst = np.empty(n, dtype=np.float32 if fdt == F32 else np.uint16)
    dequantize_slice(ty, dst, src, fdt)
    return dst


def slices(jobs):
    """`ggq_slices`: several slice calls streamed through one pipeline, as ONE synchronous call.
    `jobs` is a list of ("quantize" | "dequantize", type, dst_array, src_array[, fdt]) with the same
    array conventions as quantize_slice / dequantize_slice.  Raises like they do; validates all first."""
    from ._lib import SliceJob
    arr = (SliceJob * len(jobs))()
    keep = []
    for i, job in enumerate(jobs):
        kind, ty, dst, src = job[:4]
        fdt = job[4] if len(job) > 4 else None
        if kind == "quantize":
            fdt = _fdt_of(src, fdt)
            arr[i] = SliceJob(ty, fdt, 1, dst.ctypes.data, _blocks_len(dst, ty), src.ctypes.data, _float_len(src, fdt))
        elif kind == "dequantize":
            fdt = _fdt_of(dst, fdt)
            arr[i] = SliceJob(ty, fdt, 0, dst.ctypes.data, _float_len(dst, fdt), src.ctypes.data, _blocks_len(src, ty))
        else:
            raise ValueError(kind)
        keep.append((dst, src))
    _check(lib().ggq_slices(arr, len(jobs)))


# ---- device-pointer API (torch tensors or raw addresses) ------------------------------------------
def _ptr(x):
    return x.data_ptr() if hasattr(x, "data_ptr") else int(x)


def quantize_slice_device(ty, fdt, dst, dst_blocks, src, src_elems, stream=0):
    _check(lib().ggq_quantize_slice_device(ty, fdt, _ptr(dst), dst_blocks, _ptr(src), src_elems, stream))


def dequantize_slice_device(ty, fdt, dst, dst_elems, src, src_blocks, stream=0):
    _check(lib().ggq_dequantize_slice_device(ty, fdt, _ptr(dst), dst_elems, _ptr(src), src_blocks, stream))


def slices_device(jobs, stream=0):
    """`ggq_slices_device`: several device-pointer slice calls enqueued as one call; consecutive dequantize jobs of
    one float side share a single grid.  `jobs`: ("quantize" | "dequantize", type, fdt, dst, dst_len, src, src_len)
    with the lengths of the per-call entry points; or a prebuilt ctypes SliceJob array."""
    from ._lib import SliceJob
    if isinstance(jobs, ctypes.Array):
        arr = jobs
    else:
        arr = (SliceJob * len(jobs))()
        for i, (kind, ty, fdt, dst, dst_len, src, src_len) in enumerate(jobs):
            arr[i] = SliceJob(ty, fdt, 1 if kind == "quantize" else 0, _ptr(dst), dst_len, _ptr(src), src_len)
    _check(lib().ggq_slices_device(arr, len(arr), stream))


class PinnedBuffer:
    """Page-locked host buffer from `ggq_host_alloc`, exposed as a numpy uint8 array (`.array`)."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = lib().ggq_host_alloc(self.nbytes)
        if not self.ptr:
            raise GgqError(-2, lib().ggq_last_error().decode())
        buf = (ctypes.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.uint8, count=self.nbytes)

    def view(self, dtype):
        return self.array.view(dtype)

    def free(self):
        if self.ptr:
            self.array = None
            lib().ggq_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
