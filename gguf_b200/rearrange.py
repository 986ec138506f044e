"""Block-granular tensor rearrangement on the GPU: host mirror of the reference's operators that move
whole blocks / rows (xtask/src/utils/operator/merge.rs `concat` / `split`, permute_qk.rs `permute_qk`)
over `ggq_rearrange[_device]` (include/ggq.h), the device counterpart of mem-rearrange's `Rearranging`.

`ArrayLayout` restates the four ndarray-layout 0.2.1 operations those operators call; shapes count
elements of `unit` bytes, strides and offsets are bytes, dim 0 is the fastest one (ggml's ne[0]).
All data movement happens on the GPU; there is no CPU path here.
"""
import ctypes

import numpy as np

from ._lib import Layout, lib
from .quants import GgqError, QuantizeError, block_info

_FLOAT_SIZES = {0: (1, 4), 1: (1, 2), 30: (1, 2)}


def type_size(ty):
    """`GGmlType::size()` -> (block elements, block bytes)."""
    return _FLOAT_SIZES[ty] if ty in _FLOAT_SIZES else block_info(ty)


class ArrayLayout:
    def __init__(self, shape, strides, offset=0):
        self.shape, self.strides, self.offset = list(map(int, shape)), list(map(int, strides)), int(offset)

    @classmethod
    def new_contiguous(cls, shape, unit):
        """`ArrayLayout::new_contiguous(shape, LittleEndian, unit)` (merge.rs:359-364, permute_qk.rs:57,60)."""
        strides, mul = [], int(unit)
        for d in shape:
            strides.append(mul)
            mul *= int(d)
        return cls(shape, strides)

    def tile_le(self, axis, tiles):
        """permute_qk.rs:58 — one dim becomes several, the first tile fastest."""
        assert int(np.prod(tiles)) == self.shape[axis]
        shape, strides = [], []
        for i, (d, s) in enumerate(zip(self.shape, self.strides)):
            if i != axis:
                shape.append(d)
                strides.append(s)
                continue
            for t in tiles:
                shape.append(int(t))
                strides.append(s)
                s *= int(t)
        return ArrayLayout(shape, strides, self.offset)

    def transpose(self, perm):
        """permute_qk.rs:59 — the dims in `perm` are placed, in that order, at the sorted positions of `perm`."""
        pos = sorted(perm)
        assert len(set(pos)) == len(pos)
        shape, strides = list(self.shape), list(self.strides)
        for k, j in zip(pos, perm):
            shape[k], strides[k] = self.shape[j], self.strides[j]
        return ArrayLayout(shape, strides, self.offset)

    def split(self, axis, parts):
        """merge.rs:310,337 — consecutive sub-ranges of one dim."""
        assert sum(parts) == self.shape[axis]
        out, start = [], 0
        for p in parts:
            shape = list(self.shape)
            shape[axis] = int(p)
            out.append(ArrayLayout(shape, self.strides, self.offset + start * self.strides[axis]))
            start += int(p)
        return out

    def c(self):
        if len(self.shape) > 4:
            raise ValueError("ggq_layout holds at most 4 dims")
        l = Layout()
        l.ndim = len(self.shape)
        for i, (d, s) in enumerate(zip(self.shape, self.strides)):
            l.shape[i], l.strides[i] = d, s
        l.offset = self.offset
        return l


def _check(rc):
    if rc == 1:
        raise QuantizeError("Indivisible")
    if rc == 2:
        raise QuantizeError("LengthMismatch")
    if rc != 0:
        raise GgqError(rc, lib().ggq_last_error().decode())


def rearrange(dst, dst_layout, src, src_layout, unit):
    """`Rearranging::new(&dst, &src, unit).launch(dst, src)` on host numpy arrays (uint8 views), synchronous."""
    d, s = dst_layout.c(), src_layout.c()
    _check(lib().ggq_rearrange(dst.ctypes.data, ctypes.byref(d), src.ctypes.data, ctypes.byref(s), int(unit)))


def rearrange_device(dst_ptr, dst_layout, src_ptr, src_layout, unit, stream=0):
    """The same on device pointers, enqueued on `stream`."""
    d, s = dst_layout.c(), src_layout.c()
    _check(lib().ggq_rearrange_device(int(dst_ptr), ctypes.byref(d), int(src_ptr), ctypes.byref(s), int(unit), int(stream)))


def block_layout(ty, shape):
    """merge.rs:359-364 `layout(ty, shape)`: shape[0] in blocks, element = one block."""
    be, bb = type_size(ty)
    shape = list(shape)
    assert shape[0] % be == 0
    shape[0] //= be
    return ArrayLayout.new_contiguous(shape, bb), bb


def permute_qk_layouts(ty, shape, nh):
    """(dst, src, unit) of permute_qk.rs:46-66 for a tensor of `shape` (ggml order) with `nh` heads."""
    be, bb = type_size(ty)
    if len(shape) == 1:
        assert be == 1
        c, r = bb, shape[0]
    else:
        assert shape[0] % be == 0
        c, r = shape[0] // be * bb, shape[1]
    src = ArrayLayout.new_contiguous([c, r], 1).tile_le(1, [r // nh // 2, 2, nh]).transpose([2, 1])
    return ArrayLayout.new_contiguous(src.shape, 1), src, 1


def permute_qk(data, ty, shape, nh):
    """permute_qk.rs:46-69 on a host uint8 array; returns the permuted bytes."""
    data = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    out = np.empty_like(data)
    dl, sl, unit = permute_qk_layouts(ty, shape, nh)
    rearrange(out, dl, data, sl, unit)
    return out


def concat(axis, tensors):
    """merge.rs:288-325 — `tensors` = [(ty, shape, uint8 data)]; returns (ty, shape, data)."""
    ty, shape = tensors[0][0], list(tensors[0][1])
    if len(shape) == 1:
        axis = 0
    for t, s, _ in tensors[1:]:
        assert t == ty and len(s) == len(shape)
        for i, d in enumerate(s):
            if i == axis:
                shape[i] += d
            else:
                assert shape[i] == d
    be, bb = type_size(ty)
    whole, unit = block_layout(ty, shape)
    parts = [s[axis] // be if axis == 0 else s[axis] for _, s, _ in tensors]
    out = np.zeros(int(np.prod(shape)) // be * bb, np.uint8)
    for (t, s, d), view in zip(tensors, whole.split(axis, parts)):
        rearrange(out, view, np.ascontiguousarray(d).view(np.uint8).reshape(-1), block_layout(t, s)[0], unit)
    return ty, tuple(shape), out


def split(axis, tensor, parts):
    """merge.rs:327-357 — returns [(ty, shape, data)] for the element counts `parts` along `axis`."""
    ty, shape, data = tensor
    if len(shape) == 1:
        axis = 0
    assert shape[axis] == sum(parts)
    be, bb = type_size(ty)
    whole, unit = block_layout(ty, shape)
    data = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    outs = []
    for p, view in zip(parts, whole.split(axis, [p // be if axis == 0 else p for p in parts])):
        s = list(shape)
        s[axis] = p
        out = np.empty(int(np.prod(s)) // be * bb, np.uint8)
        rearrange(out, block_layout(ty, s)[0], data, view, unit)
        outs.append((ty, tuple(s), out))
    return outs
