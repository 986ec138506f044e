"""CPU oracle for the block-granular rearrangement operators — TEST INFRASTRUCTURE ONLY (imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; never by the product).

Restates, in numpy, what the reference does on the host for
  * `concat` / `split`      /root/reference/xtask/src/utils/operator/merge.rs:288-357
  * `permute_qk`            /root/reference/xtask/src/utils/operator/permute_qk.rs:46-69
  * `merge_linear` naming / grouping / ordering       merge.rs:22-83, 106-277
  * `Content::permute_qk` tensor selection            permute_qk.rs:11-44
through two third-party crates that are NOT in /root/reference (crates.io dependencies, not vendored):
  ndarray-layout 0.2.1 (Cargo.lock:340-343)  `ArrayLayout::{new_contiguous, tile_le, transpose, split}`
  mem-rearrange  0.1.0 (Cargo.lock:314-317)  `Rearranging::new(dst, src, unit).launch(dst_ptr, src_ptr)`
Their published semantics are restated below (`layout_*`, `rearrange`).

PARITY STATUS: **parity unpinned** — the reference has no test, fixture or golden vector for these
operators (its only operator tests are sort.rs:51 and cast.rs:218), and it cannot be compiled here (no
Rust toolchain).  What pins this oracle instead: every operator is written twice, once through the
restated layout algebra exactly as the reference composes it and once directly with numpy
reshape / swapaxes / concatenate from the meaning of the operation (`*_direct`); tests/test_oracle.py
checks that the two agree, and that `permute_qk_direct` equals the well-known llama.cpp
`convert_hf_to_gguf.py` permutation  w.reshape(n_head, 2, rows // n_head // 2, cols).swapaxes(1, 2).
"""
import numpy as np

TYPE_SIZE = {0: (1, 4), 1: (1, 2), 2: (32, 18), 3: (32, 20), 6: (32, 22), 7: (32, 24), 8: (32, 34), 9: (32, 36), 10: (256, 84),
             11: (256, 110), 12: (256, 144), 13: (256, 176), 14: (256, 210), 15: (256, 290), 30: (1, 2),
             24: (1, 1), 25: (1, 2), 26: (1, 4), 27: (1, 8), 28: (1, 8)}


def elements_to_bytes(ty, shape):
    """ggus/src/tensor.rs:83-96."""
    be, bb = TYPE_SIZE[ty]
    if len(shape) == 0:
        assert be == 1
        return bb
    assert shape[0] % be == 0
    return int(np.prod(shape[1:], dtype=np.int64)) * shape[0] // be * bb


# ---- ndarray-layout 0.2.1: a layout is (shape, strides, offset), strides / offset in bytes -------------
def layout_contiguous_le(shape, unit):
    strides = [unit * int(np.prod(shape[:i], dtype=np.int64)) for i in range(len(shape))]
    return (tuple(int(d) for d in shape), tuple(strides), 0)


def layout_tile_le(l, axis, tiles):
    shape, strides, off = l
    assert int(np.prod(tiles)) == shape[axis]
    new_strides = [strides[axis] * int(np.prod(tiles[:i], dtype=np.int64)) for i in range(len(tiles))]
    return (shape[:axis] + tuple(tiles) + shape[axis + 1:], strides[:axis] + tuple(new_strides) + strides[axis + 1:], off)


def layout_transpose(l, perm):
    shape, strides, off = l
    order = list(range(len(shape)))
    for slot, dim in zip(sorted(perm), perm):
        order[slot] = dim
    return (tuple(shape[i] for i in order), tuple(strides[i] for i in order), off)


def layout_split(l, axis, parts):
    shape, strides, off = l
    assert sum(parts) == shape[axis]
    out, at = [], 0
    for p in parts:
        out.append((shape[:axis] + (int(p),) + shape[axis + 1:], strides, off + at * strides[axis]))
        at += int(p)
    return out


# ---- mem-rearrange 0.1.0 ---------------------------------------------------------------------------------
def rearrange(dst, dl, src, sl, unit):
    """For every index of the common shape copy `unit` bytes src[...] -> dst[...]; other dst bytes untouched."""
    assert dl[0] == sl[0], "ShapeMismatch"
    shape = dl[0]
    if any(d == 0 for d in shape):
        return
    didx = np.full((), dl[2], np.int64)
    sidx = np.full((), sl[2], np.int64)
    for n, ds, ss in zip(shape, dl[1], sl[1]):  # outer-product accumulate of the byte addresses
        i = np.arange(n, dtype=np.int64)
        didx = (didx[..., None] + i * ds)
        sidx = (sidx[..., None] + i * ss)
    b = np.arange(unit, dtype=np.int64)
    dst[(didx[..., None] + b).reshape(-1)] = src[(sidx[..., None] + b).reshape(-1)]


# ---- the operators as the reference composes them -----------------------------------------------------
def block_layout(ty, shape):
    """merge.rs:359-364."""
    be, bb = TYPE_SIZE[ty]
    shape = list(shape)
    shape[0] //= be
    return layout_contiguous_le(shape, bb), bb


def concat(axis, tensors):
    """merge.rs:288-325; tensors = [(ty, shape, uint8 array)] -> (ty, shape, uint8 array)."""
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
    be, _ = TYPE_SIZE[ty]
    whole, unit = block_layout(ty, shape)
    parts = [s[axis] // be if axis == 0 else s[axis] for _, s, _ in tensors]
    out = np.zeros(elements_to_bytes(ty, shape), np.uint8)
    for (t, s, d), view in zip(tensors, layout_split(whole, axis, parts)):
        rearrange(out, view, np.asarray(d, np.uint8).reshape(-1), block_layout(t, s)[0], unit)
    return ty, tuple(shape), out


def split(axis, tensor, parts):
    """merge.rs:327-357."""
    ty, shape, data = tensor
    if len(shape) == 1:
        axis = 0
    assert shape[axis] == sum(parts)
    be, _ = TYPE_SIZE[ty]
    whole, unit = block_layout(ty, shape)
    data = np.asarray(data, np.uint8).reshape(-1)
    outs = []
    for p, view in zip(parts, layout_split(whole, axis, [p // be if axis == 0 else p for p in parts])):
        s = list(shape)
        s[axis] = p
        out = np.zeros(elements_to_bytes(ty, s), np.uint8)
        rearrange(out, block_layout(ty, s)[0], data, view, unit)
        outs.append((ty, tuple(s), out))
    return outs


def permute_qk(tensor, nh):
    """permute_qk.rs:46-69."""
    ty, shape, data = tensor
    if len(shape) == 1:
        c, r = 1, shape[0]
    else:
        c, r = shape
    c = elements_to_bytes(ty, [c])
    src = layout_transpose(layout_tile_le(layout_contiguous_le([c, r], 1), 1, (r // nh // 2, 2, nh)), (2, 1))
    dst = layout_contiguous_le(src[0], 1)
    out = np.zeros(c * r, np.uint8)
    rearrange(out, dst, np.asarray(data, np.uint8).reshape(-1), src, 1)
    return ty, tuple(shape), out


# ---- the same operators written directly from their meaning (cross-check) -------------------------------
def _as_rows(ty, shape, data):
    """numpy view [slower dims..., row bytes] (C order: the last axis is ggml's ne[0])."""
    be, bb = TYPE_SIZE[ty]
    if len(shape) == 1:
        return np.asarray(data, np.uint8).reshape(shape[0] // be, bb)  # 1-D: "rows" are blocks
    return np.asarray(data, np.uint8).reshape(tuple(reversed(shape[1:])) + (shape[0] // be * bb,))


def concat_direct(axis, tensors):
    ty, nd = tensors[0][0], len(tensors[0][1])
    if nd == 1:
        out = np.concatenate([_as_rows(t, s, d) for t, s, d in tensors], axis=0)
        return ty, (sum(s[0] for _, s, _ in tensors),), out.reshape(-1)
    assert axis >= 1
    out = np.concatenate([_as_rows(t, s, d) for t, s, d in tensors], axis=nd - 1 - axis)
    shape = list(tensors[0][1])
    shape[axis] = sum(s[axis] for _, s, _ in tensors)
    return ty, tuple(shape), np.ascontiguousarray(out).reshape(-1)


def split_direct(axis, tensor, parts):
    ty, shape, data = tensor
    be, _ = TYPE_SIZE[ty]
    rows = _as_rows(ty, shape, data)
    outs, at = [], 0
    for p in parts:
        s = list(shape)
        if len(shape) == 1:
            piece = rows[at // be:(at + p) // be]
            s[0] = p
        else:
            ax = len(shape) - 1 - axis
            piece = np.take(rows, np.arange(at, at + p), axis=ax)
            s[axis] = p
        outs.append((ty, tuple(s), np.ascontiguousarray(piece).reshape(-1)))
        at += p
    return outs


def permute_qk_direct(tensor, nh):
    """Within each head the first and second half of the rows are interleaved: out[2*i + j] = in[j*half + i]
    (the HF -> GGUF rotary permutation of llama.cpp's convert_hf_to_gguf.py `permute`)."""
    ty, shape, data = tensor
    r = shape[0] if len(shape) == 1 else shape[1]
    row_bytes = elements_to_bytes(ty, [1 if len(shape) == 1 else shape[0]])
    w = np.asarray(data, np.uint8).reshape(nh, 2, r // nh // 2, row_bytes)
    return ty, tuple(shape), np.ascontiguousarray(w.swapaxes(1, 2)).reshape(-1)


# ---- tensor-list level: names, grouping, order ----------------------------------------------------------
_MERGE = ("attn_q", "attn_k", "attn_v", "ffn_gate", "ffn_up", "ffn_gate_exps", "ffn_up_exps")


def _match(name, alts):
    for wb in ("weight", "bias"):
        if name.endswith("." + wb):
            stem = name[:-len(wb) - 1]
            for a in alts:
                if stem.endswith(a):
                    return stem[:-len(a)], a, wb
    return None


def merge_linear(tensors):
    """merge.rs:22-39 + collectors.  tensors: ordered list of (name, (ty, shape, data))."""
    out, groups = [], {}
    for name, t in tensors:
        m = _match(name, _MERGE)
        if not m:
            out.append((name, t))
            continue
        pre, which, wb = m
        layer = "attn" if which.startswith("attn") else "moe" if which.endswith("_exps") else "ffn"
        idx = {"attn_q": 0, "attn_k": 1, "attn_v": 2, "ffn_gate": 0, "ffn_up": 1, "ffn_gate_exps": 0, "ffn_up_exps": 1}[which]
        g = groups.setdefault((pre, layer, wb), {})
        g[idx] = (name, t)
        if len(g) == (3 if layer == "attn" else 2):
            parts = [g[i][1] for i in sorted(g)]
            del groups[(pre, layer, wb)]
            new = {"attn": "attn_qkv", "ffn": "ffn_gate_up", "moe": "ffn_gate_up_exps"}[layer]
            if layer == "attn":
                rows = [p[1][-1] if len(p[1]) <= 2 else None for p in parts]
                assert rows[0] % rows[1] == 0 and rows[0] >= rows[1] and rows[1] == rows[2]
            out.append((f"{pre}{new}.{wb}", concat(1, parts)))
    for g in groups.values():  # incomplete groups stay unmerged, moved to the end
        for i in sorted(g):
            out.append(g[i])
    return out


def split_linear(tensors, nh, nkvh):
    """merge.rs:40-81."""
    out = []
    for name, t in tensors:
        m = _match(name, ("attn_qkv", "ffn_gate_up"))
        if not m:
            out.append((name, t))
            continue
        pre, which, wb = m
        if which == "attn_qkv":
            r = t[1][-1] if len(t[1]) <= 2 else None
            dh = r // (nh + 2 * nkvh)
            for n, p in zip(("attn_q", "attn_k", "attn_v"), split(1, t, [nh * dh, nkvh * dh, nkvh * dh])):
                out.append((f"{pre}{n}.{wb}", p))
        else:
            r = t[1][1] // 2
            for n, p in zip(("ffn_gate", "ffn_up"), split(1, t, [r, r])):
                out.append((f"{pre}{n}.{wb}", p))
    return out


def permute_qk_all(tensors, nh, nkvh):
    """permute_qk.rs:11-44."""
    out = []
    for name, t in tensors:
        m = _match(name, ("attn_qkv", "attn_q", "attn_k"))
        if m:
            which = m[1]
            if which == "attn_q":
                t = permute_qk(t, nh)
            elif which == "attn_k":
                t = permute_qk(t, nkvh)
            else:
                r = t[1][-1]
                dh = r // (nh + 2 * nkvh)
                q, k, v = split(1, t, [nh * dh, nkvh * dh, nkvh * dh])
                t = concat(1, [permute_qk(q, nh), permute_qk(k, nkvh), v])
        out.append((name, t))
    return out
