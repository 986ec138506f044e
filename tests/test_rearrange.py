"""Block-granular rearrangement: merge-linear / split-linear / permute-qk (xtask/src/utils/operator/merge.rs,
permute_qk.rs) — oracle self-consistency and host logic on the CPU, `ggq_rearrange[_device]` and the convert
steps against the oracle on the GPU."""
import ctypes

import numpy as np
import pytest

from data import F16, F32, gaussian, to_fdt
from gguf_util import STRING, U32, U64, read_gguf, write_gguf

Q4_0, Q8_0, Q4K, Q6K, BF16 = 2, 8, 12, 14, 30


@pytest.fixture(scope="module")
def R():
    from oracle import rearrange
    return rearrange


def rand_tensor(R, ty, shape, seed):
    rng = np.random.default_rng(seed)
    return (ty, tuple(shape), rng.integers(0, 256, R.elements_to_bytes(ty, list(shape)), dtype=np.uint8))


# ---------------------------------------------------------------- oracle: the two formulations agree
@pytest.mark.parametrize("ty,shape,nh", [(F16, (64, 48), 4), (F32, (32, 64), 8), (Q4_0, (64, 32), 2), (Q8_0, (96, 24), 3),
                                         (Q4K, (512, 16), 4), (Q6K, (256, 8), 1), (F32, (64,), 4), (BF16, (8, 128), 16)])
def test_oracle_permute_qk_two_ways(R, ty, shape, nh):
    t = rand_tensor(R, ty, shape, 1)
    a, b = R.permute_qk(t, nh), R.permute_qk_direct(t, nh)
    assert a[1] == b[1] == tuple(shape) and np.array_equal(a[2], b[2])
    assert sorted(a[2].tolist()) == sorted(t[2].tolist())  # a permutation of the bytes
    if len(shape) == 2 and ty == F16:  # llama.cpp convert_hf_to_gguf.py `permute`
        w = t[2].view(np.float16).reshape(shape[1], shape[0])
        hf = w.reshape(nh, 2, shape[1] // nh // 2, shape[0]).swapaxes(1, 2).reshape(shape[1], shape[0])
        assert np.array_equal(hf.view(np.uint8).reshape(-1), a[2])


@pytest.mark.parametrize("ty,shapes,axis", [
    (F16, [(64, 32), (64, 8), (64, 8)], 1), (Q8_0, [(64, 6, 3), (64, 10, 3)], 1), (F32, [(32,), (8,), (8,)], 1),
    (Q4_0, [(64,), (32,)], 1), (Q4K, [(256, 4, 2), (256, 4, 2)], 1), (F16, [(16, 3, 2, 2), (16, 5, 2, 2)], 1),
    (F16, [(16, 3, 4), (16, 3, 2)], 2),
])
def test_oracle_concat_split_two_ways(R, ty, shapes, axis):
    ts = [rand_tensor(R, ty, s, 10 + i) for i, s in enumerate(shapes)]
    a, b = R.concat(axis, ts), R.concat_direct(axis, ts)
    assert a[1] == b[1] and np.array_equal(a[2], b[2])
    ax = 0 if len(shapes[0]) == 1 else axis
    parts = [s[ax] for s in shapes]
    back, back2 = R.split(axis, a, parts), R.split_direct(axis, a, parts)
    for t, x, y in zip(ts, back, back2):
        assert x[1] == y[1] == t[1]
        assert np.array_equal(x[2], t[2]) and np.array_equal(y[2], t[2])


def test_oracle_layout_algebra_known_answers(R):
    """The documented examples of ndarray-layout's tile_le / transpose."""
    l = ((2, 3, 6), (18, 6, 1), 0)
    assert R.layout_tile_le(l, 2, (2, 3)) == ((2, 3, 2, 3), (18, 6, 1, 2), 0)
    assert R.layout_transpose(((2, 3, 4), (12, 4, 1), 0), (1, 0)) == ((3, 2, 4), (4, 12, 1), 0)
    assert R.layout_contiguous_le([4, 3, 2], 2) == ((4, 3, 2), (2, 8, 24), 0)
    parts = R.layout_split(((4, 6), (2, 8), 5), 1, [1, 2, 3])
    assert [p[0] for p in parts] == [(4, 1), (4, 2), (4, 3)] and [p[2] for p in parts] == [5, 13, 29]


def test_product_layout_mirror_matches_oracle_algebra(R, ggq):
    """gguf_b200.rearrange.ArrayLayout (what feeds ggq_layout) against the oracle's independent restatement."""
    from gguf_b200.rearrange import ArrayLayout, permute_qk_layouts
    a = ArrayLayout.new_contiguous([128, 64], 1).tile_le(1, [8, 2, 4]).transpose([2, 1])
    o = R.layout_transpose(R.layout_tile_le(R.layout_contiguous_le([128, 64], 1), 1, (8, 2, 4)), (2, 1))
    assert (tuple(a.shape), tuple(a.strides), a.offset) == o
    dl, sl, unit = permute_qk_layouts(Q8_0, (64, 64), 4)
    assert unit == 1 and sl.shape == [68, 8, 2, 4][:1] + [2, 8, 4] and dl.strides == [1, 68, 136, 1088]
    sp = ArrayLayout.new_contiguous([4, 6, 2], 34).split(1, [2, 4])
    so = R.layout_split(R.layout_contiguous_le([4, 6, 2], 34), 1, [2, 4])
    assert [(tuple(x.shape), tuple(x.strides), x.offset) for x in sp] == so


def test_rearrange_validation_needs_no_gpu(ggq):
    """Shape / ndim mismatch = mem-rearrange's SchemeError::ShapeMismatch; rejected before any CUDA call."""
    from gguf_b200._lib import lib
    from gguf_b200.rearrange import ArrayLayout
    buf = np.zeros(64, np.uint8)
    a, b = ArrayLayout.new_contiguous([4, 4], 2).c(), ArrayLayout.new_contiguous([4, 3], 2).c()
    c3 = ArrayLayout.new_contiguous([4, 4, 1], 2).c()
    p = buf.ctypes.data
    assert lib().ggq_rearrange(p, ctypes.byref(a), p, ctypes.byref(b), 2) == 2
    assert lib().ggq_rearrange(p, ctypes.byref(a), p, ctypes.byref(c3), 2) == 2
    assert lib().ggq_rearrange(p, ctypes.byref(a), p, ctypes.byref(a), 0) == -3
    assert lib().ggq_rearrange_device(p, ctypes.byref(a), p, ctypes.byref(b), 2, None) == 2
    bad = ArrayLayout.new_contiguous([4, 4], 2).c()
    bad.ndim = 5
    assert lib().ggq_rearrange(p, ctypes.byref(bad), p, ctypes.byref(bad), 2) == -3
    zero = ArrayLayout([4, 4], [2, 0]).c()
    assert lib().ggq_rearrange(p, ctypes.byref(zero), p, ctypes.byref(a), 2) == -3
    empty = ArrayLayout.new_contiguous([4, 0], 2).c()
    assert lib().ggq_rearrange(p, ctypes.byref(empty), p, ctypes.byref(empty), 2) == 0  # nothing to move


# ---------------------------------------------------------------- convert planner on the CPU (--no-data)
def qkv_model(R, dtype=F16, hidden=128, nh=4, nkvh=2, ffn=192, layers=2, experts=0, bias=True, seed=0):
    """(name, shape, type, bytes) of a small llama-like model with separate q/k/v and gate/up."""
    ts = []
    dh = hidden // nh

    def add(name, shape, ty=dtype):
        nonlocal seed
        seed += 1
        x = gaussian(int(np.prod(shape)), seed)
        ts.append((name, tuple(shape), ty, to_fdt(x, ty).tobytes()))
    add("token_embd.weight", (hidden, 96))
    for l in range(layers):
        add(f"blk.{l}.attn_norm.weight", (hidden,), F32)
        add(f"blk.{l}.attn_q.weight", (hidden, nh * dh))
        add(f"blk.{l}.attn_k.weight", (hidden, nkvh * dh))
        if bias:
            add(f"blk.{l}.attn_q.bias", (nh * dh,), F32)
            add(f"blk.{l}.attn_k.bias", (nkvh * dh,), F32)
        add(f"blk.{l}.attn_v.weight", (hidden, nkvh * dh))
        if bias:
            add(f"blk.{l}.attn_v.bias", (nkvh * dh,), F32)
        add(f"blk.{l}.attn_output.weight", (nh * dh, hidden))
        if experts:
            add(f"blk.{l}.ffn_gate_exps.weight", (hidden, ffn, experts))
            add(f"blk.{l}.ffn_down_exps.weight", (ffn, hidden, experts))
            add(f"blk.{l}.ffn_up_exps.weight", (hidden, ffn, experts))
        else:
            add(f"blk.{l}.ffn_gate.weight", (hidden, ffn))
            add(f"blk.{l}.ffn_down.weight", (ffn, hidden))
            add(f"blk.{l}.ffn_up.weight", (hidden, ffn))
    add("blk.7.attn_q.weight", (hidden, nh * dh))  # an incomplete group: stays unmerged, moves to the end
    add("output.weight", (hidden, 96))
    return ts


def model_kvs(nh=4, nkvh=2, kv_type=U32):
    kvs = [("general.architecture", STRING, "llama"), ("llama.attention.head_count", kv_type, nh)]
    if nkvh is not None:
        kvs.append(("llama.attention.head_count_kv", U64, nkvh))
    return kvs


def run_oracle_steps(R, oracle, ts, steps, nh, nkvh):
    """The reference pipeline over (name, (ty, shape, data)) with the oracle's operators and codecs."""
    from test_convert import _expect
    cur = [(n, (ty, tuple(s), np.frombuffer(d, np.uint8))) for n, s, ty, d in ts]
    for st in steps:
        if st == "merge-linear":
            cur = R.merge_linear(cur)
        elif st in ("split-linear", "!merge-linear"):
            cur = R.split_linear(cur, nh, nkvh)
        elif st == "permute-qk":
            cur = R.permute_qk_all(cur, nh, nkvh)
        else:  # dict of cast rules
            want = _expect(oracle, [(n, t[1], t[0], t[2].tobytes()) for n, t in cur], [st])
            cur = [(n, (want[n][0], t[1], np.frombuffer(want[n][1], np.uint8))) for n, t in cur]
    return cur


def check_file(path, want):
    _, tensors, _, _ = read_gguf(path)
    assert list(tensors) == [n for n, _ in want]
    for n, (ty, shape, data) in want:
        got_shape, got_ty, got = tensors[n]
        assert tuple(got_shape) == tuple(shape) and got_ty == ty, n
        assert got == data.tobytes(), n


@pytest.mark.parametrize("steps,expect_names", [
    ("merge-linear", ["token_embd.weight", "blk.0.attn_norm.weight", "blk.0.attn_qkv.weight", "blk.0.attn_qkv.bias",
                      "blk.0.attn_output.weight", "blk.0.ffn_down.weight", "blk.0.ffn_gate_up.weight"]),
    ("merge-linear -> split-linear", ["token_embd.weight", "blk.0.attn_norm.weight", "blk.0.attn_q.weight", "blk.0.attn_k.weight",
                                      "blk.0.attn_v.weight", "blk.0.attn_q.bias", "blk.0.attn_k.bias", "blk.0.attn_v.bias"]),
])
def test_planner_names_shapes_order_without_gpu(ggq, R, tmp_path, steps, expect_names):
    """--no-data: header only, no device needed.  Order: a merged tensor stands where its LAST part stood."""
    from gguf_b200.convert import convert
    src, dst = tmp_path / "in.gguf", tmp_path / "out.gguf"
    ts = qkv_model(R)
    write_gguf(src, model_kvs(), ts)
    st = convert(src, dst, steps, no_data=True)
    buf = open(dst, "rb").read()
    import struct
    (nt,) = struct.unpack_from("<Q", buf, 8)
    want = run_oracle_steps(R, None, ts, [s.strip() for s in steps.split("->")], 4, 2)
    assert nt == len(want) == st["n_tensors"]
    # parse the infos by hand (no data region in a --no-data file)
    p = 24
    (nkv,) = struct.unpack_from("<Q", buf, 16)

    def rs():
        nonlocal p
        (n,) = struct.unpack_from("<Q", buf, p)
        s = buf[p + 8:p + 8 + n].decode()
        p += 8 + n
        return s
    for _ in range(nkv):
        rs()
        (ty,) = struct.unpack_from("<I", buf, p)
        p += 4
        if ty == STRING:
            rs()
        else:
            p += {U32: 4, U64: 8}[ty]
    got = []
    for _ in range(nt):
        name = rs()
        (nd,) = struct.unpack_from("<I", buf, p)
        shape = struct.unpack_from("<%dQ" % nd, buf, p + 4)
        ty, _off = struct.unpack_from("<IQ", buf, p + 4 + 8 * nd)
        p += 4 + 8 * nd + 12
        got.append((name, tuple(shape), ty))
    assert got == [(n, tuple(t[1]), t[0]) for n, t in want]
    assert [g[0] for g in got][:len(expect_names)] == expect_names
    assert got[-2][0] == "output.weight" and got[-1][0] == "blk.7.attn_q.weight" or "split" in steps


def test_step_errors_without_gpu(ggq, R, tmp_path):
    from gguf_b200.convert import convert
    src, dst = tmp_path / "in.gguf", tmp_path / "out.gguf"
    ts = qkv_model(R)
    write_gguf(src, [("general.architecture", STRING, "llama")], ts)
    with pytest.raises(ggq.GgqError) as e:  # merge.rs:41 `.unwrap()` on NotExist
        convert(src, dst, "permute-qk", no_data=True)
    assert "head_count" in str(e.value)
    write_gguf(src, model_kvs(nh=3), ts)    # 128 rows / (2*3): tile_le would panic
    with pytest.raises(ggq.GgqError):
        convert(src, dst, "permute-qk", no_data=True)
    write_gguf(src, model_kvs(nkvh=None), ts)  # head_count_kv missing -> nh (permute_qk.rs:14-18): k has 64 rows, 64 % 8 == 0
    convert(src, dst, "permute-qk", no_data=True)
    with pytest.raises(ggq.GgqError) as e:
        convert(src, dst, "sort", no_data=True)
    assert "implemented" in str(e.value)


# ---------------------------------------------------------------- GPU: ggq_rearrange[_device] vs the oracle
def dev_rearrange(ggq, dl, sl, unit, src_bytes, dst_init):
    import torch
    from gguf_b200.rearrange import rearrange_device
    s = torch.from_numpy(src_bytes.copy()).cuda()
    d = torch.from_numpy(dst_init.copy()).cuda()
    rearrange_device(d.data_ptr(), dl, s.data_ptr(), sl, unit, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d.cpu().numpy()


def as_product_layout(l):
    from gguf_b200.rearrange import ArrayLayout
    return ArrayLayout(l[0], l[1], l[2])


@pytest.mark.gpu
@pytest.mark.parametrize("ty,shape,nh", [(F16, (4096, 512), 4), (F32, (512, 256), 8), (Q4_0, (4096, 256), 2), (Q4_0, (96, 64), 4),
                                         (Q8_0, (4096, 128), 8), (Q4K, (4096, 64), 4), (Q6K, (768, 32), 2), (F32, (4096,), 32),
                                         (BF16, (8, 128), 16), (Q8_0, (32, 6), 3)])
def test_permute_qk_device_matches_oracle(ggq, R, ty, shape, nh):
    from gguf_b200.rearrange import permute_qk_layouts
    t = rand_tensor(R, ty, shape, 3)
    want = R.permute_qk(t, nh)[2]
    dl, sl, unit = permute_qk_layouts(ty, shape, nh)
    got = dev_rearrange(ggq, dl, sl, unit, t[2], np.zeros_like(t[2]))
    assert np.array_equal(got, want)
    host = ggq.permute_qk(t[2], ty, shape, nh)  # host-pointer entry
    assert np.array_equal(host, want)


@pytest.mark.gpu
@pytest.mark.parametrize("ty,shapes,axis", [
    (F16, [(1024, 256), (1024, 64), (1024, 64)], 1), (Q8_0, [(512, 48, 8), (512, 80, 8)], 1), (F32, [(256,), (64,), (64,)], 1),
    (Q4_0, [(96, 5, 3), (96, 7, 3)], 1), (Q4K, [(1024, 32, 4), (1024, 32, 4)], 1), (F16, [(16, 3, 2, 2), (16, 5, 2, 2)], 1),
    (Q6K, [(256, 3, 2), (256, 3, 5)], 2),
])
def test_concat_split_host_api_matches_oracle(ggq, R, ty, shapes, axis):
    ts = [rand_tensor(R, ty, s, 20 + i) for i, s in enumerate(shapes)]
    want = R.concat(axis, ts)
    got = ggq.concat(axis, ts)
    assert got[1] == want[1] and np.array_equal(got[2], want[2])
    ax = 0 if len(shapes[0]) == 1 else axis
    for g, t in zip(ggq.split(axis, got, [s[ax] for s in shapes]), ts):
        assert g[1] == t[1] and np.array_equal(g[2], t[2])


@pytest.mark.gpu
def test_rearrange_general_layouts_and_untouched_bytes(ggq, R):
    """Layouts the operators never build: 4 unmergeable dims, a transposing gather with 2-byte units, odd
    offsets (1-byte vectors), negative strides, grid.y overflow; bytes outside the layout keep their values."""
    rng = np.random.default_rng(5)
    cases = []
    # 4 dims, none mergeable (every dim padded): exercises the host-side peel
    cases.append((((3, 4, 5, 6), (2, 8, 40, 256), 6), ((3, 4, 5, 6), (720, 2, 144, 24), 0), 2))
    # element transpose of a 2-byte matrix
    cases.append((((64, 48), (2, 128), 0), ((64, 48), (96, 2), 0), 2))
    # odd pointers / run length: falls back to byte vectors
    cases.append((((33, 7), (1, 40), 3), ((33, 7), (1, 35), 1), 1))
    # reversed rows (negative source stride)
    cases.append((((32, 10), (1, 32), 0), ((32, 10), (1, -32), 9 * 32), 1))
    # 70 000 rows of 16 bytes swapped in pairs: middle dims exceed grid.y
    cases.append((((16, 2, 70000), (1, 16, 32), 0), ((16, 2, 70000), (1, -16, 32), 16), 1))
    for dl, sl, unit in cases:
        def span(l):
            lo = l[2] + sum((n - 1) * s for n, s in zip(l[0], l[1]) if s < 0)
            hi = l[2] + sum((n - 1) * s for n, s in zip(l[0], l[1]) if s > 0) + unit
            assert lo >= 0
            return hi
        src = rng.integers(0, 256, span(sl) + 5, dtype=np.uint8)
        dst0 = rng.integers(0, 256, span(dl) + 9, dtype=np.uint8)
        want = dst0.copy()
        R.rearrange(want, dl, src, sl, unit)
        got = dev_rearrange(ggq, as_product_layout(dl), as_product_layout(sl), unit, src, dst0)
        assert np.array_equal(got, want), (dl, sl)
        host = dst0.copy()
        ggq.rearrange(host, as_product_layout(dl), src, as_product_layout(sl), unit)
        assert np.array_equal(host, want), (dl, sl)


@pytest.mark.gpu
def test_rearrange_fuzz_random_layouts(ggq, R):
    """Random shapes (1-4 dims), units, padded / permuted destination layouts and arbitrary source layouts
    (permuted, padded, reversed, broadcast) through the device API against the oracle's index arithmetic."""
    import os
    rng = np.random.default_rng(99)
    n_cases = int(os.environ.get("GGQ_FUZZ_CASES", "300"))
    for case in range(n_cases):
        nd = int(rng.integers(1, 5))
        shape = [int(rng.choice([1, 2, 3, 4, 5, 7, 8, 16, 33])) for _ in range(nd)]
        unit = int(rng.choice([1, 2, 4, 6, 16, 18, 34]))

        def layout(allow_special):
            order = rng.permutation(nd)
            strides, mul = [0] * nd, unit
            for ax in order:
                pad = int(rng.choice([0, 0, 0, 1, 3])) * unit
                strides[ax] = mul
                mul = mul * shape[ax] + pad
            offset = int(rng.choice([0, 0, 1, 2, 5, 16]))
            if allow_special:
                for ax in range(nd):
                    k = rng.random()
                    if k < 0.15 and shape[ax] > 1:      # reversed axis
                        offset += (shape[ax] - 1) * strides[ax]
                        strides[ax] = -strides[ax]
                    elif k < 0.25:                        # broadcast axis
                        strides[ax] = 0
            return (tuple(shape), tuple(strides), offset)
        dl, sl = layout(False), layout(True)

        def span(l):
            hi = l[2] + sum((n - 1) * st for n, st in zip(l[0], l[1]) if st > 0) + unit
            lo = l[2] + sum((n - 1) * st for n, st in zip(l[0], l[1]) if st < 0)
            assert lo >= 0
            return hi
        src = rng.integers(0, 256, span(sl) + 3, dtype=np.uint8)
        dst0 = rng.integers(0, 256, span(dl) + 3, dtype=np.uint8)
        want = dst0.copy()
        R.rearrange(want, dl, src, sl, unit)
        got = dev_rearrange(ggq, as_product_layout(dl), as_product_layout(sl), unit, src, dst0)
        assert np.array_equal(got, want), (case, dl, sl, unit)


@pytest.mark.gpu
def test_rearrange_launches_are_counted_and_async(ggq, R):
    import torch
    from gguf_b200._lib import lib
    from gguf_b200.rearrange import permute_qk_layouts, rearrange_device
    t = rand_tensor(R, F16, (4096, 4096), 9)
    s = torch.from_numpy(t[2]).cuda()
    d = torch.empty_like(s)
    dl, sl, unit = permute_qk_layouts(F16, (4096, 4096), 32)
    n0 = lib().ggq_launch_count()
    rearrange_device(d.data_ptr(), dl, s.data_ptr(), sl, unit, torch.cuda.current_stream().cuda_stream)
    assert lib().ggq_launch_count() == n0 + 1  # one kernel for the whole tensor
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy(), R.permute_qk(t, 32)[2])


# ---------------------------------------------------------------- GPU: convert steps end to end
@pytest.mark.gpu
@pytest.mark.parametrize("steps,ops", [
    ("permute-qk", ["permute-qk"]),
    ("merge-linear", ["merge-linear"]),
    ("merge-linear -> permute-qk", ["merge-linear", "permute-qk"]),
    ("merge-linear -> split-linear", ["merge-linear", "split-linear"]),
    ("merge-linear -> cast:linear:q8_0 embd:q4_0", ["merge-linear", dict(linear=8, embd=2)]),
    ("cast:linear:f32 -> permute-qk -> merge-linear -> cast:linear:q8_0", [dict(linear=0), "permute-qk", "merge-linear", dict(linear=8)]),
    ("permute-qk -> cast:linear:q4_0 -> !merge-linear", ["permute-qk", dict(linear=2), "split-linear"]),
    ("merge-linear -> cast:linear:Q4K -> split-linear -> cast:linear:f16", ["merge-linear", dict(linear=12), "split-linear", dict(linear=1)]),
])
@pytest.mark.parametrize("experts", [0, 4])
def test_convert_rearrange_steps_match_oracle(ggq, oracle, R, tmp_path, steps, ops, experts):
    from gguf_b200.convert import convert
    if experts and any(isinstance(o, dict) and 12 in o.values() for o in ops):
        pytest.skip("expert rows of 192 are not a multiple of the Q4_K super-block")
    src, dst = tmp_path / "in.gguf", tmp_path / "out.gguf"
    kq = any(isinstance(o, dict) and 12 in o.values() for o in ops)
    hidden = 256 if kq else 128
    # no biases with Q4_K: a 128-element k/v bias cannot be split out of a 256-element super-block (the reference panics too)
    ts = qkv_model(R, hidden=hidden, experts=experts, seed=40, ffn=256 if kq else 192, bias=not kq)
    write_gguf(src, model_kvs(), ts)
    st = convert(src, dst, steps)
    want = run_oracle_steps(R, oracle, ts, ops, 4, 2)
    check_file(dst, want)
    assert st["n_tensors"] == len(want)
    if "permute-qk" in ops or (experts and ops == ["merge-linear"]):
        assert st["n_rearranged_tensors"] > 0  # rows really moved: device-resident evaluation


@pytest.mark.gpu
def test_convert_split_of_premerged_file_and_quantized_permute(ggq, oracle, R, tmp_path):
    """A file that already holds attn_qkv / ffn_gate_up in Q8_0: split-linear reads sub-ranges, permute-qk moves
    34-byte blocks by rows; sharded output (-t 5)."""
    from gguf_b200.convert import convert
    ts = qkv_model(R, seed=60, bias=False)  # a quantized 1-D bias cannot be permuted (tensor.rs:92 asserts in the reference)
    merged = run_oracle_steps(R, oracle, ts, ["merge-linear", dict(linear=8)], 4, 2)
    src = tmp_path / "merged.gguf"
    write_gguf(src, model_kvs(), [(n, t[1], t[0], t[2].tobytes()) for n, t in merged])
    for steps, ops in [("split-linear", ["split-linear"]), ("permute-qk", ["permute-qk"]), ("split-linear -> permute-qk -> cast:linear:f16", ["split-linear", "permute-qk", dict(linear=1)])]:
        dst = tmp_path / "o.gguf"
        convert(src, dst, steps)
        want = run_oracle_steps(R, oracle, [(n, t[1], t[0], t[2].tobytes()) for n, t in merged], ops, 4, 2)
        check_file(dst, want)
    st = convert(src, tmp_path / "sh.gguf", "permute-qk", max_tensors=5)
    assert st["n_out_files"] > 1
    want = dict(run_oracle_steps(R, oracle, [(n, t[1], t[0], t[2].tobytes()) for n, t in merged], ["permute-qk"], 4, 2))
    seen = {}
    for i in range(st["n_out_files"]):
        _, tensors, _, _ = read_gguf(tmp_path / f"sh-{i + 1:05d}-of-{st['n_out_files']:05d}.gguf")
        seen.update(tensors)
    assert set(seen) == set(want)
    for n, (ty, shape, data) in want.items():
        assert seen[n][2] == data.tobytes(), n


@pytest.mark.gpu
def test_convert_resident_budget_serialises_big_tensors(ggq, oracle, R, tmp_path, monkeypatch):
    """With a 1 MB device-memory budget every device-resident tensor is larger than the budget and runs alone;
    the output must not change."""
    from gguf_b200.convert import convert
    monkeypatch.setenv("GGQ_RESIDENT_BUDGET_MB", "1")
    src, dst = tmp_path / "in.gguf", tmp_path / "out.gguf"
    ts = qkv_model(R, hidden=512, ffn=768, experts=4, seed=80, layers=3)
    write_gguf(src, model_kvs(), ts)
    st = convert(src, dst, "permute-qk -> merge-linear -> cast:linear:q8_0")
    want = run_oracle_steps(R, oracle, ts, ["permute-qk", "merge-linear", dict(linear=8)], 4, 2)
    check_file(dst, want)
    assert st["n_rearranged_tensors"] >= 6
