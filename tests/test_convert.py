"""Whole-file conversion (`ggq_convert_gguf`, the xtask `convert --steps cast:...` path)."""
import os
import struct

import numpy as np
import pytest

from data import F16, F32, gaussian, to_fdt
from gguf_util import ARRAY, STRING, U32, read_gguf, write_gguf

Q4_0, Q8_0, Q4K, Q6K = 2, 8, 12, 14


def llama_like(rng_seed=0, hidden=256, ffn=512, vocab=320, layers=2, dtype=F16, big=None):
    """(name, shape, type, bytes) for a tiny llama-architecture model; shapes are ggml order (ne0 first)."""
    ts = []

    def add(name, shape, ty=dtype, seed=[rng_seed]):
        seed[0] += 1
        n = int(np.prod(shape))
        x = gaussian(n, seed[0])
        ts.append((name, tuple(shape), ty, to_fdt(x, ty).tobytes()))
    add("token_embd.weight", (hidden, vocab))
    for l in range(layers):
        add(f"blk.{l}.attn_norm.weight", (hidden,), F32)
        add(f"blk.{l}.attn_q.weight", (hidden, hidden))
        add(f"blk.{l}.attn_q.bias", (hidden,), F32)
        add(f"blk.{l}.ffn_down.weight", (ffn, hidden))
        add(f"blk.{l}.ffn_up.weight", (hidden, ffn))
    add("output_norm.weight", (hidden,), F32)
    add("output.weight", (hidden, vocab))
    add("rope_freqs.weight", (64,), F32)  # 1-D, not norm/bias -> "else"
    if big:
        add("blk.9.ffn_gate.weight", big)
    return ts


KVS = [("general.architecture", STRING, "llama"), ("general.name", STRING, "tiny"), ("general.alignment", U32, 64),
       ("llama.block_count", U32, 2), ("tokenizer.ggml.tokens", ARRAY, (STRING, ["a", "bc", "def"])),
       ("split.count", U32, 1)]


def test_copy_only_layout_matches_reference_writer(ggq, tmp_path):
    """No cast step => no GPU needed.  Layout: header, general.alignment first, KVs in order minus split.*,
    infos, alignment-padded data (xtask/src/utils/write.rs:70-90, read.rs:37-39)."""
    from gguf_b200.convert import convert
    src, dst = tmp_path / "in.gguf", tmp_path / "out.gguf"
    ts = llama_like()
    write_gguf(src, KVS, ts, alignment=64)
    st = convert(src, dst, "")
    kvs, tensors, alignment, size = read_gguf(dst)
    assert alignment == 64 and st["bytes_out"] == size and st["n_cast_tensors"] == 0
    assert [k for k, _, _ in kvs] == ["general.alignment", "general.architecture", "general.name", "llama.block_count", "tokenizer.ggml.tokens"]
    assert list(tensors) == [t[0] for t in ts]
    for name, shape, ty, data in ts:
        assert tensors[name] == (shape, ty, data)
    # idempotent: converting the output again reproduces it byte for byte
    dst2 = tmp_path / "out2.gguf"
    convert(dst, dst2, "")
    assert open(dst, "rb").read() == open(dst2, "rb").read()


def test_types_without_a_codec_pass_through_untouched(ggq, tmp_path):
    """A file holding tensors of types this library has no codec for (IQ*, sizes = size_of the reference's
    structs, ggus/src/tensor.rs:102-144) still converts: untouched tensors are borrowed bytes in the
    reference (utils/mod.rs:104-138); only a cast that reads or targets such a type is refused."""
    from gguf_b200.convert import convert
    from gguf_b200 import GgqError
    from gguf_util import TYPE_SIZE
    rng = np.random.default_rng(5)
    ts = llama_like(layers=1)
    for ty in (16, 17, 18, 19, 20, 21, 22, 23, 29):
        e, b = TYPE_SIZE[ty]
        ts.append((f"blk.0.iq{ty}.weight", (e * 2, 3), ty, rng.integers(0, 256, 2 * 3 * b, dtype=np.uint8).tobytes()))
    src, dst = tmp_path / "in.gguf", tmp_path / "out.gguf"
    write_gguf(src, KVS, ts, alignment=64)
    st = convert(src, dst, "")
    _, tensors, _, _ = read_gguf(dst)
    for name, shape, ty, data in ts:
        assert tensors[name] == (shape, ty, data), name
    assert st["n_tensors"] == len(ts)
    with pytest.raises(GgqError) as ei:          # "linear" matches the 2-D IQ tensors: refused by name, nothing written
        convert(src, tmp_path / "bad.gguf", "cast:linear:f16")
    assert "unsupported type" in str(ei.value)


@pytest.mark.parametrize("mutate,msg", [
    (dict(magic=b"GGUX"), "MagicMismatch"), (dict(version=2), "VersionNotSupport"),
])
def test_parse_errors(ggq, tmp_path, mutate, msg):
    """ggus/src/file.rs:148-444 error cases."""
    from gguf_b200.convert import convert
    src = tmp_path / "bad.gguf"
    write_gguf(src, KVS, llama_like(), alignment=64, **mutate)
    with pytest.raises(ggq.GgqError) as e:
        convert(src, tmp_path / "o.gguf", "")
    assert msg in str(e.value)


def test_parse_errors_duplicates_and_truncation(ggq, tmp_path):
    from gguf_b200.convert import convert
    src = tmp_path / "dup.gguf"
    write_gguf(src, KVS + [("general.name", STRING, "again")], llama_like(), alignment=64)
    with pytest.raises(ggq.GgqError) as e:
        convert(src, tmp_path / "o.gguf", "")
    assert "DuplicateMetaKey" in str(e.value)
    ts = llama_like()
    write_gguf(src, KVS, ts + [ts[0]], alignment=64)
    with pytest.raises(ggq.GgqError) as e:
        convert(src, tmp_path / "o.gguf", "")
    assert "DuplicateTensorName" in str(e.value)
    write_gguf(src, KVS, ts, alignment=64)
    blob = open(src, "rb").read()
    open(src, "wb").write(blob[:-100])
    with pytest.raises(ggq.GgqError) as e:
        convert(src, tmp_path / "o.gguf", "")
    assert "Eos" in str(e.value)
    write_gguf(src, [("general.architecture", STRING, "mamba")], ts)
    with pytest.raises(ggq.GgqError) as e:
        convert(src, tmp_path / "o.gguf", "cast:linear:q8_0")
    assert "Unsupported architecture" in str(e.value)
    with pytest.raises(ggq.GgqError):
        convert(src, tmp_path / "o.gguf", "sort:")


def _expect(oracle, ts, rules_per_step):
    """Apply cast.rs:28-90 with the oracle: dict name -> (type, bytes)."""
    out = {}
    for name, shape, ty, data in ts:
        if name in ("token_embd.weight", "output.weight"):
            cls = "embd"
        elif name.endswith("_norm.weight") or name.endswith("_norm.bias"):
            cls = "norm"
        elif len(shape) > 1 or name.endswith(".bias"):
            cls = "linear"
        else:
            cls = "else"
        cur_ty, cur = ty, np.frombuffer(data, np.uint8)
        for rules in rules_per_step:
            to = rules.get(cls)
            if to is None or to == cur_ty:
                continue
            # cast.rs:93-138: float -> anything = quantize; block -> float = dequantize; block -> block via F32
            def as_float(t, b):
                return b.view(np.float32) if t == 0 else b.view(np.uint16)
            if cur_ty in (0, 1, 30) and to in (0,):
                cur = oracle.dequantize(cur_ty, 0, cur).view(np.uint8)
            elif cur_ty in (0, 1, 30):
                cur = oracle.quantize(to, cur_ty, as_float(cur_ty, cur)).view(np.uint8)
            elif to in (0, 1, 30):
                cur = oracle.dequantize(cur_ty, to, cur).view(np.uint8)
            else:
                cur = oracle.quantize(to, 0, oracle.dequantize(cur_ty, 0, cur)).view(np.uint8)
            cur_ty = to
        out[name] = (cur_ty, cur.tobytes())
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("steps,rules", [
    ("cast:linear:q8_0 embd:q4_0", [dict(linear=8, embd=2)]),
    ("cast:linear:q8_0 embd:q8_0 -> cast:linear:f32 embd:f32 -> cast:linear:f16 embd:f16", [dict(linear=8, embd=8), dict(linear=0, embd=0), dict(linear=1, embd=1)]),
    ("cast:linear:Q4K embd:q6k norm:f16 else:bf16", [dict(linear=12, embd=14, norm=1, **{"else": 30})]),
    ("cast:linear:q5_1 -> cast:linear:q4_0", [dict(linear=7), dict(linear=2)]),
])
def test_convert_matches_oracle(ggq, oracle, tmp_path, steps, rules):
    from gguf_b200.convert import convert
    src, dst = tmp_path / "in.gguf", tmp_path / "out.gguf"
    ts = llama_like(rng_seed=5)
    write_gguf(src, KVS, ts, alignment=64)
    st = convert(src, dst, steps)
    _, tensors, _, size = read_gguf(dst)
    want = _expect(oracle, ts, rules)
    for name, shape, ty, data in ts:
        got_shape, got_ty, got = tensors[name]
        assert got_shape == shape and got_ty == want[name][0], name
        assert got == want[name][1], name
    assert st["bytes_out"] == size and st["n_cast_tensors"] > 0


@pytest.mark.gpu
def test_convert_multichunk_tensor_and_gguf_py_reader(ggq, oracle, tmp_path):
    """A tensor larger than one 8 Mi-element pipeline chunk; the output is also readable by gguf-py."""
    from gguf_b200.convert import convert
    src, dst = tmp_path / "in.gguf", tmp_path / "out.gguf"
    ts = llama_like(rng_seed=9, big=(4096, 2304))  # 9.4 M elements
    write_gguf(src, [kv for kv in KVS if kv[0] != "general.alignment"], ts)
    convert(src, dst, "cast:linear:q8_0 embd:q8_0")
    _, tensors, alignment, _ = read_gguf(dst)
    assert alignment == 32
    want = _expect(oracle, ts, [dict(linear=8, embd=8)])
    for name, *_ in ts:
        assert tensors[name][2] == want[name][1], name
    gr = pytest.importorskip("gguf.gguf_reader")
    r = gr.GGUFReader(str(dst))
    by_name = {t.name: t for t in r.tensors}
    assert set(by_name) == {t[0] for t in ts}
    t = by_name["blk.9.ffn_gate.weight"]
    assert int(t.tensor_type) == 8 and bytes(t.data.tobytes()) == want["blk.9.ffn_gate.weight"][1]


@pytest.mark.gpu
def test_cast_api_pairs_the_reference_leaves_unimplemented(ggq, oracle):
    """cast.rs:132-136: Q8_0 -> F16 is todo!(), Q4_0 -> anything recurses forever; ggq_cast implements them."""
    import ctypes
    from gguf_b200._lib import lib
    x = gaussian(32 * 1000, 3)
    q8 = oracle.quantize(8, 0, x)
    out = np.empty(x.size, np.uint16)
    chain = (ctypes.c_uint32 * 2)(8, 1)
    assert lib().ggq_cast(chain, 2, out.ctypes.data, q8.ctypes.data, x.size) == 0
    assert np.array_equal(out, oracle.dequantize(8, 1, q8))
    q4 = oracle.quantize(2, 0, x)
    out2 = np.empty(x.size // 32 * 34, np.uint8)
    chain = (ctypes.c_uint32 * 2)(2, 8)
    assert lib().ggq_cast(chain, 2, out2.ctypes.data, q4.ctypes.data, x.size) == 0
    assert np.array_equal(out2, oracle.quantize(8, 0, oracle.dequantize(2, 0, q4)))
    assert lib().ggq_cast(chain, 2, out2.ctypes.data, q4.ctypes.data, 31) == 1  # Indivisible


def _plan_reference(kvs, tensors, alignment, max_tensors=None, max_bytes=None, no_tensor_first=False):
    """Independent restatement of xtask/src/utils/write.rs:23-51 + ggus/src/write/simulator.rs:75-96."""
    from gguf_util import kv_bytes
    pad = lambda p: (alignment - p % alignment) % alignment
    kvb = sum(len(kv_bytes(k, t, v)) for k, t, v in kvs if k != "general.alignment" and not k.startswith("split."))
    akv = 8 + len("general.alignment") + 4 + 4
    info = lambda name, shape: 8 + len(name.encode()) + 4 + 8 * len(shape) + 4 + 8
    max_tensors = max_tensors or 1 << 62
    max_bytes = max_bytes or 1 << 62

    class Sim:
        def __init__(self, kvb):
            self.w, self.data = 24 + akv + kvb, []
        def write(self, t):
            self.w += info(t[0], t[1]); self.data.append(len(t[3]))
        def total(self):
            tot = self.w
            for n in self.data:
                tot += pad(tot) + n
            return tot
    sim, shards = Sim(kvb), [[]]
    for t in tensors:
        if len(shards) == 1 and no_tensor_first:
            sim = Sim(0); sim.write(t); shards.append([t[0]]); continue
        sim.write(t)
        if len(shards[-1]) < max_tensors and sim.total() < max_bytes:
            shards[-1].append(t[0])
        else:
            sim = Sim(0); sim.write(t); shards.append([t[0]])
    return shards


@pytest.mark.parametrize("opts", [dict(max_tensors=3), dict(max_bytes="200K"), dict(max_bytes=150000, max_tensors=4),
                                  dict(no_tensor_first=True), dict(no_tensor_first=True, max_tensors=5)])
def test_output_sharding_follows_reference_planner(ggq, tmp_path, opts):
    """Copy-only (no GPU needed): shard contents, names, KV placement and round trip through a merge."""
    from gguf_b200.convert import convert, parse_mem_size
    src = tmp_path / "in.gguf"
    ts = llama_like(rng_seed=3)
    write_gguf(src, KVS, ts, alignment=64)
    st = convert(src, tmp_path / "model.gguf", "", **opts)
    want = _plan_reference(KVS, ts, 64, opts.get("max_tensors"), parse_mem_size(opts.get("max_bytes")), opts.get("no_tensor_first", False))
    n = len(want)
    assert st["n_out_files"] == n and n > 1
    paths = [tmp_path / f"model-{i + 1:05d}-of-{n:05d}.gguf" for i in range(n)]
    by_name = {t[0]: t for t in ts}
    for i, (path, names) in enumerate(zip(paths, want)):
        kvs, tensors, alignment, _ = read_gguf(path)
        assert alignment == 64 and list(tensors) == names
        assert [k for k, _, _ in kvs] == (["general.alignment", "general.architecture", "general.name", "llama.block_count",
                                           "tokenizer.ggml.tokens"] if i == 0 else ["general.alignment"])
        for name in names:
            assert tensors[name] == by_name[name][1:]
    # merging the shards back (Content::new over several files) reproduces the single-file conversion
    convert(src, tmp_path / "single.gguf", "")
    convert(paths, tmp_path / "merged.gguf", "")
    assert open(tmp_path / "single.gguf", "rb").read() == open(tmp_path / "merged.gguf", "rb").read()


def test_no_data_writes_infos_only(ggq, tmp_path):
    """--no-data (output.rs:22-24): header, KVs and tensor infos, no tensor bytes — how the reference's own
    fixture test-files/TinyLlama-1.1B-Chat-v1.0-F16.gguf was produced."""
    from gguf_b200.convert import convert
    import struct
    src, dst = tmp_path / "in.gguf", tmp_path / "nodata.gguf"
    ts = llama_like()
    write_gguf(src, KVS, ts, alignment=64)
    st = convert(src, dst, "cast:linear:q8_0", no_data=True)   # types change in the infos, nothing is computed
    blob = open(dst, "rb").read()
    assert st["bytes_out"] == len(blob) and st["n_cast_tensors"] == 0
    assert struct.unpack_from("<IQQ", blob, 4) == (3, len(ts), 5)
    full = tmp_path / "full.gguf"
    assert len(blob) < os.path.getsize(src) // 4
    assert blob.count(b"blk.0.attn_q.weight") == 1


def test_merge_rejects_duplicates_across_inputs(ggq, tmp_path):
    from gguf_b200.convert import convert
    a, b = tmp_path / "a.gguf", tmp_path / "b.gguf"
    ts = llama_like()
    write_gguf(a, KVS, ts[:4], alignment=64)
    write_gguf(b, [("general.alignment", U32, 32)], ts[3:6], alignment=32)
    with pytest.raises(ggq.GgqError) as e:
        convert([a, b], tmp_path / "o.gguf", "")
    assert "DuplicateTensorName" in str(e.value)
    write_gguf(b, [("general.alignment", U32, 32), ("split.no", U32, 1)], ts[4:6], alignment=32)
    st = convert([a, b], tmp_path / "o.gguf", "")
    kvs, tensors, alignment, _ = read_gguf(tmp_path / "o.gguf")
    assert alignment == 64 and list(tensors) == [t[0] for t in ts[:6]]   # alignment = max over inputs


@pytest.mark.gpu
def test_sharded_convert_with_cast_matches_oracle(ggq, oracle, tmp_path):
    from gguf_b200.convert import convert
    src = tmp_path / "in.gguf"
    ts = llama_like(rng_seed=21)
    write_gguf(src, KVS, ts, alignment=64)
    st = convert(src, tmp_path / "q.gguf", "cast:linear:q4k embd:q6k", max_tensors=4)
    n = st["n_out_files"]
    want = _expect(oracle, ts, [dict(linear=12, embd=14)])
    seen = {}
    for i in range(n):
        _, tensors, _, _ = read_gguf(tmp_path / f"q-{i + 1:05d}-of-{n:05d}.gguf")
        seen.update(tensors)
    assert list(seen) == [t[0] for t in ts]
    for name, (shape, ty, data) in seen.items():
        assert (ty, data) == want[name], name


@pytest.mark.gpu
def test_convert_leaves_the_callers_thread_state_alone(ggq, oracle, tmp_path):
    """ggq_convert_gguf must not pin the calling thread to a device (its workers are all spawned threads) nor
    change the thread's current CUDA device; a sharded slice call afterwards still shards and also restores it."""
    import torch
    from gguf_b200._lib import lib
    from gguf_b200.convert import convert
    ndev = lib().ggq_device_count()
    cur = ndev - 1
    torch.cuda.set_device(cur)
    src, dst = tmp_path / "in.gguf", tmp_path / "out.gguf"
    write_gguf(src, KVS, llama_like(), alignment=64)
    convert(src, dst, "cast:linear:q8_0")
    assert torch.cuda.current_device() == cur
    n = (1 << 23) * 3 + 32 * 5
    x = to_fdt(gaussian(n, 77), F16)
    assert lib().ggq_set_shard_devices(0) == ndev          # accepted: the thread was not pinned by convert
    try:
        got = ggq.quantize(Q8_0, x, F16)
    finally:
        lib().ggq_set_shard_devices(1)
    assert np.array_equal(got, oracle.quantize(Q8_0, F16, x, threads=8))
    assert torch.cuda.current_device() == cur
    y = torch.zeros(4, device="cuda")                      # the caller's own CUDA work still targets its device
    assert y.device.index == cur


@pytest.mark.gpu
def test_convert_direct_io_gives_the_same_bytes(ggq, oracle, tmp_path):
    """ggq_convert_options.direct_io: tensors that stream through a cast are read with O_DIRECT, 4 KiB-aligned, straight
    into the pinned staging buffers (tensor offsets are only 64-byte aligned; one tensor spans several pipeline chunks
    and ends at the end of the file).  The output must equal the buffered run byte for byte and the oracle per tensor."""
    from gguf_b200.convert import convert
    src = tmp_path / "in.gguf"
    ts = llama_like(big=(4096, 2400))            # 9.8 M elements: two pipeline chunks, not a multiple of 4 KiB
    write_gguf(src, KVS, ts, alignment=64)       # KVS declares general.alignment = 64
    a, b = tmp_path / "buffered.gguf", tmp_path / "direct.gguf"
    st0 = convert(src, a, "cast:linear:q8_0 embd:q4k")
    st1 = convert(src, b, "cast:linear:q8_0 embd:q4k", direct_io=True)
    assert st0["n_direct_inputs"] == 0
    print("inputs opened with O_DIRECT:", st1["n_direct_inputs"])     # 0 where the file system refuses it (fallback)
    assert open(a, "rb").read() == open(b, "rb").read()
    _, tensors, _, _ = read_gguf(b)
    name, shape, ty, data = ts[-1]
    want = oracle.quantize(Q8_0, F16, np.frombuffer(data, np.uint16), threads=8)
    assert tensors[name][1] == Q8_0 and np.array_equal(np.frombuffer(tensors[name][2], np.uint8), want)
