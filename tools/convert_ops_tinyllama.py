"""The block-granular operators at BASELINE configs[0] size: TinyLlama-1.1B-shaped synthetic F16 GGUF (merged
attn_qkv / ffn_gate_up, 32 heads, 4 KV heads), through `split-linear`, `permute-qk`, `merge-linear` and
`split-linear -> permute-qk -> cast:linear:q8_0 embd:q8_0`.  Every output tensor is compared with the
oracle's operators (oracle/rearrange.py) and codecs; merge(split(file)) must reproduce the input file's
tensor bytes.  Writes gpurun_out/convert_ops_tinyllama.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from gguf_util import STRING, U32, read_gguf, write_gguf
from gguf_b200.convert import convert
from oracle import oracle as O
from oracle import rearrange as R

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 0
layers = max(1, int(22 * scale)); vocab = max(256, int(32000 * scale) // 32 * 32)
NH, NKVH = 32, 4
tmp = os.environ.get("TMPDIR", "/tmp")
src = os.path.join(tmp, "tinyllama_ops_f16.gguf")

tensors, idx = [], 0
def add(name, shape, f32=False):
    global idx
    n = int(np.prod(shape))
    x = np.random.default_rng(idx).standard_normal(n, dtype=np.float32) * np.float32(0.02)
    tensors.append((name, tuple(shape), 0 if f32 else 1, x.tobytes() if f32 else x.astype(np.float16).tobytes()))
    idx += 1
add("token_embd.weight", (2048, vocab))
for l in range(layers):
    add(f"blk.{l}.attn_norm.weight", (2048,), True)
    add(f"blk.{l}.attn_qkv.weight", (2048, 2560))
    add(f"blk.{l}.attn_output.weight", (2048, 2048))
    add(f"blk.{l}.ffn_norm.weight", (2048,), True)
    add(f"blk.{l}.ffn_gate_up.weight", (2048, 11264))
    add(f"blk.{l}.ffn_down.weight", (5632, 2048))
add("output_norm.weight", (2048,), True)
add("output.weight", (2048, vocab))
kvs = [("general.architecture", STRING, "llama"), ("llama.block_count", U32, layers), ("llama.embedding_length", U32, 2048),
       ("llama.attention.head_count", U32, NH), ("llama.attention.head_count_kv", U32, NKVH)]
write_gguf(src, kvs, tensors)
print(f"generated {len(tensors)} tensors, {os.path.getsize(src)/1e9:.2f} GB", flush=True)
convert(src, os.path.join(tmp, "warm.gguf"), "cast:linear:q8_0", gpus)   # contexts, pipelines, page cache

def timed(inp, out, steps):
    runs = []
    for _ in range(3):
        if os.path.exists(out): os.unlink(out)
        t = time.time(); st = convert(inp, out, steps, gpus); runs.append(time.time() - t)
    return min(runs), st

def listed(ts):
    return [(n, (ty, tuple(s), np.frombuffer(d, np.uint8))) for n, s, ty, d in ts]

def compare(path, want):
    _, got, _, _ = read_gguf(path)
    assert list(got) == [n for n, _ in want], "tensor order"
    bad = 0
    for n, (ty, shape, data) in want:
        bad += not (tuple(got[n][0]) == tuple(shape) and got[n][1] == ty and got[n][2] == data.tobytes())
    return bad

res = {"config": "TinyLlama-1.1B-shaped synthetic F16, 135 tensors (BASELINE configs[0] shapes)", "file_GB": os.path.getsize(src) / 1e9, "cases": {}}
cur = listed(tensors)
p = lambda name: os.path.join(tmp, name)

t, st = timed(src, p("split.gguf"), "split-linear")
t0 = time.time(); want_split = R.split_linear(cur, NH, NKVH); tc = time.time() - t0
res["cases"]["split-linear"] = {"gpu_s_per_file": t, "oracle_numpy_s": tc, "mismatching_tensors": compare(p("split.gguf"), want_split), "stats": st}
print("split-linear", res["cases"]["split-linear"], flush=True)

t, st = timed(p("split.gguf"), p("merged.gguf"), "merge-linear")
_, got, _, _ = read_gguf(p("merged.gguf"))
bad = sum(got[n][2] != d for n, s, ty, d in tensors)
res["cases"]["merge-linear(split-linear(file)) == file"] = {"gpu_s_per_file": t, "mismatching_tensors": int(bad), "stats": st}
print("merge-linear", res["cases"]["merge-linear(split-linear(file)) == file"], flush=True)

t, st = timed(src, p("perm.gguf"), "permute-qk")
t0 = time.time(); want_perm = R.permute_qk_all(cur, NH, NKVH); tc = time.time() - t0
res["cases"]["permute-qk"] = {"gpu_s_per_file": t, "oracle_numpy_s": tc, "mismatching_tensors": compare(p("perm.gguf"), want_perm), "stats": st}
print("permute-qk", res["cases"]["permute-qk"], flush=True)

steps = "split-linear -> permute-qk -> cast:linear:q8_0 embd:q8_0"
t, st = timed(src, p("full.gguf"), steps)
t0 = time.time()
want = R.permute_qk_all(want_split, NH, NKVH)
threads = os.cpu_count() or 1
full = []
for n, (ty, shape, data) in want:
    if ty == 1:
        full.append((n, (8, shape, O.quantize(8, O.F16, data.view(np.uint16), threads=threads).view(np.uint8))))
    else:
        full.append((n, (ty, shape, data)))
tc = time.time() - t0
res["cases"][steps] = {"gpu_s_per_file": t, "oracle_s_permute_plus_quantize": tc, "cpu_threads": threads, "mismatching_tensors": compare(p("full.gguf"), full), "stats": st}
print(steps, res["cases"][steps], flush=True)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "convert_ops_tinyllama.json"), "w"), indent=1)
assert all(c["mismatching_tensors"] == 0 for c in res["cases"].values()), "parity failure"
print("OK")
