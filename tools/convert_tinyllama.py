"""BASELINE.json configs[0] at full size: TinyLlama-1.1B-shaped synthetic F16 GGUF (the reference's
fixture test-files/TinyLlama-1.1B-Chat-v1.0-F16.gguf holds 135 tensor infos and NO data),
`cast:linear:q8_0 embd:q8_0 -> cast:..:f32 -> cast:..:f16` (the only legal spelling of "Q8_0 and back to
F16" in the reference).  Converts on the GPU(s), verifies EVERY tensor byte against the CPU
oracle, and reports seconds per file for both.  Writes gpurun_out/convert_tinyllama.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from gguf_util import STRING, U32, read_gguf, write_gguf
from gguf_b200.convert import convert
from oracle import oracle as O

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0   # <1 shrinks the vocabulary / layer count for quick runs
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 0
layers = max(1, int(22 * scale)); vocab = max(256, int(32000 * scale) // 32 * 32)
tmp = os.environ.get("TMPDIR", "/tmp")
src, dst = os.path.join(tmp, "tinyllama_f16.gguf"), os.path.join(tmp, "tinyllama_rt.gguf")

t0 = time.time()
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
kvs = [("general.architecture", STRING, "llama"), ("general.name", STRING, "TinyLlama-1.1B-Chat-v1.0 (synthetic data)"),
       ("llama.block_count", U32, layers), ("llama.embedding_length", U32, 2048)]
write_gguf(src, kvs, tensors)
n_f16 = sum(int(np.prod(s)) for _, s, ty, _ in tensors if ty == 1)
print(f"generated {len(tensors)} tensors, {n_f16/1e9:.3f} G f16 elements, {os.path.getsize(src)/1e9:.2f} GB in {time.time()-t0:.1f}s", flush=True)

steps = "cast:linear:q8_0 embd:q8_0 -> cast:linear:f32 embd:f32 -> cast:linear:f16 embd:f16"
convert(src, dst, "cast:linear:q8_0", gpus)        # warm-up: contexts, pipelines, page cache
os.unlink(dst)
runs = []
for _ in range(3):
    if os.path.exists(dst): os.unlink(dst)   # the reference never overwrites (find_path, write.rs:104-126)
    t = time.time(); st = convert(src, dst, steps, gpus); runs.append(time.time() - t)
print("gpu convert s/file:", ["%.3f" % r for r in runs], st, flush=True)

# single-step Q8_0 conversion too (the common production step)
t = time.time(); st8 = convert(src, os.path.join(tmp, "tinyllama_q8.gguf"), "cast:linear:q8_0 embd:q8_0", gpus); t_q8 = time.time() - t

# ---- oracle: same chain on the host cores, every byte compared ----
threads = os.cpu_count() or 1
_, got, _, _ = read_gguf(dst)
_, got8, _, _ = read_gguf(os.path.join(tmp, "tinyllama_q8.gguf"))
t_cpu, bad = 0.0, 0
for name, shape, ty, data in tensors:
    if ty != 1:
        bad += got[name][2] != data
        continue
    x = np.frombuffer(data, np.uint16)
    t = time.time()
    q = O.quantize(8, O.F16, x, threads=threads)            # F16 -> Q8_0   (quantize::<Q8_0, f16, 32>)
    y32 = O.dequantize(8, O.F32, q, threads=threads)        # Q8_0 -> F32   (dequantize::<Q8_0, f32, 32>)
    y16 = O.quantize(1, O.F32, y32, threads=threads)        # F32 -> F16    (quantize::<f16, f32, 1>)
    t_cpu += time.time() - t
    bad += got[name][2] != y16.tobytes()
    bad += got8[name][2] != q.tobytes()
    assert got[name][1] == 1 and got8[name][1] == 8
res = {"config": "TinyLlama-1.1B-shaped synthetic F16 -> Q8_0 -> F32 -> F16 (BASELINE configs[0])", "scale": scale, "tensors": len(tensors),
       "f16_elements": n_f16, "file_GB": os.path.getsize(src) / 1e9, "gpu_seconds_per_file_roundtrip": min(runs), "gpu_runs": runs,
       "gpu_seconds_per_file_q8_0_only": t_q8, "n_devices": st["n_devices"], "stats_last": st,
       "cpu_oracle_seconds_compute_only": t_cpu, "cpu_threads": threads, "mismatching_tensors": int(bad),
       "note": "GPU time is wall clock of ggq_convert_gguf incl. mmap page faults, bounce copies, H2D/D2H and msync; CPU time is the oracle's compute only (no file IO)"}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "convert_tinyllama.json"), "w"), indent=1)
print(json.dumps(res))
assert bad == 0, "parity failure"
