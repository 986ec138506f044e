"""BASELINE.json configs[4], scaled to what one box's disk holds: Llama-3-8B-shaped synthetic F16 GGUF
(16 GB; the 70B file would be 141 GB) -> `cast:linear:q5k embd:q6k` whole-file convert with pinned async
H2D/D2H, then a strided sample of output super-blocks is re-quantised by the CPU oracle and compared.
Writes gpurun_out/convert_llama8b.json."""
import json, os, struct, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from gguf_util import STRING, U32, index_gguf, kv_bytes
from gguf_b200.convert import convert
from oracle import oracle as O

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 0
model = sys.argv[3] if len(sys.argv) > 3 else "8b"     # "70b": Llama-3-70B tensor shapes (hidden 8192, ffn 28672, 80 layers in full)
H, FFN, KV, VOCAB = (8192, 28672, 1024, 128256) if model == "70b" else (4096, 14336, 1024, 128256)
tmp = os.environ.get("TMPDIR", "/tmp")
src, dst = os.path.join(tmp, f"llama{model}_f16.gguf"), os.path.join(tmp, f"llama{model}_q5k.gguf")
shapes = [("token_embd.weight", (H, VOCAB))]
for l in range(layers):
    shapes += [(f"blk.{l}.attn_norm.weight", (H,)), (f"blk.{l}.attn_q.weight", (H, H)), (f"blk.{l}.attn_k.weight", (H, KV)),
               (f"blk.{l}.attn_v.weight", (H, KV)), (f"blk.{l}.attn_output.weight", (H, H)), (f"blk.{l}.ffn_norm.weight", (H,)),
               (f"blk.{l}.ffn_gate.weight", (H, FFN)), (f"blk.{l}.ffn_up.weight", (H, FFN)), (f"blk.{l}.ffn_down.weight", (FFN, H))]
shapes += [("output_norm.weight", (H,)), ("output.weight", (H, VOCAB))]
# ---- write the input file streaming (header + infos first, then GPU-generated tensor data) ----
t0 = time.time()
kvs = [("general.architecture", STRING, "llama"), ("general.name", STRING, f"Llama-3-{model.upper()}-shaped synthetic"), ("llama.block_count", U32, layers)]
def s_(x): b = x.encode(); return struct.pack("<Q", len(b)) + b
infos, off, offs = b"", 0, []
for name, shape in shapes:
    ty = 0 if len(shape) == 1 else 1
    n = int(np.prod(shape)); nbytes = n * (4 if ty == 0 else 2)
    off += (32 - off % 32) % 32; offs.append(off)
    infos += s_(name) + struct.pack("<I", len(shape)) + b"".join(struct.pack("<Q", d) for d in shape) + struct.pack("<IQ", ty, off)
    off += nbytes
head = b"GGUF" + struct.pack("<IQQ", 3, len(shapes), len(kvs)) + b"".join(kv_bytes(*kv) for kv in kvs) + infos
head += b"\0" * ((32 - len(head) % 32) % 32)
gen = torch.Generator(device="cuda"); gen.manual_seed(4)
with open(src, "wb") as f:
    f.write(head)
    base = len(head)
    for (name, shape), o in zip(shapes, offs):
        n = int(np.prod(shape))
        x = torch.randn(n, device="cuda", generator=gen) * 0.02
        f.seek(base + o)
        (x if len(shape) == 1 else x.to(torch.float16)).cpu().numpy().tofile(f)
n_lin = sum(int(np.prod(s)) for _, s in shapes if len(s) > 1)
print(f"generated {len(shapes)} tensors, {n_lin/1e9:.2f} G linear elements, {os.path.getsize(src)/1e9:.2f} GB in {time.time()-t0:.1f}s", flush=True)

steps = "cast:linear:q5k embd:q6k"
# warm-up (CUDA contexts, stream pipelines, pinned buffers) on a one-tensor file
from gguf_util import write_gguf
write_gguf(os.path.join(tmp, "warm_in.gguf"), kvs, [("blk.0.attn_q.weight", (4096, 4096), 1, np.zeros(4096 * 4096, np.float16).tobytes())])
convert(os.path.join(tmp, "warm_in.gguf"), os.path.join(tmp, "warm.gguf"), "cast:linear:q5k", gpus)
os.unlink(os.path.join(tmp, "warm.gguf")); os.unlink(os.path.join(tmp, "warm_in.gguf"))
runs = []
for _ in range(2):
    if os.path.exists(dst): os.unlink(dst)
    t = time.time(); st = convert(src, dst, steps, gpus); runs.append(time.time() - t)
print("gpu convert s/file:", ["%.2f" % r for r in runs], st, flush=True)
# sharded output (-s 1G): each shard is its own inode, so the buffered writes no longer serialise
import glob
t = time.time(); st_sh = convert(src, os.path.join(tmp, f"llama{model}_sh.gguf"), steps, gpus, max_bytes="1G"); t_sharded = time.time() - t
for f_ in glob.glob(os.path.join(tmp, f"llama{model}_sh-*.gguf")): os.unlink(f_)
print("sharded (-s 1G):", "%.2f s" % t_sharded, st_sh["n_out_files"], "files", flush=True)

# ---- verify a strided sample of super-blocks of every cast tensor against the oracle ----
threads = os.cpu_count() or 1
out_t, _, out_mm = index_gguf(dst)
out_size = os.path.getsize(dst)
inp = np.memmap(src, dtype=np.uint8, mode="r")
bad = sampled = 0
t_cpu = 0.0
for (name, shape), o in zip(shapes, offs):
    if len(shape) == 1:
        continue
    ty = 14 if name in ("token_embd.weight", "output.weight") else 13
    _, b = O.block_info(ty)
    n = int(np.prod(shape)); nsb = n // 256
    idx = np.arange(0, nsb, 4001)
    x = np.frombuffer(inp, dtype=np.uint16, count=n, offset=base + o).reshape(-1, 256)[idx].reshape(-1)
    t = time.time(); want = O.quantize(ty, O.F16, np.ascontiguousarray(x), threads=threads).reshape(-1, b); t_cpu += time.time() - t
    got = out_mm[out_t[name][2]:out_t[name][2] + out_t[name][3]].reshape(-1, b)[idx]
    assert out_t[name][1] == ty
    bad += int((got != want).any(axis=1).sum()); sampled += len(idx)
# CPU baseline rate: the oracle's K-quant quantizers on a buffer big enough for every host thread
xb = (np.random.default_rng(1).standard_normal(256 * 65536, dtype=np.float32) * 0.02).astype(np.float16).view(np.uint16)
t = time.time(); O.quantize(13, O.F16, xb, threads=threads); t5 = time.time() - t
t = time.time(); O.quantize(14, O.F16, xb, threads=threads); t6 = time.time() - t
n_q6 = 2 * H * VOCAB
cpu_seconds = (n_lin - n_q6) / (xb.size / t5) + n_q6 / (xb.size / t6)
cpu_rate = n_lin / cpu_seconds
res = {"config": f"Llama-3-{model.upper()}-shaped synthetic F16 ({layers} layers) -> Q5_K (linear) / Q6_K (embd) whole-file convert (BASELINE configs[4]"
                 + (" scaled from 70B / 8 GPUs)" if model == "8b" else f": 70B tensor shapes, {layers} of 80 layers = what one box's RAM disk holds)"),
       "layers": layers, "tensors": len(shapes), "linear_elements": n_lin, "file_in_GB": os.path.getsize(src) / 1e9, "file_out_GB": out_size / 1e9,
       "gpu_seconds_per_file": min(runs), "gpu_runs": runs, "n_devices": st["n_devices"], "stats_last": st,
       "gpu_seconds_per_file_sharded_1G": t_sharded, "sharded_files": st_sh["n_out_files"],
       "sampled_super_blocks": sampled, "mismatching_super_blocks": bad,
       "cpu_oracle_elements_per_second": cpu_rate, "cpu_threads": threads,
       "cpu_oracle_q5k_Melem_per_s": xb.size / t5 / 1e6, "cpu_oracle_q6k_Melem_per_s": xb.size / t6 / 1e6,
       "cpu_oracle_seconds_per_file_extrapolated": cpu_seconds,
       "note": "GPU time is the wall clock of ggq_convert_gguf (file read, H2D, K-quant kernels, D2H, file write); the CPU figure extrapolates the oracle's K-quant rate (16.8 M-element buffer, all host threads) to the whole file, compute only"}
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"convert_llama{model}.json"), "w"), indent=1)
print(json.dumps(res))
os.unlink(src); os.unlink(dst)
assert bad == 0
