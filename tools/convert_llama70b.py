"""BASELINE.json configs[4] as written: Llama-3-70B-shaped synthetic F16 GGUF (hidden 8192, ffn 28672, kv 1024,
vocab 128256, 80 layers = 141 GB) -> `cast:linear:q5k embd:q6k` whole-file convert through ggq_convert_gguf_ex
across every visible B200 (pread -> pinned -> H2D -> K-quant kernels -> D2H -> pinned -> pwrite per worker, tensors
spread over the GPUs, no inter-GPU traffic), end to end against the host-core CPU baseline.

The input lives on the box's RAM disk; the layer count drops below 80 only if the RAM disk cannot hold input + output
(the JSON says so).  Reports wall s/file, per-stage worker time and GB/s, the PCIe-bound fraction, a sampled byte
parity check against the oracle and the oracle's own K-quant rate on all host threads.
usage: python tools/convert_llama70b.py [layers=80] [gpus=0 (all)] ; writes gpurun_out/convert_llama70b_<N>gpu.json"""
import glob
import json
import os
import shutil
import struct
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from gguf_util import STRING, U32, index_gguf, kv_bytes, write_gguf
from gguf_b200._lib import lib
from gguf_b200.convert import convert
from oracle import oracle as O

want_layers = int(sys.argv[1]) if len(sys.argv) > 1 else 80
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ndev = lib().ggq_device_count() if gpus <= 0 else gpus
H, FFN, KV, VOCAB = 8192, 28672, 1024, 128256
tmp = os.environ.get("GGQ_TMP", "/dev/shm")
per_layer_in = (2 * H * H + 2 * H * KV + 3 * H * FFN) * 2
per_layer_out = per_layer_in // 2 * 176 // 256
fixed_in, fixed_out = 2 * H * VOCAB * 2, 2 * H * VOCAB * 210 // 256
free = shutil.disk_usage(tmp).free
layers = want_layers
while layers > 1 and fixed_in + fixed_out + layers * (per_layer_in + per_layer_out) > free - (24 << 30):
    layers -= 1
src, dst = os.path.join(tmp, "llama70b_f16.gguf"), os.path.join(tmp, "llama70b_q5k.gguf")
shapes = [("token_embd.weight", (H, VOCAB))]
for l in range(layers):
    shapes += [(f"blk.{l}.attn_norm.weight", (H,)), (f"blk.{l}.attn_q.weight", (H, H)), (f"blk.{l}.attn_k.weight", (H, KV)),
               (f"blk.{l}.attn_v.weight", (H, KV)), (f"blk.{l}.attn_output.weight", (H, H)), (f"blk.{l}.ffn_norm.weight", (H,)),
               (f"blk.{l}.ffn_gate.weight", (H, FFN)), (f"blk.{l}.ffn_up.weight", (H, FFN)), (f"blk.{l}.ffn_down.weight", (FFN, H))]
shapes += [("output_norm.weight", (H,)), ("output.weight", (H, VOCAB))]

# ---- the input file: header + infos, then tensor data cut from a 4.3 GB pool of GPU-generated N(0, 0.02^2) f16 values at
# a different offset per tensor (distinct bytes per tensor without 141 GB of random-number generation), written by 16 threads
t0 = time.time()
kvs = [("general.architecture", STRING, "llama"), ("general.name", STRING, "Llama-3-70B-shaped synthetic"), ("llama.block_count", U32, layers)]


def s_(x):
    b = x.encode()
    return struct.pack("<Q", len(b)) + b


infos, off, offs = b"", 0, []
for name, shape in shapes:
    ty = 0 if len(shape) == 1 else 1
    n = int(np.prod(shape))
    off += (32 - off % 32) % 32
    offs.append(off)
    infos += s_(name) + struct.pack("<I", len(shape)) + b"".join(struct.pack("<Q", d) for d in shape) + struct.pack("<IQ", ty, off)
    off += n * (4 if ty == 0 else 2)
head = b"GGUF" + struct.pack("<IQQ", 3, len(shapes), len(kvs)) + b"".join(kv_bytes(*kv) for kv in kvs) + infos
head += b"\0" * ((32 - len(head) % 32) % 32)
base = len(head)
gen = torch.Generator(device="cuda"); gen.manual_seed(4)
POOL = H * VOCAB * 2 + (1 << 24)
pool = np.empty(POOL, np.float16)
for a in range(0, POOL, 1 << 28):
    nn = min(1 << 28, POOL - a)
    pool[a:a + nn] = (torch.randn(nn, device="cuda", generator=gen) * 0.02).to(torch.float16).cpu().numpy()
pool_u8 = pool.view(np.uint8)
fd = os.open(src, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
os.ftruncate(fd, base + off)
os.pwrite(fd, head, 0)
pool_off = {}


def put(i):
    (name, shape), o = shapes[i], offs[i]
    n = int(np.prod(shape))
    if len(shape) == 1:
        os.pwrite(fd, np.ones(n, np.float32).tobytes(), base + o)
        return
    start = (i * 1000003 * 256) % (POOL - n) // 256 * 256           # whole super-blocks, differs per tensor
    pool_off[name] = start
    mv = memoryview(pool_u8[2 * start:2 * (start + n)])
    for a in range(0, len(mv), 1 << 26):
        os.pwrite(fd, mv[a:a + (1 << 26)], base + o + a)


with ThreadPoolExecutor(16) as ex:
    list(ex.map(put, range(len(shapes))))
os.close(fd)
n_lin = sum(int(np.prod(s)) for _, s in shapes if len(s) > 1)
t_gen = time.time() - t0
print(f"generated {len(shapes)} tensors, {layers} layers, {n_lin/1e9:.2f} G linear elements, {os.path.getsize(src)/1e9:.2f} GB in {t_gen:.1f}s", flush=True)

steps = "cast:linear:q5k embd:q6k"
# warm-up on a small file: CUDA contexts, stream pipelines and pinned buffers of every worker on every device
write_gguf(os.path.join(tmp, "warm_in.gguf"), kvs, [(f"blk.{i}.attn_q.weight", (4096, 4096), 1, np.zeros(4096 * 4096, np.float16).tobytes()) for i in range(8 * ndev)])
convert(os.path.join(tmp, "warm_in.gguf"), os.path.join(tmp, "warm.gguf"), "cast:linear:q5k", gpus)
os.unlink(os.path.join(tmp, "warm.gguf")); os.unlink(os.path.join(tmp, "warm_in.gguf"))


def stage_report(st, wall):
    w = max(st["n_workers"], 1)
    return {"wall_seconds": wall, "n_devices": st["n_devices"], "n_workers": st["n_workers"], "out_files": st["n_out_files"],
            "file_read_GBps": st["bytes_in"] / wall / 1e9, "file_write_GBps": st["bytes_out"] / wall / 1e9,
            "h2d_GB": st["h2d_bytes"] / 1e9, "d2h_GB": st["d2h_bytes"] / 1e9,
            "worker_seconds": {"read": st["worker_seconds_read"], "write": st["worker_seconds_write"], "gpu_wait": st["worker_seconds_gpu_wait"]},
            "avg_share_of_a_worker's_wall_time": {"read": st["worker_seconds_read"] / w / wall, "write": st["worker_seconds_write"] / w / wall,
                                                  "gpu_wait": st["worker_seconds_gpu_wait"] / w / wall},
            # lower bound of the PCIe stage: the busier direction over the devices' links at the measured 55 GB/s per direction
            "pcie_seconds_lower_bound": max(st["h2d_bytes"], st["d2h_bytes"]) / st["n_devices"] / 55e9,
            "pcie_bound_fraction": max(st["h2d_bytes"], st["d2h_bytes"]) / st["n_devices"] / 55e9 / wall}


runs = []
for _ in range(int(os.environ.get("GGQ_70B_SINGLE_RUNS", "2"))):
    if os.path.exists(dst):
        os.unlink(dst)
    t = time.time(); st = convert(src, dst, steps, gpus); wall = time.time() - t
    runs.append(stage_report(st, wall))
    print("convert, one output file:", json.dumps(runs[-1]), flush=True)
sharded_runs = {}
for size in os.environ.get("GGQ_70B_SHARDS", "4G").split(","):
    t = time.time(); st_sh = convert(src, os.path.join(tmp, "llama70b_sh.gguf"), steps, gpus, max_bytes=size); wall = time.time() - t
    sharded_runs[size] = stage_report(st_sh, wall)
    print(f"convert, -s {size}:", json.dumps(sharded_runs[size]), flush=True)
    if not runs and size == os.environ.get("GGQ_70B_SHARDS", "4G").split(",")[-1]:
        # keep the last sharded output for the parity check when no single-file run was made
        keep = sorted(glob.glob(os.path.join(tmp, "llama70b_sh-*.gguf")))
    else:
        for f_ in glob.glob(os.path.join(tmp, "llama70b_sh-*.gguf")):
            os.unlink(f_)
sharded = sharded_runs[min(sharded_runs, key=lambda k: sharded_runs[k]["wall_seconds"])]
one_gpu = None
if ndev > 1 and os.environ.get("GGQ_70B_ALSO_1GPU", "1") == "1":
    t = time.time(); st1 = convert(src, os.path.join(tmp, "llama70b_1.gguf"), steps, 1, max_bytes="4G"); wall = time.time() - t
    one_gpu = stage_report(st1, wall)
    print("convert on ONE gpu, -s 4G:", json.dumps(one_gpu), flush=True)
    for f_ in glob.glob(os.path.join(tmp, "llama70b_1-*.gguf")):
        os.unlink(f_)

# ---- parity: a strided sample of super-blocks of every cast tensor re-quantised by the oracle ----
threads = os.cpu_count() or 1
if runs:
    shard_index = [index_gguf(dst)]
    out_size = os.path.getsize(dst)
else:
    shard_index = [index_gguf(f_) for f_ in keep]
    out_size = sum(os.path.getsize(f_) for f_ in keep)
where = {}
for ti, (tt, _, mm) in enumerate(shard_index):
    for name_ in tt:
        where[name_] = ti
bad = sampled = 0
for name, shape in shapes:
    if len(shape) == 1:
        continue
    ty = 14 if name in ("token_embd.weight", "output.weight") else 13
    _, b = O.block_info(ty)
    n = int(np.prod(shape)); nsb = n // 256
    idx = np.arange(0, nsb, 8009)
    x = pool[pool_off[name]:pool_off[name] + n].view(np.uint16).reshape(-1, 256)[idx].reshape(-1)
    want = O.quantize(ty, O.F16, np.ascontiguousarray(x), threads=threads).reshape(-1, b)
    out_t, _, out_mm = shard_index[where[name]]
    assert out_t[name][1] == ty
    got = out_mm[out_t[name][2]:out_t[name][2] + out_t[name][3]].reshape(-1, b)[idx]
    bad += int((got != want).any(axis=1).sum()); sampled += len(idx)
# ---- CPU baseline: the oracle's K-quant quantizers on every host thread, extrapolated to the file (compute only) ----
xb = pool[:256 * 65536 * 2].view(np.uint16)
t = time.time(); O.quantize(13, O.F16, xb, threads=threads); t5 = time.time() - t
t = time.time(); O.quantize(14, O.F16, xb, threads=threads); t6 = time.time() - t
n_q6 = 2 * H * VOCAB
cpu_seconds = (n_lin - n_q6) / (xb.size / t5) + n_q6 / (xb.size / t6)
best = min(runs + list(sharded_runs.values()), key=lambda r: r["wall_seconds"])
res = {"config": f"Llama-3-70B-shaped synthetic F16 ({layers} of 80 layers) -> Q5_K (linear) / Q6_K (embd) whole-file convert across {ndev} B200 (BASELINE configs[4])",
       "layers": layers, "layers_requested": want_layers, "ram_disk_free_GB_at_start": free / 1e9, "tensors": len(shapes), "linear_elements": n_lin,
       "file_in_GB": os.path.getsize(src) / 1e9, "file_out_GB": out_size / 1e9, "generate_seconds": t_gen,
       "seconds_per_file": best["wall_seconds"], "runs_one_output_file": runs, "runs_sharded": sharded_runs, "run_one_gpu_sharded_4G": one_gpu,
       "speedup_N_gpus_vs_1_sharded": (one_gpu["wall_seconds"] / sharded["wall_seconds"]) if one_gpu else None,
       "sampled_super_blocks": sampled, "mismatching_super_blocks": bad,
       "cpu_baseline": {"kind": "port", "threads": threads, "q5k_Melem_per_s": xb.size / t5 / 1e6, "q6k_Melem_per_s": xb.size / t6 / 1e6,
                        "seconds_per_file_extrapolated_compute_only": cpu_seconds,
                        "speedup_end_to_end_vs_cpu_compute": cpu_seconds / min(best["wall_seconds"], sharded["wall_seconds"])},
       "host": {"cpus": threads, "ram_disk": tmp},
       "note": "wall_seconds is the wall clock of ggq_convert_gguf_ex (file read, H2D, K-quant kernels, D2H, file write); worker_seconds are summed over the "
               "worker threads; the CPU figure extrapolates the oracle's K-quant rate on all host threads to the whole file, compute only (no file I/O)"}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"convert_llama70b_{ndev}gpu.json"), "w"), indent=1)
print(json.dumps(res))
os.unlink(src)
for f_ in ([dst] if runs else keep):
    os.unlink(f_)
assert bad == 0
