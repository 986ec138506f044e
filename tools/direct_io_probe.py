"""O_DIRECT vs buffered input for a file that is not in the page cache: an 8B-shaped slice (default 6 layers, ~3 GB F16) on
the box's DISK (not the RAM disk), page cache dropped before each run, output to the RAM disk so only the read side
differs.  Writes gpurun_out/direct_io_probe.json.  usage: python tools/direct_io_probe.py [layers] [disk_dir]"""
import json, os, struct, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from gguf_util import STRING, U32, kv_bytes
from gguf_b200.convert import convert

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 6
disk = sys.argv[2] if len(sys.argv) > 2 else "/var/tmp"
H, FFN, KV = 4096, 14336, 1024
src, out = os.path.join(disk, "dio_in.gguf"), "/dev/shm/dio_out.gguf"
shapes = []
for l in range(layers):
    shapes += [(f"blk.{l}.attn_q.weight", (H, H)), (f"blk.{l}.attn_k.weight", (H, KV)), (f"blk.{l}.ffn_gate.weight", (H, FFN)),
               (f"blk.{l}.ffn_up.weight", (H, FFN)), (f"blk.{l}.ffn_down.weight", (FFN, H))]
kvs = [("general.architecture", STRING, "llama"), ("llama.block_count", U32, layers)]
def s_(x): b = x.encode(); return struct.pack("<Q", len(b)) + b
infos, off, offs = b"", 0, []
for name, shape in shapes:
    n = int(np.prod(shape)); off += (32 - off % 32) % 32; offs.append(off)
    infos += s_(name) + struct.pack("<I", len(shape)) + b"".join(struct.pack("<Q", d) for d in shape) + struct.pack("<IQ", 1, off)
    off += n * 2
head = b"GGUF" + struct.pack("<IQQ", 3, len(shapes), len(kvs)) + b"".join(kv_bytes(*kv) for kv in kvs) + infos
head += b"\0" * ((32 - len(head) % 32) % 32) + b"\0" * 32   # data region deliberately NOT 4 KiB aligned
pool = (np.random.default_rng(0).standard_normal(H * FFN + 4096, dtype=np.float32) * np.float32(0.02)).astype(np.float16)
t0 = time.time()
with open(src, "wb") as f:
    f.write(head[:-32]); base = f.tell()
    for (name, shape), o in zip(shapes, offs):
        n = int(np.prod(shape)); f.seek(base + o); pool[(o // 2) % 4096:(o // 2) % 4096 + n].tofile(f)
    f.flush(); os.fsync(f.fileno())
size = os.path.getsize(src)
print(f"wrote {size/1e9:.2f} GB to {src} in {time.time()-t0:.1f}s", flush=True)
def drop():
    subprocess.run("sync; echo 3 > /proc/sys/vm/drop_caches", shell=True, check=False)
rows = []
convert(src, out, "cast:linear:q8_0")   # warm-up of contexts / pipelines (leaves the file cached)
for mode in ("buffered_cold", "direct_cold", "buffered_cold", "direct_cold", "buffered_warm"):
    if mode.endswith("cold"): drop()
    t = time.time(); st = convert(src, out, "cast:linear:q8_0", direct_io=mode.startswith("direct")); dt = time.time() - t
    rows.append({"mode": mode, "seconds": dt, "read_GBps": size / dt / 1e9, "n_direct_inputs": st["n_direct_inputs"],
                 "worker_seconds_read": st["worker_seconds_read"], "n_workers": st["n_workers"]})
    print(rows[-1], flush=True)
    os.unlink(out)
os.unlink(src)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"file_GB": size / 1e9, "disk_dir": disk, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "direct_io_probe.json"), "w"), indent=1)
