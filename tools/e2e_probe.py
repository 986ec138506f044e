"""Host-pointer C ABI throughput (H2D + kernel + D2H inside the timing): pinned vs pageable caller
memory, one GPU vs block-range sharding over all visible GPUs."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gguf_b200 as g
from gguf_b200._lib import lib

n = 4096 * 14336 * 2          # 117 M elements: 14 pipeline chunks
x = (np.random.default_rng(0).standard_normal(n, dtype=np.float32) * 0.02).astype(np.float16)
rows = []
ndev_all = lib().ggq_device_count()
for ndev in sorted({1, ndev_all}):
    lib().ggq_set_shard_devices(ndev)
    for ty in (g.Q8_0, g.Q4_0):
        e, b = g.block_info(ty)
        nbytes = n // e * b
        for mem in os.environ.get("E2E_MEM", "pinned,pageable").split(","):
            if mem == "pinned":
                pf, pq = g.PinnedBuffer(n * 2), g.PinnedBuffer(nbytes)
                fa, qa = pf.view(np.uint16), pq.array
            else:
                fa, qa = np.empty(n, np.uint16), np.empty(nbytes, np.uint8)
            fa[:] = x.view(np.uint16)
            for direction in ("quant", "dequant"):
                fn = (lambda: g.quantize_slice(ty, qa, fa, g.F16)) if direction == "quant" else (lambda: g.dequantize_slice(ty, fa, qa, g.F16))
                fn()
                t0 = time.perf_counter()
                reps = 3
                for _ in range(reps):
                    fn()
                dt = (time.perf_counter() - t0) / reps
                gbs = (n * 2 + nbytes) / dt / 1e9
                rows.append({"gpus": ndev, "type": g.TYPE_NAMES[ty], "dir": direction, "memory": mem, "ms": dt * 1e3, "GBps": gbs})
                print(f"gpus={ndev} {g.TYPE_NAMES[ty]:5s} {direction:8s} {mem:9s} {dt*1e3:8.1f} ms {gbs:7.1f} GB/s", flush=True)
            if mem == "pinned":
                pf.free(); pq.free()
lib().ggq_set_shard_devices(1)
json.dump(rows, open("gpurun_out/e2e_probe.json", "w"), indent=1)
