"""Device-resident GB/s of every codec direction and type (FFN-shaped tensor, rotating buffers > L2).
Writes a table to stdout and JSON to gpurun_out/codec_sweep.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gguf_b200 as g

PEAK = 6543.4
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 14336
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
st = torch.cuda.current_stream().cuda_stream
TD = {g.F32: torch.float32, g.F16: torch.float16, g.BF16: torch.bfloat16}
x32 = torch.randn(n, device="cuda") * 0.02
rows = []


def timeit(fn, nsets, reps):
    for i in range(nsets):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i % nsets)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


for ty in [g.Q4_0, g.Q4_1, g.Q5_0, g.Q5_1, g.Q8_0, g.Q8_1, g.Q8K, g.Q2K, g.Q3K, g.Q4K, g.Q5K, g.Q6K, g.F16, g.BF16]:
    e, b = g.block_info(ty)
    name = g.TYPE_NAMES[ty]
    kq = ty in (g.Q2K, g.Q3K, g.Q4K, g.Q5K, g.Q6K)
    for fdt in (g.F16, g.F32, g.BF16):
        if ty in (g.F16, g.BF16) and fdt == ty:
            continue
        nsets = 4
        xs = [x32.to(TD[fdt]).clone() for _ in range(nsets)]
        packed = [torch.empty(n // e * b, dtype=torch.uint8, device="cuda") for _ in range(nsets)]
        outs = [torch.empty(n, dtype=TD[fdt], device="cuda") for _ in range(nsets)]
        for direction in ("quant", "dequant"):
            if only and f"{name}:{direction}" not in only and name not in only and direction not in only:
                continue
            if direction == "quant":
                fn = lambda i: g.quantize_slice_device(ty, fdt, packed[i], n // e, xs[i], n, st)
                reps = 6 if kq else 40
            else:
                g.quantize_slice_device(ty, fdt, packed[0], n // e, xs[0], n, st)
                for i in range(1, nsets):
                    packed[i].copy_(packed[0])
                fn = lambda i: g.dequantize_slice_device(ty, fdt, outs[i], n, packed[i], n // e, st)
                reps = 40
            t = timeit(fn, nsets, reps)
            nbytes = n // e * b + n * (4 if fdt == g.F32 else 2)
            gbs = nbytes / t / 1e9
            rows.append({"type": name, "fdt": g.TYPE_NAMES[fdt], "dir": direction, "us": t * 1e6, "GBps": gbs, "frac_measured_peak": gbs / PEAK})
            print(f"{name:5s} {direction:8s} {g.TYPE_NAMES[fdt]:4s} {t*1e6:10.1f} us {gbs:8.1f} GB/s  {100*gbs/PEAK:5.1f}% of measured peak", flush=True)
        del xs, packed, outs
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"n_elems": n, "peak_GBps": PEAK, "rows": rows}, open("gpurun_out/codec_sweep.json", "w"), indent=1)
