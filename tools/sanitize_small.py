"""Small run of every kernel (all types, f32/f16/bf16, ragged sizes, misaligned device pointers) — the
target of compute-sanitizer memcheck / racecheck.  Also checks results against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import gguf_b200 as g
from oracle import oracle as O
from data import gaussian, to_fdt, same_blocks, same_floats
st = torch.cuda.current_stream().cuda_stream
bad = 0
for ty in [g.Q4_0, g.Q4_1, g.Q5_0, g.Q5_1, g.Q8_0, g.Q8_1, g.Q8K, g.Q2K, g.Q3K, g.Q4K, g.Q5K, g.Q6K]:
    e, b = g.block_info(ty)
    nb = (300 if e == 32 else 70) + 3       # ragged: > 1 tile for the row kernel, partial last tile
    for fdt in (g.F32, g.F16, g.BF16):
        x = to_fdt(gaussian(e * nb, ty * 7 + fdt), fdt)
        want_q = O.quantize(ty, fdt, x)
        want_d = O.dequantize(ty, fdt, want_q)
        for off in (0, 4 if fdt == g.F32 else 2):   # 16-byte aligned, and only element-aligned device pointers
            src = torch.zeros(x.nbytes + 32, dtype=torch.uint8, device="cuda")
            src[off:off + x.nbytes] = torch.from_numpy(x.view(np.uint8)).cuda()
            q = torch.zeros(nb * b + 32, dtype=torch.uint8, device="cuda")
            g.quantize_slice_device(ty, fdt, q.data_ptr() + off, nb, src.data_ptr() + off, e * nb, st)
            d = torch.zeros(x.nbytes + 32, dtype=torch.uint8, device="cuda")
            g.dequantize_slice_device(ty, fdt, d.data_ptr() + off, e * nb, q.data_ptr() + off, nb, st)
            torch.cuda.synchronize()
            gq = q.cpu().numpy()[off:off + nb * b]
            gd = d.cpu().numpy()[off:off + x.nbytes].view(want_d.dtype)
            if not same_blocks(gq, want_q, ty, b) or not same_floats(gd, want_d):
                bad += 1; print("MISMATCH", g.TYPE_NAMES[ty], fdt, off)
for s_, d_ in [(g.F32, g.F16), (g.F16, g.F32), (g.BF16, g.F16), (g.F32, g.BF16)]:   # casts, ragged + tail
    n = 8192 * 3 + 37
    x = to_fdt(gaussian(n, 5), s_)
    out = g.quantize(d_, x, s_) if d_ != g.F32 else g.dequantize(s_, x.view(np.uint8), g.F32)
    want = O.quantize(d_, s_, x) if d_ != g.F32 else O.dequantize(s_, g.F32, x.view(np.uint8))
    bad += not np.array_equal(out.view(np.uint8), want.view(np.uint8))
# the batched descriptor-table grid with wild scale fields (exact pass for NaN / inf scales) and a ragged last tile per job
from data import random_packed
jobs, want, keep = [], [], []
for ty in [g.Q4_0, g.Q4_1, g.Q5_0, g.Q5_1, g.Q8_0, g.Q8_1, g.Q8K, g.Q2K, g.Q3K, g.Q4K, g.Q5K, g.Q6K]:
    e, b = g.block_info(ty)
    nb = 16384 // e + 5
    blocks = random_packed(ty, nb, b, 40 + ty, wild=True)
    dq, dy = torch.from_numpy(blocks).cuda(), torch.zeros(nb * e, dtype=torch.float16, device="cuda")
    jobs.append(("dequantize", ty, g.F16, dy, nb * e, dq, nb)); keep.append((dq, dy)); want.append(O.dequantize(ty, O.F16, blocks))
g.slices_device(jobs, st)
torch.cuda.synchronize()
for (_, dy), w in zip(keep, want):
    bad += not same_floats(dy.cpu().numpy().view(np.uint16), w)
x = to_fdt(gaussian(32 * 70000, 9), g.F16)                                           # host pipeline, pageable
bad += not np.array_equal(g.quantize(g.Q8_0, x, g.F16), O.quantize(g.Q8_0, g.F16, x))
print("sanitize_small:", "OK" if not bad else f"{bad} MISMATCHES")
sys.exit(1 if bad else 0)
