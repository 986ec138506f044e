"""Launch each headline dequant kernel a few times (FFN shape) — the target of the ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gguf_b200 as g

n = 4096 * 14336
types = [int(t) for t in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2, 8, 12, 14]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
direction = sys.argv[3] if len(sys.argv) > 3 else "dequant"
st = torch.cuda.current_stream().cuda_stream
x = (torch.randn(n, device="cuda") * 0.02).to(torch.float16)
for ty in types:
    e, b = g.block_info(ty)
    packed = torch.empty(n // e * b, dtype=torch.uint8, device="cuda")
    g.quantize_slice_device(ty, g.F16, packed, n // e, x, n, st)
    out = torch.empty(n, dtype=torch.float16, device="cuda")
    torch.cuda.synchronize()
    for _ in range(reps):
        if direction == "dequant":
            g.dequantize_slice_device(ty, g.F16, out, n, packed, n // e, st)
        else:
            g.quantize_slice_device(ty, g.F16, packed, n // e, x, n, st)
    torch.cuda.synchronize()
print("ok")
