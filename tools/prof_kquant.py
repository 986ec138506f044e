import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gguf_b200 as g
n = 4096 * 4096
st = torch.cuda.current_stream().cuda_stream
x = (torch.randn(n, device="cuda") * 0.02).to(torch.float16)
for ty in [int(t) for t in sys.argv[1].split(",")]:
    e, b = g.block_info(ty)
    packed = torch.empty(n // e * b, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        g.quantize_slice_device(ty, g.F16, packed, n // e, x, n, st)
    torch.cuda.synchronize()
print("ok")
