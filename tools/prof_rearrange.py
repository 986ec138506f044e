"""Launch the rearrangement kernel on the operator shapes a few times — the target of the ncu capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gguf_b200 as g
from gguf_b200.rearrange import block_layout, permute_qk_layouts, rearrange_device, type_size

st = torch.cuda.current_stream().cuda_stream
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2


def nbytes(ty, shape):
    be, bb = type_size(ty)
    n = 1
    for d in shape:
        n *= d
    return n // be * bb


for ty, shape, nh in [(g.F16, (8192, 8192), 64), (g.Q4_0, (8192, 8192), 64), (g.Q8_0, (4096, 4096), 32)]:
    n = nbytes(ty, shape)
    src = torch.empty(n, dtype=torch.uint8, device="cuda").random_(0, 256)
    dst = torch.empty_like(src)
    dl, sl, unit = permute_qk_layouts(ty, shape, nh)
    for _ in range(reps):
        rearrange_device(dst.data_ptr(), dl, src.data_ptr(), sl, unit, st)
# Mixtral expert merge, F16
part, whole = (4096, 14336, 8), (4096, 28672, 8)
n = nbytes(g.F16, part)
gate = torch.empty(n, dtype=torch.uint8, device="cuda").random_(0, 256)
out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
wl, unit = block_layout(g.F16, whole)
views = wl.split(1, [14336, 14336])
pl, _ = block_layout(g.F16, part)
for _ in range(reps):
    rearrange_device(out.data_ptr(), views[0], gate.data_ptr(), pl, unit, st)
torch.cuda.synchronize()
print("ok")
