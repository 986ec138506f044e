import sys, os
sys.path.insert(0, "/root/repo")
import torch
import gguf_b200 as g
from gguf_b200.rearrange import ArrayLayout, rearrange_device
st = torch.cuda.current_stream().cuda_stream
for mb in (117, 235, 940):
    n = mb * (1 << 20)
    bufs = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(max(2, 600 // mb))]
    src = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    dl = ArrayLayout.new_contiguous([n // 16], 16)
    sl = ArrayLayout([n // 16], [0])
    def fn(i): rearrange_device(bufs[i % len(bufs)].data_ptr(), dl, src.data_ptr(), sl, 16, st)
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 40
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps * 1e-3
    print(f"write-only {mb} MiB: {t*1e6:.1f} us  {n/t/1e9:.1f} GB/s")
    # copy 1:1
    srcb = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(len(bufs))]
    cl = ArrayLayout.new_contiguous([n // 16], 16)
    def fc(i): rearrange_device(bufs[i % len(bufs)].data_ptr(), dl, srcb[i % len(bufs)].data_ptr(), cl, 16, st)
    for i in range(3): fc(i)
    torch.cuda.synchronize()
    e0.record()
    for i in range(reps): fc(i)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps * 1e-3
    print(f"copy 1:1  {mb} MiB: {t*1e6:.1f} us  {2*n/t/1e9:.1f} GB/s")
    del bufs, srcb
