"""Raw PCIe ceilings on the box: pinned<->device copies alone and concurrently, 16 MB and 256 MB."""
import torch, time
dev = torch.device("cuda")
def bw(fn, nbytes, reps=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t) / 1e9
for mb in (16, 64, 256):
    n = mb << 20
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a, d_b = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def h2d():
        with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
    def d2h():
        with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    def both(): h2d(); d2h()
    print(f"{mb:4d} MB  H2D {bw(h2d, n):6.1f} GB/s   D2H {bw(d2h, n):6.1f} GB/s   concurrent sum {bw(both, 2 * n):6.1f} GB/s", flush=True)
