"""Device-resident GB/s of the rearrangement kernel (rearrange.cu) on the shapes the operators produce:
permute-qk of attn_q / attn_k (8B and 70B class), the 3-D expert merge of Mixtral, a split of a merged
attn_qkv.  Algorithmic bytes = 2 x tensor bytes (read once, write once); buffers rotate over > 126 MB.
Writes gpurun_out/rearrange_sweep.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gguf_b200 as g
from gguf_b200.rearrange import ArrayLayout, block_layout, permute_qk_layouts, rearrange_device, type_size

PEAK = 6543.4
st = torch.cuda.current_stream().cuda_stream
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


def nbytes(ty, shape):
    be, bb = type_size(ty)
    n = 1
    for d in shape:
        n *= d
    return n // be * bb


def graph_time(jobs, nsets, reps):
    """Device-side rate: the same launches captured once into a CUDA graph (no per-call host overhead)."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        global st
        keep, st = st, s.cuda_stream
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for i in range(reps):
                jobs(i % nsets)
        st = keep
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def run(label, jobs, total_bytes):
    """jobs(i) enqueues the launches of one operator application on buffer set i."""
    nsets = max(2, int(300e6 // total_bytes) + 1)
    t = timeit(jobs, nsets, 40)
    if t < 30e-6:  # launch-bound through ctypes: report the graph-replayed rate
        t = graph_time(jobs, nsets, 40)
        label += " [graph]"
    gbs = 2 * total_bytes / t / 1e9
    rows.append({"case": label, "us": t * 1e6, "GBps": gbs, "frac_measured_peak": gbs / PEAK, "tensor_MB": total_bytes / 1e6})
    print(f"{label:58s} {t*1e6:9.1f} us {gbs:8.1f} GB/s  {100*gbs/PEAK:5.1f}% of measured peak", flush=True)


def bufs(n, count):
    return [torch.empty(n, dtype=torch.uint8, device="cuda").random_(0, 256) for _ in range(count)]


for ty, name in [(g.F16, "F16"), (g.F32, "F32"), (g.Q8_0, "Q8_0"), (g.Q4_0, "Q4_0"), (g.Q4K, "Q4K"), (g.Q6K, "Q6K")]:
    for shape, nh, what in [((4096, 4096), 32, "attn_q 8B"), ((4096, 1024), 8, "attn_k 8B"), ((8192, 8192), 64, "attn_q 70B")]:
        n = nbytes(ty, shape)
        nsets = max(2, int(300e6 // n) + 1)
        src, dst = bufs(n, nsets), bufs(n, nsets)
        dl, sl, unit = permute_qk_layouts(ty, shape, nh)
        run(f"permute-qk {what} {name} {shape}", lambda i: rearrange_device(dst[i % nsets].data_ptr(), dl, src[i % nsets].data_ptr(), sl, unit, st), n)
        del src, dst

# Mixtral expert merge: gate/up [4096,14336,8] -> [4096,28672,8]
for ty, name in [(g.F16, "F16"), (g.Q8_0, "Q8_0"), (g.Q4K, "Q4K")]:
    part, whole = (4096, 14336, 8), (4096, 28672, 8)
    n = nbytes(ty, part)
    gate, up = bufs(n, 2), bufs(n, 2)
    out = bufs(2 * n, 2)
    wl, unit = block_layout(ty, whole)
    views = wl.split(1, [14336, 14336])
    pl, _ = block_layout(ty, part)

    def job(i, views=views, pl=pl, unit=unit, gate=gate, up=up, out=out):
        rearrange_device(out[i % 2].data_ptr(), views[0], gate[i % 2].data_ptr(), pl, unit, st)
        rearrange_device(out[i % 2].data_ptr(), views[1], up[i % 2].data_ptr(), pl, unit, st)
    run(f"merge ffn_gate_up_exps {name} 2 x {part}", job, 2 * n)
    del gate, up, out

# split of a merged attn_qkv [4096, 6144] into q / k / v (contiguous sub-ranges: plain copies)
for ty, name in [(g.F16, "F16"), (g.Q4_0, "Q4_0")]:
    whole = (4096, 6144)
    n = nbytes(ty, whole)
    src, dst = bufs(n, 6), bufs(n, 6)
    wl, unit = block_layout(ty, whole)
    views = wl.split(1, [4096, 1024, 1024])
    offs = [0, nbytes(ty, (4096, 4096)), nbytes(ty, (4096, 5120))]

    def job(i, views=views, unit=unit, src=src, dst=dst, offs=offs, ty=ty):
        for v, o, r in zip(views, offs, (4096, 1024, 1024)):
            rearrange_device(dst[i % 6].data_ptr() + o, block_layout(ty, (4096, r))[0], src[i % 6].data_ptr(), v, unit, st)
    run(f"split attn_qkv {name} {whole}", job, n)
    del src, dst

os.makedirs("gpurun_out", exist_ok=True)
json.dump({"peak_GBps": PEAK, "rows": rows}, open("gpurun_out/rearrange_sweep.json", "w"), indent=1)
