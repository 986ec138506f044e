"""BASELINE.json configs[3]: Mixtral-8x7B-shaped expert tensors (4096x14336x8 = 469.8 M elements each),
dequant sweep over all 12 block types -> f16, sharded BY TENSOR over the ranks (torchrun, one GPU per rank,
no collective on the data path).  Each rank owns its share of `n_tensors` expert tensors per type; aggregate
GB/s = bytes of all ranks / max-over-ranks time."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import gguf_b200 as g


def max_over_ranks(value, dist=None, device=None):
    if dist is None:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
PEAK = 6543.4
n = 4096 * 14336 * 8
n_tensors = 8 * world if len(sys.argv) < 2 else int(sys.argv[1])   # global tensor count, sharded by tensor
mine = list(range(rank, n_tensors, world))   # equal-sized tensors: by-tensor sharding is round-robin
st = torch.cuda.current_stream().cuda_stream
x = (torch.randn(n, device=dev) * 0.02).to(torch.float16)
out = [torch.empty(n, dtype=torch.float16, device=dev) for _ in range(2)]
rows = []
for ty in [g.Q4_0, g.Q4_1, g.Q5_0, g.Q5_1, g.Q8_0, g.Q8_1, g.Q2K, g.Q3K, g.Q4K, g.Q5K, g.Q6K, g.Q8K]:
    e, b = g.block_info(ty)
    packed = [torch.empty(n // e * b, dtype=torch.uint8, device=dev) for _ in range(2)]
    g.quantize_slice_device(ty, g.F16, packed[0], n // e, x, n, st)
    packed[1].copy_(packed[0])
    for i in range(2):
        g.dequantize_slice_device(ty, g.F16, out[i], n, packed[i], n // e, st)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k, _ in enumerate(mine):
        g.dequantize_slice_device(ty, g.F16, out[k % 2], n, packed[k % 2], n // e, st)
    e1.record(); torch.cuda.synchronize()
    t = max_over_ranks(e0.elapsed_time(e1) * 1e-3, dist if world > 1 else None, dev)
    nbytes = (n // e * b + n * 2) * n_tensors
    if rank == 0:
        rows.append({"type": g.TYPE_NAMES[ty], "n_gpus": world, "tensors": n_tensors, "seconds": t, "aggregate_GBps": nbytes / t / 1e9,
                     "frac_of_aggregate_measured_peak": nbytes / t / 1e9 / (PEAK * world)})
        print(f"{g.TYPE_NAMES[ty]:5s} gpus={world} {n_tensors} tensors {t*1e3:8.2f} ms {nbytes/t/1e9:9.1f} GB/s  {100*nbytes/t/1e9/(PEAK*world):5.1f}% of {world}x peak", flush=True)
    del packed
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"mixtral_sweep_n{world}.json"), "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
