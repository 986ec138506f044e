"""Multi-GPU partitioning of the codec path: blocks are independent, so work is split
by tensor (largest-first greedy) and, inside a large tensor, by contiguous block range.  There is no
collective on the data path; `torch.distributed` is only used by harnesses to agree on timings."""


def assign_tensors(sizes, world):
    """LPT greedy: returns `world` lists of tensor indices, largest tensors placed first on the least
    loaded rank.  Deterministic (ties -> lowest rank, lowest index)."""
    loads = [0] * world
    out = [[] for _ in range(world)]
    for i in sorted(range(len(sizes)), key=lambda i: (-sizes[i], i)):
        r = min(range(world), key=lambda r: (loads[r], r))
        out[r].append(i)
        loads[r] += sizes[i]
    return out


def split_block_range(n_blocks, parts, align=8):
    """Contiguous [begin, end) block ranges, boundaries at multiples of `align` blocks so every shard of
    a 16-byte aligned tensor stays 16-byte aligned.  Empty ranges are dropped."""
    units = (n_blocks + align - 1) // align
    out = []
    for k in range(parts):
        b = min(n_blocks, units * k // parts * align)
        e = min(n_blocks, units * (k + 1) // parts * align)
        if e > b:
            out.append((b, e))
    return out


def max_over_ranks(value, dist=None, device=None):
    """Max of a python float over all ranks (identity when not distributed)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
