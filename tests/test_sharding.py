"""N > 1 host logic on CPU: partitioning helpers and a world_size-2 gloo run in which each rank
processes its share of the tensors (compute = the CPU oracle, this is a test) and the gathered result
must equal the single-process result byte for byte — i.e. sharding never changes the bytes."""
import hashlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_assign_tensors_lpt():
    from gguf_b200.sharding import assign_tensors
    sizes = [58, 58, 58, 16, 16, 4, 4, 128, 128]
    for world in (1, 2, 4, 8):
        parts = assign_tensors(sizes, world)
        assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(sizes)
    assert assign_tensors(sizes, 2) == assign_tensors(sizes, 2)


def test_split_block_range_alignment_and_cover():
    from gguf_b200.sharding import split_block_range
    for n in (0, 1, 7, 8, 9, 1000, 1835008):
        for parts in (1, 2, 3, 8):
            r = split_block_range(n, parts)
            assert sum(e - b for b, e in r) == n
            for (b, e), nxt in zip(r, r[1:] + [(n, n)]):
                assert e == nxt[0] and b % 8 == 0


WORKER = r'''
import os, sys, hashlib, json
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
from gguf_b200.sharding import assign_tensors, split_block_range, max_over_ranks
from oracle import oracle as O
from data import gaussian, to_fdt
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
rank, world = dist.get_rank(), dist.get_world_size()
shapes = [(2, 32 * 700), (8, 32 * 300), (12, 256 * 90), (14, 256 * 40), (3, 32 * 64), (8, 32 * 1201)]
mine = assign_tensors([n for _, n in shapes], world)[rank]
digests = {}
for i in mine:
    ty, n = shapes[i]
    q = O.quantize(ty, 1, to_fdt(gaussian(n, 100 + i), 1))
    digests[i] = hashlib.sha256(q.tobytes()).hexdigest()
# one large tensor split by block range: every rank quantizes its contiguous range
ty, n = 8, 32 * 5000
x = to_fdt(gaussian(n, 999), 1)
rngs = split_block_range(n // 32, world)
b, e = rngs[rank] if rank < len(rngs) else (0, 0)
part = O.quantize(ty, 1, x[b * 32:e * 32]).tobytes()
gathered = [None] * world
dist.all_gather_object(gathered, (digests, b, part))
t = max_over_ranks(0.1 * (rank + 1), dist)
if rank == 0:
    allg = {}
    for d, _, _ in gathered: allg.update(d)
    whole = b"".join(p for _, _, p in sorted(gathered, key=lambda g: g[1]))
    print(json.dumps({"digests": {str(k): v for k, v in allg.items()}, "range_sha": hashlib.sha256(whole).hexdigest(), "tmax": t}))
dist.barrier(); dist.destroy_process_group()
'''


def test_world2_gloo_sharded_equals_single_process(tmp_path, oracle):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from data import gaussian, to_fdt
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = 29650 + os.getpid() % 200
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    import json
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    shapes = [(2, 32 * 700), (8, 32 * 300), (12, 256 * 90), (14, 256 * 40), (3, 32 * 64), (8, 32 * 1201)]
    for i, (ty, n) in enumerate(shapes):
        want = hashlib.sha256(oracle.quantize(ty, 1, to_fdt(gaussian(n, 100 + i), 1)).tobytes()).hexdigest()
        assert res["digests"][str(i)] == want
    whole = oracle.quantize(8, 1, to_fdt(gaussian(32 * 5000, 999), 1))
    assert res["range_sha"] == hashlib.sha256(whole.tobytes()).hexdigest()
    assert abs(res["tmax"] - 0.2) < 1e-9
