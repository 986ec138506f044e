"""N > 1 host logic on CPU, against the SHIPPED partitioner: `ggq_plan_shards` (include/ggq.h) is the rule
`ggq_slices` / the slice calls / `ggq_cast` split their work by after `ggq_set_shard_devices(n)`; it is a
pure host function, so it is tested here without a GPU.  The world_size-2 gloo run has each rank compute
exactly the pieces the library assigns to "its" device (compute = the CPU oracle, this is a test) and
checks that the gathered bytes equal the single-process result — sharding never changes the bytes
(the reference's blocks are independent: ggml-quants/src/lib.rs:129-131)."""
import ctypes
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BLOCK = {2: (32, 18), 3: (32, 20), 6: (32, 22), 7: (32, 24), 8: (32, 34), 9: (32, 36), 10: (256, 84), 11: (256, 110),
         12: (256, 144), 13: (256, 176), 14: (256, 210), 15: (256, 290)}
FSIZE = {0: 4, 1: 2, 30: 2}


def plan(jobs, ndev):
    """jobs: list of (quantize, type, fdt, n_elems) -> list of (job, device, e0, e1) from libggq."""
    from gguf_b200._lib import ShardPiece, SliceJob, lib
    arr = (SliceJob * len(jobs))()
    for i, (q, ty, fdt, n) in enumerate(jobs):
        e, _ = BLOCK[ty]
        arr[i] = SliceJob(ty, fdt, int(q), None, n // e if q else n, None, n if q else n // e)
    n = lib().ggq_plan_shards(arr, len(jobs), ndev, None, 0)
    out = (ShardPiece * max(n, 1))()
    assert lib().ggq_plan_shards(arr, len(jobs), ndev, out, n) == n
    return [(p.job, p.device, p.elem_begin, p.elem_end) for p in out[:n]]


def weight(job, e0, e1):
    q, ty, fdt, _ = job
    e, b = BLOCK[ty]
    return (e1 - e0) // e * b + (e1 - e0) * FSIZE[fdt]


BENCH_JOBS = [(0, ty, 1, n) for ty in (2, 8, 12, 14) for n in (4096 * 14336, 4096 * 4096)]   # bench.py's step


def test_plan_covers_every_element_once_and_cuts_are_aligned(ggq):
    jobs = BENCH_JOBS + [(1, 13, 0, 256 * 4099), (1, 2, 30, 32 * 5), (0, 15, 0, 0), (1, 9, 1, (1 << 20) * 3 + 32)]
    for ndev in (1, 2, 3, 4, 8):
        pieces = plan(jobs, ndev)
        per_job = {}
        for j, d, e0, e1 in pieces:
            assert 0 <= d < ndev and e0 < e1
            per_job.setdefault(j, []).append((e0, e1))
        for j, job in enumerate(jobs):
            n = job[3]
            rs = per_job.get(j, [])
            assert sum(b - a for a, b in rs) == n
            at = 0
            for a, b in rs:                       # ascending, contiguous, cuts on 2^20-element units
                assert a == at and (a % (1 << 20) == 0)
                at = b
            assert at == n
        # devices appear in non-decreasing order along the job stream (contiguous ranges per device)
        devs = [d for _, d, _, _ in pieces]
        assert devs == sorted(devs)


def test_plan_balances_pcie_bytes(ggq):
    for ndev in (2, 4, 8):
        pieces = plan(BENCH_JOBS, ndev)
        load = [0] * ndev
        for j, d, e0, e1 in pieces:
            load[d] += weight(BENCH_JOBS[j], e0, e1)
        total = sum(load)
        unit = max(weight(j, 0, 1 << 20) for j in BENCH_JOBS)
        assert max(load) - total / ndev <= unit, (ndev, load)      # within one unit of perfect
        assert min(load) > 0


def test_plan_small_calls_use_fewer_devices_and_is_deterministic(ggq):
    small = [(1, 8, 1, 32 * 1000)]
    assert {d for _, d, _, _ in plan(small, 8)} == {0}
    mid = [(0, 2, 1, 1 << 24)]                                     # 43 MB of traffic: at most 2 devices' worth
    assert len({d for _, d, _, _ in plan(mid, 8)}) == 2
    assert plan(BENCH_JOBS, 8) == plan(BENCH_JOBS, 8)
    # a job that fails the slice length checks yields no plan
    from gguf_b200._lib import SliceJob, lib
    bad = (SliceJob * 1)(SliceJob(2, 1, 1, None, 3, None, 33))
    assert lib().ggq_plan_shards(bad, 1, 2, None, 0) == 0


WORKER = r'''
import os, sys, hashlib, json
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
from test_sharding import plan, BLOCK
from oracle import oracle as O
from data import gaussian, to_fdt
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=int(sys.argv[4]))
rank, world = dist.get_rank(), dist.get_world_size()
jobs = [(1, 2, 1, (1 << 20) * 9 + 32 * 77), (1, 14, 1, 256 * 9000), (1, 8, 1, (1 << 20) * 3)]
mine = []
for j, d, e0, e1 in plan(jobs, world):
    if d != rank: continue
    q, ty, fdt, n = jobs[j]
    x = to_fdt(gaussian(n, 100 + j), fdt)[e0:e1]
    mine.append((j, e0, O.quantize(ty, fdt, x).tobytes()))
gathered = [None] * world
dist.all_gather_object(gathered, mine)
t = torch.tensor([0.1 * (rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    parts = sorted(p for g in gathered for p in g)
    sha = {}
    for j in range(len(jobs)):
        sha[str(j)] = hashlib.sha256(b"".join(b for jj, _, b in parts if jj == j)).hexdigest()
    print(json.dumps({"sha": sha, "tmax": float(t.item()), "ranks_with_work": sum(1 for g in gathered if g)}))
dist.barrier(); dist.destroy_process_group()
'''


def test_world2_gloo_pieces_of_the_shipped_plan_equal_single_process(tmp_path, oracle, ggq):
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
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    jobs = [(1, 2, 1, (1 << 20) * 9 + 32 * 77), (1, 14, 1, 256 * 9000), (1, 8, 1, (1 << 20) * 3)]
    for j, (q, ty, fdt, n) in enumerate(jobs):
        want = hashlib.sha256(oracle.quantize(ty, fdt, to_fdt(gaussian(n, 100 + j), fdt)).tobytes()).hexdigest()
        assert res["sha"][str(j)] == want
    assert res["ranks_with_work"] == 2 and abs(res["tmax"] - 0.2) < 1e-9
