"""Parity soak: random (type, float side, block count, pointer offsets, input family) cases through the device API for a
time budget, every result compared with the CPU oracle bit for bit.  Input families go after rounding boundaries rather
than typical weights: values a few ulps either side of every code boundary of the block's own scale, rows with +max and
-max (tie path), constant rows, signed zeros, subnormals, coarse grids (bf16-representable values: ties everywhere),
heavy tails, NaN / inf sprinkles (legacy types).  usage: python tools/soak.py [seconds] [seed]; SOAK_CASES=N runs exactly N
cases instead (tests/test_parity_gpu.py::test_adversarial_input_families does), SOAK_TYPES / SOAK_FAMILIES restrict the draw.
Round 2's first run of this tool found the two K-quant deviations on subnormal weights that quant_k_kernel.cuh now handles
(an overflowing iscale, and inf * 0 in the scale codes)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import gguf_b200 as g
from data import BF16, F16, F32, same_blocks, same_floats, to_fdt
from oracle import oracle as O

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
LEGACY, Q8K, KQ = [2, 3, 6, 7, 8, 9], [15], [10, 11, 12, 13, 14]
st = torch.cuda.current_stream().cuda_stream
LEVELS = {2: 8, 3: 15, 6: 16, 7: 31, 8: 127, 9: 127, 15: 127, 10: 3, 11: 4, 12: 15, 13: 31, 14: 32}


def ulp_jitter(x, k):
    """move the magnitude of every nonzero element by a random number of f32 ulps in [-k, k] (zeros stay zeros: one ulp
    below +0 would be a NaN bit pattern)"""
    u = x.view(np.int32).copy()
    nz = (u & 0x7FFFFFFF) > 16
    u[nz] += rng.integers(-k, k + 1, int(nz.sum())).astype(np.int32)
    return u.view(np.float32)


def family(name, n, nb, ty):
    if name == "gauss":
        return (rng.standard_normal(n * nb) * float(rng.choice([1e-3, 0.02, 1.0, 40.0]))).astype(np.float32)
    if name == "student":
        return (rng.standard_t(3, n * nb) * 0.02).astype(np.float32)
    if name == "boundary":  # x = max * (code + 0.5) / levels, jittered by a few ulps: every element next to a rounding boundary
        L = LEVELS[ty]
        x = np.empty((nb, n), np.float32)
        for b in range(nb):
            m = np.float32(rng.standard_normal() * 0.05 + 0.1) * np.float32(rng.choice([-1, 1]))
            codes = rng.integers(-L, L, n).astype(np.float32) + np.float32(rng.choice([0.5, 0.0, 0.25]))
            x[b] = m * codes / np.float32(L)
            x[b, rng.integers(0, n)] = m
        return ulp_jitter(x.reshape(-1), int(rng.choice([0, 1, 3])))
    if name == "ties":  # +max and -max in the same row, random order; coarse grid so that ties abound
        x = np.round(rng.standard_normal((nb, n)) * 4).astype(np.float32) * np.float32(rng.choice([2.0 ** -6, 0.125, 1.0]))
        for b in range(nb):
            i, j = rng.choice(n, 2, replace=False)
            m = np.abs(x[b]).max() + np.float32(1.0) * np.float32(rng.choice([0, 2.0 ** -6]))
            x[b, i], x[b, j] = m, -m
        return x.reshape(-1)
    if name == "const":
        return np.repeat((rng.standard_normal(nb) * 0.1).astype(np.float32), n) * (rng.random(n * nb) < 0.97)
    if name == "zeros":
        x = np.zeros(n * nb, np.float32)
        x[rng.random(n * nb) < 0.3] = -0.0
        x[rng.random(n * nb) < 0.02] = np.float32(rng.standard_normal() * 1e-3)
        return x
    if name == "subnormal":
        return (rng.standard_normal(n * nb) * float(rng.choice([1e-41, 1e-39, 6e-8, 1e-30]))).astype(np.float32)
    if name == "nonfinite":
        x = (rng.standard_t(3, n * nb) * 0.02).astype(np.float32)
        x[rng.random(x.size) < 0.01] = np.nan
        x[rng.random(x.size) < 0.005] = np.inf
        x[rng.random(x.size) < 0.005] = -np.inf
        return x
    raise ValueError(name)


TYPES = [int(t) for t in os.environ["SOAK_TYPES"].split(",")] if os.environ.get("SOAK_TYPES") else LEGACY + Q8K + KQ
ONLY_FAM = os.environ.get("SOAK_FAMILIES", "").split(",") if os.environ.get("SOAK_FAMILIES") else None
t_end = time.time() + budget
max_cases = int(os.environ.get("SOAK_CASES", "0"))
cases = fails = 0
by_family = {}
while (cases < max_cases) if max_cases else (time.time() < t_end):
    ty = int(rng.choice(TYPES))
    fdt = int(rng.choice([F32, F16, BF16]))
    n, b = O.block_info(ty)
    big = rng.random() < 0.08
    nb = int(rng.integers(1, 40000 if n == 32 else 5000)) if big else int(rng.choice([1, 2, 3, 7, 8, 9, 31, 63, 64, 65, 127, 128, 129, 255, 257, 511, 513, 1025, 2049] if n == 32 else [1, 2, 3, 7, 8, 9, 15, 17, 31, 33, 63, 65, 129, 300]))
    fams = ["gauss", "student", "boundary", "ties", "const", "zeros", "subnormal"] + (["nonfinite"] if ty in LEGACY + Q8K else [])
    if ONLY_FAM:
        fams = [f for f in fams if f in ONLY_FAM]
    fam = str(rng.choice(fams))
    x = to_fdt(np.ascontiguousarray(family(fam, n, nb, ty), dtype=np.float32), fdt)
    foff = int(rng.choice([0, 4, 8, 16])) if fdt == F32 else int(rng.choice([0, 2, 6, 16]))
    poff = int(rng.choice([0, 2, 4, 16]))
    want_q = O.quantize(ty, fdt, x, threads=8)
    src = torch.zeros(x.nbytes + 64, dtype=torch.uint8, device="cuda")
    src[foff:foff + x.nbytes] = torch.from_numpy(x.view(np.uint8)).cuda()
    q = torch.zeros(nb * b + 64, dtype=torch.uint8, device="cuda")
    g.quantize_slice_device(ty, fdt, q.data_ptr() + poff, nb, src.data_ptr() + foff, n * nb, st)
    d = torch.zeros(x.nbytes + 64, dtype=torch.uint8, device="cuda")
    g.dequantize_slice_device(ty, fdt, d.data_ptr() + foff, n * nb, q.data_ptr() + poff, nb, st)
    torch.cuda.synchronize()
    gq = q.cpu().numpy()[poff:poff + nb * b]
    ok_q = same_blocks(gq, want_q, ty, b)
    want_d = O.dequantize(ty, fdt, want_q, threads=8)
    gd = d.cpu().numpy()[foff:foff + x.nbytes].view(want_d.dtype)
    ok_d = (not ok_q) or same_floats(gd, want_d)
    cases += 1
    s = by_family.setdefault(fam, [0, 0])
    s[0] += 1
    if not (ok_q and ok_d):
        fails += 1
        s[1] += 1
        if fails <= 10:
            bad = np.flatnonzero((gq.reshape(nb, b) != want_q.reshape(nb, b)).any(1)) if not ok_q else []
            if len(bad):  # keep the first differing blocks for an offline look: input elements, both encodings
                k = bad[:4]
                os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                np.savez(os.path.join(ROOT, "gpurun_out", f"soak_fail_{fails}.npz"), ty=ty, fdt=fdt, fam=fam, blocks=k,
                         x=x.reshape(nb, -1)[k], got=gq.reshape(nb, b)[k], want=want_q.reshape(nb, b)[k])
            print(f"MISMATCH case {cases}: type {g.TYPE_NAMES[ty]} fdt {g.TYPE_NAMES[fdt]} nb {nb} family {fam} foff {foff} poff {poff} quant_ok {ok_q} dequant_ok {ok_d} bad blocks {list(bad[:5])}", flush=True)
print(f"soak: {cases} cases ({'fixed count' if max_cases else f'{budget:.0f} s'}), seed {seed}: {fails} mismatches; per family (cases, mismatches): {by_family}")
sys.exit(1 if fails else 0)
