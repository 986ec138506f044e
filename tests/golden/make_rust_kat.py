#!/usr/bin/env python
"""Generates tests/golden/rust_kat.npz: known answers derived from tests/golden/rust_restatement.py (numpy float32,
one operation per Rust operation, no code shared with oracle/ or gguf_b200/) for the cases round 1 had no non-oracle
vector for: Q8K, Q5_1, Q8_1 (sum field), every type's f16 / bf16 float side, NaN / inf / tie / zero rows, and
dequantize -> f32 / f16 / bf16 of both quantized and random-byte blocks.
Run from the repo root:  python tests/golden/make_rust_kat.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import rust_restatement as R  # noqa: E402

TYPES = [2, 3, 6, 7, 8, 9, 15]
SIZE = {2: 18, 3: 20, 6: 22, 7: 24, 8: 34, 9: 36, 15: 290}
FIELDS = {2: [0], 3: [0, 2], 6: [0], 7: [0, 2], 8: [0], 9: [0, 2], 15: [0]}
rng = np.random.default_rng(20261019)


def inputs(n):
    rows = [np.arange(1, n + 1, dtype=np.float32) * np.float32(0.1)]                       # README block pattern
    for s in (1e-4, 0.02, 1.0, 300.0):
        rows += [(rng.standard_normal(n) * s).astype(np.float32) for _ in range(3)]
    rows.append(rng.random(n, dtype=np.float32))
    rows.append((rng.standard_t(3, n) * 0.02).astype(np.float32))
    z = np.zeros(n, np.float32)
    rows += [z.copy(), np.full(n, 0.37, np.float32), np.full(n, -2.5, np.float32)]
    a = z.copy(); a[5] = 1.0; a[20] = -1.0; rows.append(a)                                 # |x| tie: first wins (+)
    a = z.copy(); a[5] = -1.0; a[20] = 1.0; rows.append(a)                                 # (-)
    a = z.copy(); a[n - 1] = 3.0; rows.append(a)
    a = rng.standard_normal(n).astype(np.float32); a[3] = np.nan; rows.append(a)           # NaN element: Q4_0 -> 15, Q5_0 -> 0
    a = rng.standard_normal(n).astype(np.float32); a[7] = np.inf; rows.append(a)
    a = rng.standard_normal(n).astype(np.float32); a[9] = -np.inf; rows.append(a)
    rows.append((rng.standard_normal(n) * 1e-41).astype(np.float32))                       # f32 denormals
    rows.append((rng.standard_normal(n) * 1e6).astype(np.float32))                         # delta overflows f16
    rows.append(np.linspace(-1, 1, n, dtype=np.float32))                                   # exact .5 rounding cases
    rows.append((np.arange(n, dtype=np.float32) - n / 2) * np.float32(0.5))
    return np.concatenate(rows)


def widen(bits, fdt):
    if fdt == 1:
        return bits.view(np.float16).astype(np.float32)
    return (bits.astype(np.uint32) << 16).view(np.float32)


def narrow(x32, fdt):
    """half::{f16, bf16}::from_f32 incl. NaN: quiet bit set, top payload bits kept."""
    u = x32.view(np.uint32)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    if fdt == 1:
        with np.errstate(over="ignore", invalid="ignore"):
            r = x32.astype(np.float16).view(np.uint16)
        r[nan] = (((u[nan] >> 16) & 0x8000) | 0x7C00 | 0x0200 | ((u[nan] >> 13) & 0x03FF)).astype(np.uint16)
        return r
    r = ((u.astype(np.uint64) + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    r[nan] = ((u[nan] >> 16) | 0x40).astype(np.uint16)
    return r


out = {}
for ty in TYPES:
    n = R.QUANT[ty][0]
    x = inputs(n)
    out[f"x_{ty}"] = x
    out[f"q_{ty}_f32"] = R.quantize(ty, x)
    for fdt, name in ((1, "f16"), (30, "bf16")):
        with np.errstate(over="ignore", invalid="ignore"):
            bits = narrow(x, fdt)
        out[f"x_{ty}_{name}"] = bits
        out[f"q_{ty}_{name}"] = R.quantize(ty, widen(bits, fdt))      # lib.rs:66-69 / 82-84: widen, then the f32 quantizer
    # dequantize: the f32-side quantized blocks whose scales are finite, plus random bytes with finite scale fields
    q = out[f"q_{ty}_f32"].reshape(-1, SIZE[ty])
    blk = rng.integers(0, 256, size=(24, SIZE[ty]), dtype=np.uint8)
    for o in FIELDS[ty]:
        h = (rng.standard_normal(24) * 0.01).astype(np.float16).view(np.uint16)
        blk[:, o] = h & 0xFF
        blk[:, o + 1] = h >> 8
    fin = np.ones(len(q), bool)
    for o in FIELDS[ty]:
        hb = q[:, o].astype(np.uint16) | (q[:, o + 1].astype(np.uint16) << 8)
        fin &= (hb & 0x7C00) != 0x7C00
    blocks = np.concatenate([q[fin], blk]).reshape(-1)
    y = R.dequantize(ty, blocks)
    out[f"b_{ty}"] = blocks
    out[f"d_{ty}_f32"] = y
    out[f"d_{ty}_f16"] = narrow(y, 1)
    out[f"d_{ty}_bf16"] = narrow(y, 30)
np.savez_compressed(os.path.join(HERE, "rust_kat.npz"), **out)
print("rust_kat.npz:", {k: v.shape for k, v in list(out.items())[:6]}, "...", len(out), "arrays")
