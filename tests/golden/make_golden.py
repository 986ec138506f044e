#!/usr/bin/env python
"""Generates tests/golden/*.npz — vectors that pin the CPU oracle WITHOUT the oracle.

Sources (none of them is oracle/ or gguf_b200/):
  * readme_block.npz — the reference's README example block (ggml-quants/README.md:32-37),
    x[i] = 0.1*(i+1), with the packed bytes derived by hand-restating the Rust in numpy float32
    during the survey.
  * gguf_py_legacy.npz — seeded inputs + blocks produced by the independent gguf-py 0.19.0 numpy
    quantizers (gguf/quants.py:220-402) for Q4_0 Q4_1 Q5_0 Q5_1 Q8_0.  gguf-py implements upstream
    ggml semantics, which equal the reference's on every block that is not all-zero,
    so inputs contain no all-zero block.
  * gguf_py_kdequant.npz — random packed K-quant blocks (finite f16 fields) + gguf-py's f32 dequant
    (gguf/quants.py:404-572) for Q2_K..Q6_K.
The reference itself (Rust) cannot be run in this image (no cargo/rustc); it has no golden vectors.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os

import numpy as np
from gguf import GGMLQuantizationType as T
from gguf import quants as GQ

HERE = os.path.dirname(os.path.abspath(__file__))

# ---- README block (App. C.1) ----
x = np.array([0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4, 1.5, 1.6, 1.7, 1.8, 1.9, 2.0, 2.1, 2.2,
              2.3, 2.4, 2.5, 2.6, 2.7, 2.8, 2.9, 3.0, 3.1, 3.2], dtype=np.float32)
kat = {
    "x": x,
    "q8_0": np.frombuffer(bytes.fromhex("732604080c1014181c2024282c3034383c4043474b4f53575b5f63676b6f73777b7f"), np.uint8),
    "q4_0": np.frombuffer(bytes.fromhex("66b648483737373726262626151515150404"), np.uint8),
    "q4_1": np.frombuffer(bytes.fromhex("9d32662e80809191a2a2b3b3c4c4d5d5e6e6f7f7"), np.uint8),
    "q5_0": np.frombuffer(bytes.fromhex("66b201000000807f7f6e6e5d5d4c4c3b3b2a2a191908"), np.uint8),
    "q8_1": np.frombuffer(bytes.fromhex("73269a5204080c1014181c2024282c3034383c4043474b4f53575b5f63676b6f73777b7f"), np.uint8),
}
np.savez_compressed(os.path.join(HERE, "readme_block.npz"), **kat)

# ---- gguf-py legacy quantizers ----
rng = np.random.default_rng(20261018)
parts = [rng.standard_normal(32 * 64).astype(np.float32) * np.float32(s) for s in (1e-4, 0.02, 1.0, 100.0)]
parts.append(rng.random(32 * 64, dtype=np.float32))                       # [0,1) like the reference's own tests
parts.append((rng.standard_t(3, 32 * 64) * 0.02).astype(np.float32))     # heavy tails
xin = np.concatenate(parts)
out = {"x": xin}
for name, gt in [("q4_0", T.Q4_0), ("q4_1", T.Q4_1), ("q5_0", T.Q5_0), ("q5_1", T.Q5_1), ("q8_0", T.Q8_0)]:
    blocks = GQ.quantize(xin.reshape(-1, 32), gt)
    out[name] = np.ascontiguousarray(blocks).reshape(-1)
    out[name + "_deq"] = GQ.dequantize(blocks, gt).reshape(-1).astype(np.float32)
np.savez_compressed(os.path.join(HERE, "gguf_py_legacy.npz"), **out)

# ---- gguf-py K-quant dequant ----
kd = {}
for name, gt, size, foffs in [("q2k", T.Q2_K, 84, [80, 82]), ("q3k", T.Q3_K, 110, [108]), ("q4k", T.Q4_K, 144, [0, 2]),
                              ("q5k", T.Q5_K, 176, [0, 2]), ("q6k", T.Q6_K, 210, [208])]:
    blk = rng.integers(0, 256, size=(48, size), dtype=np.uint8)
    for o in foffs:
        h = (rng.standard_normal(48) * 0.01).astype(np.float16).view(np.uint16)
        blk[:, o] = h & 0xFF
        blk[:, o + 1] = h >> 8
    kd[name] = blk.reshape(-1)
    kd[name + "_deq"] = GQ.dequantize(blk, gt).reshape(-1).astype(np.float32)
np.savez_compressed(os.path.join(HERE, "gguf_py_kdequant.npz"), **kd)
print("written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
