"""Seeded synthetic inputs shared by the parity tests (float side and packed side)."""
import numpy as np

F32, F16, BF16 = 0, 1, 30
FIELDS_F16 = {  # byte offsets of the f16 fields inside each packed block
    2: [0], 3: [0, 2], 6: [0], 7: [0, 2], 8: [0], 9: [0, 2], 10: [80, 82], 11: [108], 12: [0, 2], 13: [0, 2], 14: [208], 15: [0],
}


def f32_to_bf16_bits(x):
    """RNE f32 -> bf16 bit patterns (numpy), NaN quieted like the `half` crate."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    r[nan] = ((u[nan] >> 16) | 0x40).astype(np.uint16)
    return r


def to_fdt(x32, fdt):
    """f32 values -> the float-side representation used by the APIs (f32, or uint16 bits)."""
    if fdt == F32:
        return np.ascontiguousarray(x32, dtype=np.float32)
    if fdt == F16:
        with np.errstate(over="ignore"):
            return np.ascontiguousarray(x32, dtype=np.float32).astype(np.float16).view(np.uint16)
    return f32_to_bf16_bits(x32)


def gaussian(n, seed, sigma=0.02):
    return (np.random.default_rng(seed).standard_normal(n) * sigma).astype(np.float32)


def edge_blocks(block, kquant_domain=False):
    """A battery of nasty `block`-element inputs (f32): zeros, ties, NaN/inf, denormals, signed zeros.

    `kquant_domain=True` keeps only rows inside upstream ggml's K-quant domain: finite values whose
    scale search never divides to inf/NaN (its `nearest_int` asserts |v| <= 4194303; NaN bits that
    reach it are platform-defined, so there is nothing to be bit-exact with).  The legacy quantizers
    are total functions in the reference and take every row."""
    rng = np.random.default_rng(7)
    rows = []
    z = np.zeros(block, np.float32)
    rows.append(z.copy())                                   # all +0  -> ZEROS block
    rows.append(-z)                                         # all -0
    rows.append(np.full(block, 0.37, np.float32))           # constant (min == max)
    rows.append(np.full(block, -2.5, np.float32))
    a = z.copy(); a[5] = 1.0; a[20 % block] = -1.0; rows.append(a)      # |x| tie, first wins (+)
    a = z.copy(); a[5] = -1.0; a[20 % block] = 1.0; rows.append(a)      # |x| tie, first wins (-)
    a = z.copy(); a[block - 1] = 3.0; rows.append(a)                    # max in the last lane
    a = z.copy(); a[0] = -3.0; rows.append(a)
    a = rng.standard_normal(block).astype(np.float32); a[3] = np.nan; rows.append(a)
    a = rng.standard_normal(block).astype(np.float32); a[:] = np.nan; rows.append(a)   # all NaN
    a = rng.standard_normal(block).astype(np.float32); a[7] = np.inf; rows.append(a)
    a = rng.standard_normal(block).astype(np.float32); a[9] = -np.inf; rows.append(a)
    a = rng.standard_normal(block).astype(np.float32); a[1] = np.inf; a[2] = -np.inf; rows.append(a)
    rows.append((rng.standard_normal(block) * 1e-41).astype(np.float32))              # f32 denormals
    a = z.copy(); a[4] = np.float32(1e-45); rows.append(a)                             # delta underflows to 0
    a = z.copy(); a[4] = np.float32(-1e-45); a[6] = np.float32(1e-45); rows.append(a)
    rows.append((rng.standard_normal(block) * 6e-8).astype(np.float32))               # f16-subnormal scales
    rows.append((rng.standard_normal(block) * 3e4).astype(np.float32))                # near f16 overflow
    rows.append((rng.standard_normal(block) * 1e6).astype(np.float32))                # delta overflows f16 -> inf
    a = z.copy(); a[0] = 0.0; a[1] = -0.0; a[2] = 1.0; rows.append(a)                  # +0 first among the minima
    a = z.copy(); a[0] = -0.0; a[1] = 0.0; a[2] = 1.0; rows.append(a)                  # -0 first
    a = np.abs(rng.standard_normal(block)).astype(np.float32); a[11] = -0.0; rows.append(a)
    rows.append(np.linspace(-1, 1, block, dtype=np.float32))                           # exact .5 rounding cases
    rows.append((np.arange(block, dtype=np.float32) - block / 2) * np.float32(0.5))
    rows.append(np.arange(1, block + 1, dtype=np.float32) * np.float32(0.1))
    if kquant_domain:
        keep = []
        for r in rows:
            nz = np.abs(r[r != 0]) if np.isfinite(r).all() else None
            if nz is not None and (nz.size == 0 or (nz.min() > 1e-30 and nz.max() < 1e4)):
                keep.append(r)
        rows = keep
    return np.concatenate(rows)


def random_packed(ty, nblocks, bytes_per_block, seed, wild=False):
    """Random packed blocks exercising every code value.  `wild` keeps raw random f16 fields
    (inf / NaN / subnormal scales); otherwise the fields are forced to small finite values."""
    rng = np.random.default_rng(seed)
    blk = rng.integers(0, 256, size=(nblocks, bytes_per_block), dtype=np.uint8)
    if not wild:
        for o in FIELDS_F16[ty]:
            h = (rng.standard_normal(nblocks) * 0.01).astype(np.float16).view(np.uint16)
            blk[:, o] = h & 0xFF
            blk[:, o + 1] = h >> 8
    return blk.reshape(-1)


def same_floats(a, b):
    """Bit equality, except that NaN matches NaN of any payload (arithmetic NaNs are platform-defined)."""
    a = np.asarray(a); b = np.asarray(b)
    if a.dtype == np.float32:
        ua, ub = a.view(np.uint32), b.view(np.uint32)
        na, nb = (ua & 0x7FFFFFFF) > 0x7F800000, (ub & 0x7FFFFFFF) > 0x7F800000
    else:
        ua, ub = a.view(np.uint16), b.view(np.uint16)
        na, nb = (ua & 0x7FFF) > 0x7C00, (ub & 0x7FFF) > 0x7C00
    return bool(np.all((ua == ub) | (na & nb)))


def same_blocks(a, b, ty, bytes_per_block):
    """Byte equality of packed blocks; the f16 header fields compare NaN ~ NaN."""
    a = np.asarray(a, np.uint8).reshape(-1, bytes_per_block).copy()
    b = np.asarray(b, np.uint8).reshape(-1, bytes_per_block).copy()
    for o in FIELDS_F16[ty]:
        fa = (a[:, o].astype(np.uint16) | (a[:, o + 1].astype(np.uint16) << 8))
        fb = (b[:, o].astype(np.uint16) | (b[:, o + 1].astype(np.uint16) << 8))
        both_nan = ((fa & 0x7FFF) > 0x7C00) & ((fb & 0x7FFF) > 0x7C00)
        a[both_nan, o] = b[both_nan, o]
        a[both_nan, o + 1] = b[both_nan, o + 1]
    return bool(np.array_equal(a, b))


def nan_rule_usage(got, want, ty, blocks, bytes_per_block):
    """How much of a dequantize comparison leans on the NaN ~ NaN rule of same_floats().
    Returns (n_relaxed, n_total, ok): `n_relaxed` elements are NaN on both sides with different payloads; `ok` is False
    if any of them sits in a block with FEWER THAN TWO non-finite f16 header fields.  With one NaN / infinite field the
    reference's result is fully determined (x86: the NaN operand quieted, or 0xFFC00000 for inf * 0) and the kernels
    reproduce it bit for bit; only when two such fields meet in one expression (delta and min both NaN, ...) does the
    payload depend on the operand order the reference's compiler happened to pick."""
    got = np.asarray(got); want = np.asarray(want)
    if got.dtype == np.float32:
        ua, ub = got.view(np.uint32), want.view(np.uint32)
        na, nb = (ua & 0x7FFFFFFF) > 0x7F800000, (ub & 0x7FFFFFFF) > 0x7F800000
    else:
        ua, ub = got.view(np.uint16), want.view(np.uint16)
        na, nb = (ua & 0x7FFF) > 0x7C00, (ub & 0x7FFF) > 0x7C00
    relaxed = (ua != ub) & na & nb
    blk = np.asarray(blocks, np.uint8).reshape(-1, bytes_per_block)
    nonfinite = np.zeros(len(blk), np.int32)
    for o in FIELDS_F16[ty]:
        h = blk[:, o].astype(np.uint16) | (blk[:, o + 1].astype(np.uint16) << 8)
        nonfinite += (h & 0x7C00) == 0x7C00
    per_block = relaxed.reshape(len(blk), -1).any(axis=1)
    return int(relaxed.sum()), int(relaxed.size), not bool((per_block & (nonfinite < 2)).any())
