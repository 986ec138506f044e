"""A per-block, scalar restatement of the reference's legacy quantizers in numpy float32 — written from the Rust
text (ggml-quants/src/structs/q4_0.rs ... q8_k.rs, structs.rs:54-107, lib.rs:62-90), NOT from oracle/ggq_oracle.c.
It exists only to derive known-answer vectors (tests/golden/make_rust_kat.py -> rust_kat.npz) for the types and float
sides that had none: Q8K, Q5_1, Q8_1, and the f16 / bf16 adapters.  Every arithmetic step is one np.float32 operation
(one IEEE rounding each, like the Rust), casts are Rust's saturating `as`, `.round()` is half-away-from-zero.

Domain: finite or NaN/inf elements are fine; blocks whose minimum is a signed-zero tie are avoided by the KAT inputs
(the sign of `min` is then decided by rustc's lowering of f32::min, see oracle/ggq_oracle.c)."""
import numpy as np

f32 = np.float32


def as_u8(v):      # Rust `f32 as u8`: NaN -> 0, saturating, truncating
    if np.isnan(v):
        return 0
    return int(min(max(np.trunc(v), 0.0), 255.0))


def as_i8(v):      # Rust `f32 as i8`
    if np.isnan(v):
        return 0
    return int(min(max(np.trunc(v), -128.0), 127.0))


def rust_round(v):  # f32::round: half away from zero (exact for |v| < 2^23; NaN / inf pass through)
    if not np.isfinite(v):
        return v
    t = np.trunc(v)
    if abs(f32(v - t)) >= f32(0.5):
        t = f32(t + np.copysign(f32(1.0), v))
    return f32(t)


def rust_min(a, b):  # f32::min / f32::max: the non-NaN operand wins
    return f32(np.fmin(a, b))


def rust_max(a, b):
    return f32(np.fmax(a, b))


def f16_bits(v):   # half::f16::from_f32 for non-NaN inputs: RNE, overflow -> inf
    with np.errstate(over="ignore"):
        return int(np.array([v], np.float32).astype(np.float16).view(np.uint16)[0])


def f16_to_f32(bits):
    return f32(np.array([bits], np.uint16).view(np.float16).astype(np.float32)[0])


def bf16_bits(v):  # half::bf16::from_f32, non-NaN: RNE on the upper 16 bits
    u = int(np.array([v], np.float32).view(np.uint32)[0])
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) & 0xFFFF


def bf16_to_f32(bits):
    return f32(np.array([bits << 16], np.uint32).view(np.float32)[0])


def max_abs(x):      # structs.rs:91-94
    acc = f32(0.0)
    for v in x:
        acc = rust_max(acc, f32(abs(v)))
    return acc


def max_by_abs(x):   # structs.rs:96-100
    acc = f32(0.0)
    for v in x:
        if abs(v) > abs(acc):
            acc = f32(v)
    return acc


def min_max(x):      # structs.rs:102-107
    lo, hi = f32(np.finfo(np.float32).max), f32(np.finfo(np.float32).min)
    for v in x:
        lo, hi = rust_min(lo, v), rust_max(hi, v)
    return lo, hi


def le16(b):
    return [b & 0xFF, b >> 8]


def q4_0(x):   # q4_0.rs:23-44
    mx = max_by_abs(x)
    if mx == 0.0:
        return [0] * 18
    with np.errstate(all="ignore"):
        delta = f32(mx / f32(-8.0))
        recip = f32(f32(1.0) / delta)
        f = lambda v: as_u8(rust_min(f32(f32(v * recip) + f32(8.5)), f32(15.0)))
        return le16(f16_bits(delta)) + [(f(x[i + 16]) << 4) | f(x[i]) for i in range(16)]


def q4_1(x):   # q4_1.rs:23-47
    mn, mx = min_max(x)
    if mn == mx:
        return le16(0) + le16(f16_bits(mn)) + [0] * 16
    with np.errstate(all="ignore"):
        delta = f32(f32(mx - mn) / f32(15.0))
        recip = f32(f32(1.0) / delta)
        f = lambda v: min(as_u8(f32(f32(f32(v - mn) * recip) + f32(0.5))), 15)
        return le16(f16_bits(delta)) + le16(f16_bits(mn)) + [(f(x[i + 16]) << 4) | f(x[i]) for i in range(16)]


def _pack5(codes):  # q5_0.rs:44-57, q5_1.rs:47-60
    qh, ql = 0, []
    for i in range(16):
        l, h = codes[i], codes[i + 16]
        qh |= ((l >> 4) & 1) << i
        qh |= ((h >> 4) & 1) << (i + 16)
        ql.append(((h & 0xF) << 4) | (l & 0xF))
    return [(qh >> s) & 0xFF for s in (0, 8, 16, 24)], ql


def q5_0(x):   # q5_0.rs:26-58
    mx = max_by_abs(x)
    if mx == 0.0:
        return [0] * 22
    with np.errstate(all="ignore"):
        delta = f32(mx / f32(-16.0))
        recip = f32(f32(1.0) / delta)
        f = lambda v: min(as_u8(f32(f32(v * recip) + f32(16.5))), 31)
        qh, ql = _pack5([f(v) for v in x])
        return le16(f16_bits(delta)) + qh + ql


def q5_1(x):   # q5_1.rs:26-62
    mn, mx = min_max(x)
    if mn == mx:
        return le16(0) + le16(f16_bits(mn)) + [0] * 20
    with np.errstate(all="ignore"):
        delta = f32(f32(mx - mn) / f32(31.0))
        recip = f32(f32(1.0) / delta)
        f = lambda v: min(as_u8(f32(f32(f32(v - mn) * recip) + f32(0.5))), 31)
        qh, ql = _pack5([f(v) for v in x])
        return le16(f16_bits(delta)) + le16(f16_bits(mn)) + qh + ql


def q8_0(x):   # q8_0.rs:23-41
    amax = max_abs(x)
    if amax == 0.0:
        return [0] * 34
    with np.errstate(all="ignore"):
        delta = f32(amax / f32(127.0))
        recip = f32(f32(1.0) / delta)
        return le16(f16_bits(delta)) + [as_i8(rust_round(f32(v * recip))) & 0xFF for v in x]


def q8_1(x):   # q8_1.rs:28-55: sum = f16(sum_i16 as f32 * delta) with the UNROUNDED f32 delta
    amax = max_abs(x)
    if amax == 0.0:
        return [0] * 36
    with np.errstate(all="ignore"):
        delta = f32(amax / f32(127.0))
        recip = f32(f32(1.0) / delta)
        q = [as_i8(rust_round(f32(v * recip))) for v in x]
        s = sum(q)                                                # i16 accumulation cannot overflow: |s| <= 32 * 127
        return le16(f16_bits(delta)) + le16(f16_bits(f32(f32(s) * delta))) + [v & 0xFF for v in q]


def q8_k(x):   # q8_k.rs:27-54 — the reference's 290-byte layout {delta: f16, quants: [i8; 256], sums: [i16; 16]}
    mx = max_by_abs(x)
    if mx == 0.0:
        return [0] * 290
    with np.errstate(all="ignore"):
        delta = f32(mx / f32(-127.0))
        recip = f32(f32(1.0) / delta)
        q = [as_i8(rust_min(rust_round(f32(v * recip)), f32(127.0))) for v in x]
        sums = [sum(q[16 * g:16 * g + 16]) for g in range(16)]
        out = le16(f16_bits(delta)) + [v & 0xFF for v in q]
        for s in sums:
            out += le16(s & 0xFFFF)
        return out


QUANT = {2: (32, q4_0), 3: (32, q4_1), 6: (32, q5_0), 7: (32, q5_1), 8: (32, q8_0), 9: (32, q8_1), 15: (256, q8_k)}


def quantize(ty, x32):
    """x32: float32 array (already widened from f16 / bf16 exactly, lib.rs:66-69); returns uint8 blocks."""
    n, fn = QUANT[ty]
    x32 = np.asarray(x32, np.float32)
    return np.array([b for i in range(0, x32.size, n) for b in fn(x32[i:i + n])], np.uint8)


# ---- dequantize (f32 results; the f16 / bf16 adapters narrow them, lib.rs:70-73, 86-89) ----
def _s8(b):
    return b - 256 if b > 127 else b


def dq4_0(b):   # q4_0.rs:46-57
    d = f16_to_f32(b[0] | b[1] << 8)
    lo = [f32(f32((q & 0xF) - 8) * d) for q in b[2:18]]
    hi = [f32(f32((q >> 4) - 8) * d) for q in b[2:18]]
    return lo + hi


def dq4_1(b):   # q4_1.rs:49-60
    d, m = f16_to_f32(b[0] | b[1] << 8), f16_to_f32(b[2] | b[3] << 8)
    lo = [f32(f32(f32(q & 0xF) * d) + m) for q in b[4:20]]
    hi = [f32(f32(f32(q >> 4) * d) + m) for q in b[4:20]]
    return lo + hi


def _codes5(qh_bytes, ql):   # q5_0.rs:60-73: l | ((qh >> i) << 4 & 0x10) ; h | (qh >> (i + 12) & 0x10)
    qh = qh_bytes[0] | qh_bytes[1] << 8 | qh_bytes[2] << 16 | qh_bytes[3] << 24
    lo = [(ql[i] & 0xF) | (((qh >> i) << 4) & 0x10) for i in range(16)]
    hi = [(ql[i] >> 4) | ((qh >> (i + 12)) & 0x10) for i in range(16)]
    return lo + hi


def dq5_0(b):
    d = f16_to_f32(b[0] | b[1] << 8)
    return [f32(f32(c - 16) * d) for c in _codes5(b[2:6], b[6:22])]


def dq5_1(b):   # q5_1.rs:64-77
    d, m = f16_to_f32(b[0] | b[1] << 8), f16_to_f32(b[2] | b[3] << 8)
    return [f32(f32(f32(c) * d) + m) for c in _codes5(b[4:8], b[8:24])]


def dq8(off):
    def fn(b):      # q8_0.rs:43-47, q8_1.rs:57-61, q8_k.rs:56-60
        d = f16_to_f32(b[0] | b[1] << 8)
        n = 256 if len(b) == 290 else 32
        return [f32(f32(_s8(q)) * d) for q in b[off:off + n]]
    return fn


DEQUANT = {2: (18, dq4_0), 3: (20, dq4_1), 6: (22, dq5_0), 7: (24, dq5_1), 8: (34, dq8(2)), 9: (36, dq8(4)), 15: (290, dq8(2))}


def dequantize(ty, blocks):
    size, fn = DEQUANT[ty]
    blocks = [int(v) for v in np.asarray(blocks, np.uint8)]
    with np.errstate(all="ignore"):
        return np.array([v for i in range(0, len(blocks), size) for v in fn(blocks[i:i + size])], np.float32)
