"""CPU tests: pin the oracle (oracle/ggq_oracle.c) against vectors that do not come from it."""
import os

import numpy as np
import pytest

from data import edge_blocks, gaussian, to_fdt

G = os.path.join(os.path.dirname(__file__), "golden")
LEGACY = {"q4_0": 2, "q4_1": 3, "q5_0": 6, "q5_1": 7, "q8_0": 8}


def test_readme_block_known_answers(oracle):
    """ggml-quants/README.md:32-37 example block -> exact packed bytes."""
    k = np.load(os.path.join(G, "readme_block.npz"))
    for name, ty in [("q8_0", 8), ("q4_0", 2), ("q4_1", 3), ("q5_0", 6), ("q8_1", 9)]:
        assert oracle.quantize(ty, oracle.F32, k["x"]).tobytes() == k[name].tobytes(), name


@pytest.mark.parametrize("name", sorted(LEGACY))
def test_legacy_matches_gguf_py_golden(oracle, name):
    g = np.load(os.path.join(G, "gguf_py_legacy.npz"))
    ty = LEGACY[name]
    q = oracle.quantize(ty, oracle.F32, g["x"])
    assert np.array_equal(q, g[name])
    d = oracle.dequantize(ty, oracle.F32, g[name])
    assert np.array_equal(d.view(np.uint32), g[name + "_deq"].view(np.uint32))


@pytest.mark.parametrize("name,ty", [("q2k", 10), ("q3k", 11), ("q4k", 12), ("q5k", 13), ("q6k", 14)])
def test_kquant_dequant_matches_gguf_py_golden(oracle, name, ty):
    g = np.load(os.path.join(G, "gguf_py_kdequant.npz"))
    d = oracle.dequantize(ty, oracle.F32, g[name])
    assert np.array_equal(d.view(np.uint32), g[name + "_deq"].view(np.uint32))


def test_live_gguf_py_crosscheck(oracle):
    """Same cross-check on fresh seeds when gguf-py is importable (it is in this image)."""
    gq = pytest.importorskip("gguf.quants")
    from gguf import GGMLQuantizationType as T
    x = np.concatenate([gaussian(32 * 500, 11, s) for s in (1e-3, 0.02, 5.0)])
    for ty, gt in [(2, T.Q4_0), (3, T.Q4_1), (6, T.Q5_0), (7, T.Q5_1), (8, T.Q8_0)]:
        assert np.array_equal(oracle.quantize(ty, 0, x), gq.quantize(x.reshape(-1, 32), gt).reshape(-1))


def test_zero_block_is_all_zero_bytes(oracle):
    """the reference returns Self::ZEROS (diverges from ggml)."""
    for ty in (2, 3, 6, 7, 8, 9, 15):
        n, b = oracle.block_info(ty)
        assert not oracle.quantize(ty, 0, np.zeros(n, np.float32)).any()
    # dequantising a zero Q4_0 block gives -0.0 (sign bit set): (0 - 8) * +0.0
    y = oracle.dequantize(2, 0, np.zeros(18, np.uint8))
    assert (y.view(np.uint32) == 0x80000000).all()


def test_error_order(oracle):
    """lib.rs:293-331 known answers."""
    assert oracle.quantize_rc(8, 0, 1, 31) == oracle.INDIVISIBLE
    assert oracle.quantize_rc(8, 0, 3, 64) == oracle.LENGTH_MISMATCH
    assert oracle.dequantize_rc(8, 0, 31, 1) == oracle.INDIVISIBLE
    assert oracle.dequantize_rc(8, 0, 64, 3) == oracle.LENGTH_MISMATCH
    assert oracle.quantize_rc(8, 0, 2, 64) == oracle.OK


@pytest.mark.parametrize("ty,tol", [(1, 4e-3), (30, 8e-3), (2, 8e-2), (3, 4e-2), (6, 4e-2), (7, 2e-2), (8, 4.5e-3), (9, 4.5e-3), (15, 4.5e-3)])
def test_reference_tolerances(oracle, ty, tol):
    """the reference's own round-trip thresholds on [0,1) inputs."""
    n, _ = oracle.block_info(ty)
    x = np.random.default_rng(ty).random(n * 64, dtype=np.float32)
    y = oracle.dequantize(ty, 0, oracle.quantize(ty, 0, x))
    assert np.abs(x - y).max() <= tol


@pytest.mark.parametrize("ty,rel", [(10, 0.35), (11, 0.2), (12, 0.09), (13, 0.05), (14, 0.025)])
def test_kquant_roundtrip_quality(oracle, ty, rel):
    """K-quant quantize is unpinned by the reference; sanity: rmse/sigma in the range upstream ggml shows."""
    x = gaussian(256 * 200, 5)
    y = oracle.dequantize(ty, 0, oracle.quantize(ty, 0, x))
    assert np.sqrt(np.mean((x - y) ** 2)) / 0.02 < rel


def test_half_conversions_exhaustive(oracle):
    """f16 -> f32 -> f16 is the identity on every non-NaN pattern and matches numpy on a float sweep."""
    L = oracle.lib()
    allh = np.arange(65536, dtype=np.uint16)
    wide = oracle.dequantize(1, 0, allh.view(np.uint8))
    back = oracle.quantize(1, 0, wide).view(np.uint16)
    nan = (allh & 0x7FFF) > 0x7C00
    assert np.array_equal(back[~nan], allh[~nan])
    assert np.array_equal(wide[~nan].view(np.uint32), allh[~nan].view(np.float16).astype(np.float32).view(np.uint32))
    assert np.array_equal(back[nan], allh[nan] | 0x0200)  # quieted, payload kept
    x = (np.random.default_rng(1).standard_normal(200000) * 10.0 ** np.random.default_rng(2).integers(-9, 6, 200000)).astype(np.float32)
    with np.errstate(over="ignore"):
        assert np.array_equal(oracle.quantize(1, 0, x).view(np.uint16), x.astype(np.float16).view(np.uint16))
    assert L.ggo_f32_to_bf16(1.0) == 0x3F80 and L.ggo_f32_to_bf16(np.float32(1.00390625)) == 0x3F80  # tie -> even


def test_threaded_driver_is_deterministic(oracle):
    x = to_fdt(gaussian(32 * 4096, 3), 1)
    assert np.array_equal(oracle.quantize(8, 1, x, threads=1), oracle.quantize(8, 1, x, threads=7))


def test_edge_blocks_do_not_crash(oracle):
    for ty in (2, 3, 6, 7, 8, 9, 15, 10, 11, 12, 13, 14):
        n, _ = oracle.block_info(ty)
        x = edge_blocks(n)
        q = oracle.quantize(ty, 0, x)
        oracle.dequantize(ty, 0, q)


def test_f16c_mediation_equals_software(oracle):
    """The slice drivers widen/narrow f16 with F16C when the CPU has it; results must equal the
    software `half` restatement on every f16 pattern (NaNs included)."""
    allh = np.arange(65536, dtype=np.uint16)
    pad = np.zeros(32 - 65536 % 32 if 65536 % 32 else 0, np.uint16)
    h = np.concatenate([allh, pad])
    wide = oracle.dequantize(1, 0, h.view(np.uint8))           # software widen (1-element blocks)
    for ty in (8, 2, 3):                                        # Q8_0, Q4_0, Q4_1 quantize from f16 vs from the widened f32
        assert np.array_equal(oracle.quantize(ty, 1, h), oracle.quantize(ty, 0, wide)), ty
    blocks = np.random.default_rng(3).integers(0, 256, 34 * 4096, dtype=np.uint8)
    y32 = oracle.dequantize(8, 0, blocks)
    assert np.array_equal(oracle.dequantize(8, 1, blocks), oracle.quantize(1, 0, y32).view(np.uint16))  # hw narrow == software narrow


@pytest.mark.parametrize("ty", [10, 11, 12, 13, 14])
def test_kquant_quantize_matches_independent_numpy_restatement(oracle, ty):
    """K-quant quantize has no reference arithmetic to pin it (todo!() in structs/q{2..6}_k.rs): the C oracle must at
    least agree, byte for byte, with a second restatement of upstream's published algorithm written independently in
    numpy (tests/kquant_numpy.py) — Gaussian, heavy-tailed, mixed-magnitude, constant, all-zero and tiny super-blocks."""
    from kquant_numpy import QUANTIZERS
    rng = np.random.default_rng(1234 + ty)
    parts = [
        (rng.standard_normal(256 * 96) * 0.02).astype(np.float32),
        (rng.standard_t(3, 256 * 48) * 0.02).astype(np.float32),
        (rng.standard_normal(256 * 24) * np.repeat(10.0 ** rng.integers(-6, 3, 24 * 16), 16)).astype(np.float32),
        rng.random(256 * 8, dtype=np.float32),                       # all positive: min clamps to 0
        np.full(256 * 2, 0.37, np.float32),                          # max == min
        np.zeros(256 * 2, np.float32),
        (rng.standard_normal(256 * 4) * 1e-12).astype(np.float32),
        (rng.standard_normal(256 * 8) * 0.02).astype(np.float16).astype(np.float32),   # f16-representable inputs
    ]
    x = np.concatenate(parts)
    got = oracle.quantize(ty, oracle.F32, x)
    want = QUANTIZERS[ty](x)
    _, b = oracle.block_info(ty)
    bad = np.flatnonzero((got.reshape(-1, b) != want.reshape(-1, b)).any(axis=1))
    assert bad.size == 0, f"{bad.size} of {x.size // 256} super-blocks differ, first at {bad[:5]}"


# ---- known answers derived from the Rust text by a third, scalar numpy restatement (tests/golden/rust_restatement.py) ----
KAT_TYPES = [2, 3, 6, 7, 8, 9, 15]
KAT_FDT = [(0, "f32"), (1, "f16"), (30, "bf16")]


def test_rust_restatement_reproduces_the_hand_derived_readme_block():
    """The restatement the KATs come from is itself pinned: README block (App. C.1) and gguf-py on non-zero blocks."""
    import sys
    sys.path.insert(0, G)
    import rust_restatement as R
    k = np.load(os.path.join(G, "readme_block.npz"))
    for name, ty in [("q8_0", 8), ("q4_0", 2), ("q4_1", 3), ("q5_0", 6), ("q8_1", 9)]:
        assert R.quantize(ty, k["x"]).tobytes() == k[name].tobytes(), name
    g = np.load(os.path.join(G, "gguf_py_legacy.npz"))
    x = g["x"][:32 * 96]
    for name, ty in LEGACY.items():
        b = {2: 18, 3: 20, 6: 22, 7: 24, 8: 34}[ty]
        assert np.array_equal(R.quantize(ty, x), g[name][:96 * b]), name
        assert np.array_equal(R.dequantize(ty, g[name][:96 * b]).view(np.uint32), g[name + "_deq"][:32 * 96].view(np.uint32)), name


@pytest.mark.parametrize("ty", KAT_TYPES)
def test_oracle_matches_rust_kat(oracle, ty):
    """Q8K, Q5_1, Q8_1 (sum from the unrounded delta), the f16 / bf16 float sides, NaN / inf / tie / zero rows:
    quantize bytes and dequantize bits equal the known answers of tests/golden/rust_kat.npz."""
    k = np.load(os.path.join(G, "rust_kat.npz"))
    for fdt, name in KAT_FDT:
        x = k[f"x_{ty}"] if fdt == 0 else k[f"x_{ty}_{name}"]
        assert np.array_equal(oracle.quantize(ty, fdt, x), k[f"q_{ty}_{name}"]), (ty, name)
        want = k[f"d_{ty}_{name}"]
        got = oracle.dequantize(ty, fdt, k[f"b_{ty}"])
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), (ty, name)


def test_oracle_matches_rust_reference_dump_when_present(oracle):
    """`cargo run --example dump_golden` (rust/ggml-quants-cuda) writes the REFERENCE's own outputs to
    tests/golden/rust_reference.bin; when that file is there, every record must be reproduced by the oracle."""
    import struct
    path = os.path.join(G, "rust_reference.bin")
    if not os.path.exists(path):
        pytest.skip("no Rust toolchain here: tests/golden/rust_reference.bin has not been generated")
    from data import same_blocks, same_floats
    buf = open(path, "rb").read()
    assert buf[:4] == b"GGQR" and struct.unpack_from("<I", buf, 4)[0] == 1
    p, n_rec = 8, 0
    while p < len(buf):
        kind, ty, fdt, n, nin, nout = struct.unpack_from("<IIIQQQ", buf, p)
        p += 36
        din, dout = np.frombuffer(buf, np.uint8, nin, p), np.frombuffer(buf, np.uint8, nout, p + nin)
        p += nin + nout
        fl = np.float32 if fdt == 0 else np.uint16
        if kind == 0:
            got = oracle.quantize(ty, fdt, din.view(fl))
            _, b = oracle.block_info(ty)
            assert same_blocks(got, dout, ty, b) if ty not in (1, 30) else np.array_equal(got, dout), (kind, ty, fdt)
        else:
            got = oracle.dequantize(ty, fdt, din)
            assert same_floats(got, dout.view(fl)), (kind, ty, fdt)
        n_rec += 1
    assert n_rec >= 60


def test_rust_ffi_declares_every_header_symbol():
    """rust/ggml-quants-cuda-sys/src/lib.rs is source-only here (no cargo); at least its extern block must name
    exactly the functions include/ggq.h declares."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "ggq.h")).read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(ggq_[a-z_0-9]+)\s*\(", hdr)))
    rs = open(os.path.join(root, "rust", "ggml-quants-cuda-sys", "src", "lib.rs")).read()
    assert sorted(set(re.findall(r"pub fn (ggq_[a-z_0-9]+)\s*\(", rs))) == declared
