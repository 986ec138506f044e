"""CPU tests of the C-ABI boundary: the library loads, exports every declared symbol, validates
lengths in the reference's order, and refuses to compute without a GPU (no CPU fallback)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ggq.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ggq_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_are_exported(ggq):
    from gguf_b200._lib import SO_PATH, SYMBOLS, lib
    L = lib()
    decl = declared_symbols()
    assert len(decl) >= 12
    assert sorted(s[0] for s in SYMBOLS) == decl
    nm = subprocess.run(["nm", "-D", "--defined-only", SO_PATH], capture_output=True, text=True).stdout
    for name in decl:
        assert hasattr(L, name)
        assert re.search(rf"\bT {name}\b", nm), f"{name} not exported with C linkage"


def test_block_info_matches_reference_layouts(ggq):
    want = {1: (1, 2), 30: (1, 2), 2: (32, 18), 3: (32, 20), 6: (32, 22), 7: (32, 24), 8: (32, 34), 9: (32, 36),
            10: (256, 84), 11: (256, 110), 12: (256, 144), 13: (256, 176), 14: (256, 210), 15: (256, 290)}
    for ty, eb in want.items():
        assert ggq.block_info(ty) == eb
    with pytest.raises(ggq.GgqError):
        ggq.block_info(16)  # IQ2XXS: out of scope


def test_error_order_matches_reference(ggq):
    """lib.rs:293-331: divisibility is checked before length equality; both before touching a device."""
    src31, src64 = np.zeros(31, np.float32), np.zeros(64, np.float32)
    one, three = np.zeros(34, np.uint8), np.zeros(3 * 34, np.uint8)
    with pytest.raises(ggq.QuantizeError) as e:
        ggq.quantize_slice(ggq.Q8_0, one, src31)
    assert e.value.kind == "Indivisible"
    with pytest.raises(ggq.QuantizeError) as e:
        ggq.quantize_slice(ggq.Q8_0, three, src64)
    assert e.value.kind == "LengthMismatch"
    with pytest.raises(ggq.QuantizeError) as e:
        ggq.dequantize_slice(ggq.Q8_0, src31, one)
    assert e.value.kind == "Indivisible"
    with pytest.raises(ggq.QuantizeError) as e:
        ggq.dequantize_slice(ggq.Q8_0, src64, three)
    assert e.value.kind == "LengthMismatch"
    # 31 elements into 3 blocks: Indivisible wins
    with pytest.raises(ggq.QuantizeError) as e:
        ggq.quantize_slice(ggq.Q8_0, three, src31)
    assert e.value.kind == "Indivisible"
    # f16 / bf16 float sides and a K-quant
    with pytest.raises(ggq.QuantizeError) as e:
        ggq.quantize_slice(ggq.Q4K, np.zeros(144, np.uint8), np.zeros(255, np.uint16), fdt=ggq.BF16)
    assert e.value.kind == "Indivisible"


def test_empty_slices_are_ok_without_a_device(ggq):
    ggq.quantize_slice(ggq.Q4_0, np.zeros(0, np.uint8), np.zeros(0, np.float32))
    ggq.dequantize_slice(ggq.Q6K, np.zeros(0, np.float32), np.zeros(0, np.uint8))


def test_no_cpu_fallback(ggq):
    """Without a GPU a compute call must fail loudly, never silently compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ggq.GgqError) as e:
        ggq.quantize(ggq.Q8_0, np.ones(64, np.float32))
    assert e.value.code == -2


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gguf_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+\"[^\"]*oracle|ggo_[a-z]", txt, flags=re.M), f


def test_slices_validates_every_job_before_computing(ggq):
    """ggq_slices: the first failing job's QuantizeError is returned and nothing runs (works without a GPU)."""
    ok = ("quantize", ggq.Q8_0, np.zeros(2 * 34, np.uint8), np.zeros(64, np.float32))
    bad_len = ("dequantize", ggq.Q8_0, np.zeros(64, np.float32), np.zeros(3 * 34, np.uint8))
    bad_div = ("quantize", ggq.Q4K, np.zeros(144, np.uint8), np.zeros(255, np.float32))
    with pytest.raises(ggq.QuantizeError) as e:
        ggq.slices([ok, bad_len, bad_div])
    assert e.value.kind == "LengthMismatch"
    with pytest.raises(ggq.QuantizeError) as e:
        ggq.slices([bad_div, bad_len])
    assert e.value.kind == "Indivisible"
    ggq.slices([])
