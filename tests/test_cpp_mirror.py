"""Builds and runs tests/cpp/test_quant_ext.cpp (the reference's unit tests against the C++ mirror)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_quant_ext")


def build(ggq):
    so_dir = os.path.join(ROOT, "gguf_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", EXE, os.path.join(ROOT, "tests", "cpp", "test_quant_ext.cpp"),
                           "-L" + so_dir, "-l:libggq.so", "-Wl,-rpath," + so_dir])


def test_cpp_mirror_error_order_no_gpu(ggq):
    build(ggq)
    r = subprocess.run([EXE, "--no-gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_tensor_ops_planner_no_gpu(ggq):
    """tests/cpp/test_tensor_ops.cpp: merge-linear / split-linear / permute-qk / cast planning (host/tensor_ops.hpp)."""
    exe = os.path.join(ROOT, "tests", "cpp", "test_tensor_ops")
    so_dir = os.path.join(ROOT, "gguf_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_tensor_ops.cpp"),
                           "-L" + so_dir, "-l:libggq.so", "-Wl,-rpath," + so_dir])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_shared_divisor_division_is_correctly_rounded():
    """tests/cpp/test_shared_divisor.c: the three-operation division of quant_k.cu (div_shared) returns the bits of n / d."""
    exe = os.path.join(ROOT, "tests", "cpp", "test_shared_divisor")
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_shared_divisor.c"), "-lm"])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_relu_clamp_equals_the_two_sided_float_clamp():
    """tests/cpp/test_clamp_relu.c: `min.relu.s32` on the float's bits (the K-quant search's one-instruction clamp to
    [0, nmax]) returns the value of fminf(fmaxf(v, 0), nmax) for every non-NaN float (every 13th pattern and the
    neighbourhoods of the boundaries here; `test_clamp_relu all` walks all 2^32)."""
    exe = os.path.join(ROOT, "tests", "cpp", "test_clamp_relu")
    subprocess.check_call(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_clamp_relu.c"), "-lm"])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 differ" in r.stdout


def test_q6k_clamp_bounds_hold_for_every_candidate_group():
    """tests/cpp/test_q6k_clamp_bounds.c: the Q6K search applies its clamp to [-32, 31] only where a candidate can reach a
    bound (quant_k_kernel.cuh, make_qx_quants16, CS 1); the float expressions of the kernel, evaluated on the host for the
    extreme ratios x = +-max over every exponent and adversarial significands, stay inside the claimed ranges."""
    exe = os.path.join(ROOT, "tests", "cpp", "test_q6k_clamp_bounds")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_q6k_clamp_bounds.c"), "-lm"])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 bound violations" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_reference_unit_tests(ggq):
    build(ggq)
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
