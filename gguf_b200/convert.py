"""`xtask convert` for `cast:` steps on the GPUs (host mirror of include/ggq.h: ggq_convert_gguf).

    python -m gguf_b200.convert IN.gguf OUT.gguf -x "cast:linear:q8_0 embd:q8_0 -> cast:linear:f32 embd:f32"

Reference: xtask/src/convert.rs:9-58 (`ConvertArgs{file, --steps/-x, ...}`), operator/cast.rs:28-138.
"""
import argparse
import ctypes
import json
import sys

from ._lib import ConvertStats, lib
from .quants import GgqError, QuantizeError


def convert(in_path, out_path, steps, n_devices=0):
    """Returns a dict of ggq_convert_stats. Raises GgqError / QuantizeError like the slice API."""
    st = ConvertStats()
    rc = lib().ggq_convert_gguf(str(in_path).encode(), str(out_path).encode(), steps.encode(), int(n_devices), ctypes.byref(st))
    if rc == 1:
        raise QuantizeError("Indivisible")
    if rc != 0:
        msg = lib().ggq_convert_last_error().decode() or lib().ggq_last_error().decode()
        raise GgqError(rc, msg)
    return {k: getattr(st, k) for k, _ in ConvertStats._fields_}


def main(argv=None):
    ap = argparse.ArgumentParser(description="GGUF cast conversion on B200 GPUs")
    ap.add_argument("file")
    ap.add_argument("out")
    ap.add_argument("-x", "--steps", required=True, help='e.g. "cast:linear:q8_0 embd:f16 -> cast:linear:f32"')
    ap.add_argument("--gpus", type=int, default=0)
    a = ap.parse_args(argv)
    print(json.dumps(convert(a.file, a.out, a.steps, a.gpus)))


if __name__ == "__main__":
    sys.exit(main())
