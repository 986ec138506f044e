"""`xtask convert` for `cast:` / `merge-linear` / `split-linear` / `permute-qk` steps on the GPUs (host mirror of include/ggq.h: ggq_convert_gguf).

    python -m gguf_b200.convert IN.gguf [IN2.gguf ...] -o OUT.gguf -x "cast:linear:q8_0 embd:q8_0 -> cast:linear:f32 embd:f32" [-t N] [-s 4G]

Reference: xtask/src/convert.rs:9-58 (`ConvertArgs{file, --steps/-x, ...}`), operator/cast.rs:28-138.
"""
import argparse
import ctypes
import json
import sys

from ._lib import ConvertStats, lib
from .quants import GgqError, QuantizeError


def parse_mem_size(s):
    """`MemSize::from_str` (xtask/src/utils/output.rs:74-92): "4G" / "512M" / "64K" / plain bytes."""
    if s is None:
        return 0
    if isinstance(s, int):
        return s
    s = s.strip()
    for suffix, shift in (("G", 30), ("M", 20), ("K", 10)):
        if s.endswith(suffix):
            return int(s[:-1]) << shift
    return int(s)


def convert(in_path, out_path, steps, n_devices=0, max_tensors=None, max_bytes=None, no_tensor_first=False, no_data=False, direct_io=False):
    """`xtask convert` for `cast:` steps.  `in_path` may be one path or a list of input shards (merged like
    `Content::new`).  Returns a dict of ggq_convert_stats.  Raises GgqError / QuantizeError like the slice API."""
    from ._lib import ConvertOptions
    paths = [in_path] if isinstance(in_path, (str, bytes)) or hasattr(in_path, "__fspath__") else list(in_path)
    arr = (ctypes.c_char_p * len(paths))(*[str(p).encode() for p in paths])
    opts = ConvertOptions(int(n_devices), int(max_tensors or 0), parse_mem_size(max_bytes), int(bool(no_tensor_first)), int(bool(no_data)), int(bool(direct_io)))
    st = ConvertStats()
    rc = lib().ggq_convert_gguf_ex(arr, len(paths), str(out_path).encode(), steps.encode(), ctypes.byref(opts), ctypes.byref(st))
    if rc == 1:
        raise QuantizeError("Indivisible")
    if rc != 0:
        msg = lib().ggq_convert_last_error().decode() or lib().ggq_last_error().decode()
        raise GgqError(rc, msg)
    return {k: getattr(st, k) for k, _ in ConvertStats._fields_}


def main(argv=None):
    ap = argparse.ArgumentParser(description="GGUF cast conversion on B200 GPUs (xtask convert, cast: steps)")
    ap.add_argument("file", nargs="+", help="input GGUF file(s); several shards are merged")
    ap.add_argument("-o", "--out", required=True, help="output path (shards get -0000i-of-0000N before .gguf)")
    ap.add_argument("-x", "--steps", required=True, help='e.g. "cast:linear:q8_0 embd:f16 -> cast:linear:f32"')
    ap.add_argument("-t", "--max-tensors", type=int, default=None)
    ap.add_argument("-s", "--max-bytes", default=None, help="e.g. 4G, 512M")
    ap.add_argument("--no-tensor-first", action="store_true")
    ap.add_argument("--no-data", action="store_true")
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--direct-io", action="store_true", help="O_DIRECT reads into the pinned staging buffers (inputs larger than RAM)")
    a = ap.parse_args(argv)
    print(json.dumps(convert(a.file, a.out, a.steps, a.gpus, a.max_tensors, a.max_bytes, a.no_tensor_first, a.no_data, a.direct_io)))


if __name__ == "__main__":
    sys.exit(main())
