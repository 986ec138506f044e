"""gguf_b200 — B200-native (sm_100a) ggml block quantize / dequantize behind the reference's
`ggml-quants` slice API.  See include/ggq.h for the C ABI and DESIGN.md for the design."""
from .quants import (BF16, BLOCK_TYPES, F16, F32, FLOAT_TYPES, Q2K, Q3K, Q4_0, Q4_1, Q4K, Q5_0, Q5_1, Q5K, Q6K, Q8_0,
                     Q8_1, Q8K, TYPE_NAMES, GgqError, PinnedBuffer, QuantizeError, block_info, dequantize,
                     dequantize_slice, dequantize_slice_device, quantize, quantize_slice, quantize_slice_device, slices, slices_device)

from .rearrange import ArrayLayout, concat, permute_qk, rearrange, rearrange_device, split

__all__ = [n for n in dir() if not n.startswith("_")]
