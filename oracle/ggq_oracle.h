/*
 * ggq_oracle.h — CPU ORACLE for the ggml-quants block codec.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's per-block arithmetic, used only as the
 * checker in tests/, __graft_entry__.smoke() and as the CPU baseline leg of bench.py.  The
 * product path (gguf_b200/, libggq.so) never links, imports or calls anything in oracle/.
 *
 * Parity status
 *   - legacy blocks (Q4_0 Q4_1 Q5_0 Q5_1 Q8_0 Q8_1), Q8K, f16/bf16: restated op-for-op from
 *     /root/reference/ggml-quants/src/structs/{q4_0,...}.rs (citations at each function).  The reference
 *     holds NO golden vectors (only tolerance round trips), and its Rust cannot be compiled in
 *     this image (no cargo/rustc), so the restatement is pinned by (i) hand-derived known-answer
 *     blocks for the README example, (ii) byte equality with the independent
 *     gguf-py 0.19.0 numpy quantizers on every non-all-zero block, (iii) the reference's own
 *     tolerances.  => "parity pinned by restatement + independent cross-check".
 *   - K-quants (Q2K..Q6K): the reference has only struct layouts; quantize/dequantize are
 *     `todo!()` (e.g. ggml-quants/src/structs/q4_k.rs:24-31).  Arithmetic here follows upstream
 *     ggml `quantize_row_qN_K_ref` / `dequantize_row_qN_K` semantics restated from the published
 *     algorithm.  Dequant is cross-checked bit-for-bit against gguf-py; quantize has nothing to
 *     be pinned against => **K-quant quantize: parity unpinned** (the oracle *defines* it).
 */
#ifndef GGQ_ORACLE_H
#define GGQ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* GGmlType discriminants: /root/reference/ggus/src/tensor.rs:15-50 */
enum {
    GGO_F32 = 0, GGO_F16 = 1, GGO_Q4_0 = 2, GGO_Q4_1 = 3, GGO_Q5_0 = 6, GGO_Q5_1 = 7,
    GGO_Q8_0 = 8, GGO_Q8_1 = 9, GGO_Q2K = 10, GGO_Q3K = 11, GGO_Q4K = 12, GGO_Q5K = 13,
    GGO_Q6K = 14, GGO_Q8K = 15, GGO_BF16 = 30
};

/* QuantizeError: /root/reference/ggml-quants/src/lib.rs:107-113 */
enum { GGO_OK = 0, GGO_INDIVISIBLE = 1, GGO_LENGTH_MISMATCH = 2, GGO_UNSUPPORTED = -1 };

/* block geometry; returns 0 or GGO_UNSUPPORTED */
int ggo_block_info(uint32_t type, uint32_t *elems, uint32_t *bytes);

/* QuantExt::quantize_slice(dst, src) — lib.rs:121-133.  `fdt` is the float-side dtype
 * (GGO_F32 / GGO_F16 / GGO_BF16).  `threads` <= 1 runs inline; otherwise contiguous block
 * ranges are split over that many pthreads (the rayon driver's shape). */
int ggo_quantize_slice(uint32_t type, uint32_t fdt, void *dst, size_t dst_blocks,
                       const void *src, size_t src_elems, int threads);

/* QuantExt::dequantize_slice(dst, src) — lib.rs:135-147 */
int ggo_dequantize_slice(uint32_t type, uint32_t fdt, void *dst, size_t dst_elems,
                         const void *src, size_t src_blocks, int threads);

/* scalar conversions (half 2.6.0 semantics), exposed for tests */
uint16_t ggo_f32_to_f16(float v);
float ggo_f16_to_f32(uint16_t h);
uint16_t ggo_f32_to_bf16(float v);
float ggo_bf16_to_f32(uint16_t h);

#ifdef __cplusplus
}
#endif
#endif
