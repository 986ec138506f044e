/*
 * ggq.h — C ABI of libggq.so: the B200-native (sm_100a) ggml block quantize / dequantize path.
 *
 * This is the drop-in boundary for the reference's `ggml-quants` slice API.  Each entry point
 * cites the reference interface it replaces (paths under /root/reference/).  Plain pointers and
 * sizes only; no torch / C++ types.  The Rust-side binding a maintainer would add is shown in
 * INTEGRATION.md.
 *
 * Semantics common to every call
 *   - `type` is a GGmlType discriminant (ggus/src/tensor.rs:15-50): the packed/block side.
 *   - `fdt` is the float-side element type T of `Quantize<T, N>` (ggml-quants/src/lib.rs:53-90):
 *     GGQ_F32, GGQ_F16 or GGQ_BF16.
 *   - Lengths are in the units of the reference slices: `[Blk]` lengths in blocks, `[T]` lengths in
 *     elements.  The check ORDER is the reference's (lib.rs:122-127, 136-141): divisibility first
 *     -> GGQ_ERR_INDIVISIBLE, then length equality -> GGQ_ERR_LENGTH_MISMATCH.
 *   - Results are bit-identical to the reference's per-block functions (see DESIGN.md "Parity").
 *   - There is NO CPU fallback: without a CUDA device every compute call returns GGQ_ERR_CUDA.
 *   - All entry points are thread-safe and re-entrant (the reference is entered concurrently from
 *     one writer thread per output shard, xtask/src/utils/write.rs:64-99).
 */
#ifndef GGQ_H
#define GGQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* GGmlType discriminants — ggus/src/tensor.rs:15-50 (only the types with a codec here) */
enum ggq_type {
    GGQ_F32 = 0,
    GGQ_F16 = 1,   /* 1-element "block": structs/half.rs:8-22  */
    GGQ_Q4_0 = 2,  /* structs/q4_0.rs  18 B / 32 */
    GGQ_Q4_1 = 3,  /* structs/q4_1.rs  20 B / 32 */
    GGQ_Q5_0 = 6,  /* structs/q5_0.rs  22 B / 32 */
    GGQ_Q5_1 = 7,  /* structs/q5_1.rs  24 B / 32 */
    GGQ_Q8_0 = 8,  /* structs/q8_0.rs  34 B / 32 */
    GGQ_Q8_1 = 9,  /* structs/q8_1.rs  36 B / 32 */
    GGQ_Q2K = 10,  /* structs/q2_k.rs  84 B / 256 */
    GGQ_Q3K = 11,  /* structs/q3_k.rs 110 B / 256 */
    GGQ_Q4K = 12,  /* structs/q4_k.rs 144 B / 256 */
    GGQ_Q5K = 13,  /* structs/q5_k.rs 176 B / 256 */
    GGQ_Q6K = 14,  /* structs/q6_k.rs 210 B / 256 */
    GGQ_Q8K = 15,  /* structs/q8_k.rs 290 B / 256 (reference layout: f16 delta) */
    GGQ_BF16 = 30  /* 1-element "block": structs/half.rs:24-38 */
};

/* Return codes.  1 and 2 are `QuantizeError::{Indivisible, LengthMismatch}` (lib.rs:107-113). */
enum ggq_status {
    GGQ_OK = 0,
    GGQ_ERR_INDIVISIBLE = 1,
    GGQ_ERR_LENGTH_MISMATCH = 2,
    GGQ_ERR_UNSUPPORTED = -1, /* unknown type / fdt */
    GGQ_ERR_CUDA = -2,        /* CUDA runtime failure or no device; see ggq_last_error() */
    GGQ_ERR_INVALID = -3      /* null pointer with non-zero length, bad device index, ... */
};

/* `DataBlock::COUNT` and `size_of::<Blk>()` (lib.rs:11-21; ggus/src/tensor.rs:74-79). */
int ggq_block_info(uint32_t type, uint32_t *elems, uint32_t *bytes);

/* Human-readable description of the calling thread's last non-OK return. */
const char *ggq_last_error(void);

/* Number of CUDA devices visible (0 when none); and the device the calling thread's *_slice
 * calls run on (default: the current CUDA device of the thread). */
int ggq_device_count(void);
int ggq_set_device(int device);

/* Spread every large host-pointer call THE CALLING THREAD makes afterwards (slice API, ggq_slices, ggq_cast)
 * over the GPUs 0 .. n_devices-1 — blocks are independent (lib.rs:129-131), so there is no inter-GPU
 * traffic: each GPU pulls and pushes its own ranges over its own PCIe link.  The split is the one
 * ggq_plan_shards() reports.  <= 0 selects all visible devices.  Returns the device count now in effect
 * (>= 1) or a negative status.  Default 1.  Per-thread state (other threads' calls are unaffected); a
 * thread that called ggq_set_device() is not sharded.  The caller's current CUDA device is restored
 * before the call returns. */
int ggq_set_shard_devices(int n_devices);

/* ---- host-pointer slice API (the drop-in) ------------------------------------------------- */

/* `QuantExt::<T, N>::quantize_slice(dst: &mut [Blk], src: &[T])` — lib.rs:121-133; called from
 * xtask/src/utils/operator/cast.rs:140-148.  Host pointers, synchronous: on return every byte of
 * dst[0 .. dst_blocks*block_bytes) has been written.  Pageable or pinned memory accepted; only
 * the Rust alignment (align_of::<T>/<Blk>, i.e. 2 or 4) is assumed. */
int ggq_quantize_slice(uint32_t type, uint32_t fdt, void *dst, size_t dst_blocks,
                       const void *src, size_t src_elems);

/* `QuantExt::<T, N>::dequantize_slice(dst: &mut [T], src: &[Blk])` — lib.rs:135-147; called from
 * cast.rs:150-156. */
int ggq_dequantize_slice(uint32_t type, uint32_t fdt, void *dst, size_t dst_elems,
                         const void *src, size_t src_blocks);

/* Several slice calls as ONE synchronous call: the jobs stream back to back through a single
 * H2D -> kernel -> D2H pipeline, so the copies of job k+1 start while job k is still draining (a
 * sequence of the calls above drains the pipeline between tensors: ~0.2 ms of idle PCIe per call).
 * Every job is validated first with the checks and order of its slice call; the first failing job's
 * status is returned and nothing is computed.  This is what the per-shard writer loop of
 * ggus/src/write/file_writer.rs:121-134 becomes when it hands the library all of a shard's tensors. */
struct ggq_slice_job {
    uint32_t type;     /* block type */
    uint32_t fdt;      /* float-side type */
    int quantize;      /* non-zero: quantize_slice(dst: blocks, src: elements); zero: dequantize_slice */
    void *dst;
    size_t dst_len;    /* blocks when quantizing, elements when dequantizing */
    const void *src;
    size_t src_len;    /* elements when quantizing, blocks when dequantizing */
};
int ggq_slices(const struct ggq_slice_job *jobs, size_t n_jobs);

/* The library's multi-GPU partitioner, as a pure host function (no CUDA call, usable without a GPU): how a
 * ggq_slices / slice / cast call with these jobs is split over `n_devices` GPUs — "by tensor and by block
 * range"  as one rule: the jobs are laid end to end in units of 2^20 elements (a multiple of
 * every block size and of 8 blocks, so both sides of every cut stay 16-byte aligned), a unit weighs the
 * bytes it moves over PCIe (input + output representation), and device d takes the units whose weight
 * midpoint lies in [W d / n, W (d+1) / n).  Pieces come out in job order, element ranges ascending;
 * together they cover every element of every job exactly once.  Fewer devices are used when a device
 * would move less than 16 MiB.  Writes at most `cap` pieces to `out` (may be NULL) and returns the number
 * of pieces of the full plan; 0 when a job fails the slice calls' length checks. */
struct ggq_shard_piece {
    uint32_t job;      /* index into `jobs` */
    int device;        /* 0 .. n_devices-1 */
    size_t elem_begin; /* element range [elem_begin, elem_end) of that job's float side */
    size_t elem_end;
};
size_t ggq_plan_shards(const struct ggq_slice_job *jobs, size_t n_jobs, int n_devices,
                       struct ggq_shard_piece *out, size_t cap);

/* ---- device-pointer slice API (device-resident chains, pipelining, kernel timing) ---------- */

/* Same contracts, but `dst`/`src` are device pointers on the current device and the work is
 * enqueued on `stream` (a cudaStream_t; NULL = legacy default stream) without synchronising.
 * Pointers need only the Rust alignment; 16-byte aligned buffers take the fast (TMA) path. */
int ggq_quantize_slice_device(uint32_t type, uint32_t fdt, void *dst, size_t dst_blocks,
                              const void *src, size_t src_elems, void *stream);
int ggq_dequantize_slice_device(uint32_t type, uint32_t fdt, void *dst, size_t dst_elems,
                                const void *src, size_t src_blocks, void *stream);

/* Several device-pointer slice calls as ONE call, enqueued on `stream` in job order without synchronising
 * (the device-side counterpart of ggq_slices).  Every job is validated first, with the checks and order of its
 * slice call; on a failure nothing is enqueued.  Runs of consecutive dequantize jobs with the same `fdt` share a
 * single grid — one descriptor-table launch whose CTAs each take one tile of whichever tensor they fall into —
 * so the small tensors of a step (a 4096x4096 attention weight is a 7-9 us kernel on its own) pay one ramp and
 * one tail between them instead of one each.  Results are identical to the per-call entry points. */
int ggq_slices_device(const struct ggq_slice_job *jobs, size_t n_jobs, void *stream);

/* ---- tensor casts (the caller of the slice API) ---------------------------------------------- */

/* `cast(row, data, from, to)` — xtask/src/utils/operator/cast.rs:93-138, for every pair of supported
 * tensor types (F32, F16, BF16 and the block types), including the pairs the reference leaves as
 * `todo!()` (cast.rs:132-135) or sends into unbounded recursion (cast.rs:136; a quantized source is
 * mediated by F32 exactly as that arm intends).  `types[0]` is the type of `src`, `types[n-1]` the
 * type of `dst`; intermediate entries apply chained casts (`convert --steps "a -> b -> c"`,
 * xtask/src/convert.rs:38-51) whose intermediates stay on the device.  Host pointers, synchronous.
 * Returns GGQ_ERR_INDIVISIBLE when `n_elems` is not a multiple of every block size in the chain. */
int ggq_cast(const uint32_t *types, int n_types, void *dst, const void *src, size_t n_elems);

/* `GGmlType::size().elements_to_bytes(shape)` (ggus/src/tensor.rs:83-96) for a flat element count;
 * 0 when the type is unknown or `n_elems` is not a whole number of blocks. */
size_t ggq_type_nbytes(uint32_t type, size_t n_elems);

/* ---- block-granular rearrangement (the tensor operators around the casts) ------------------------ */

/* `ArrayLayout<4>` of the crate ndarray-layout 0.2.1 as the reference builds it
 * (xtask/src/utils/operator/merge.rs:359-364 `layout(ty, shape)`: shape[0] counted in blocks, element
 * = one block of `type_size` bytes; operator/permute_qk.rs:57-60: byte-granular rows tiled by head).
 * Shape in elements of `unit` bytes, strides and offset in BYTES (may be negative). */
#define GGQ_MAX_NDIM 4
struct ggq_layout {
    uint32_t ndim;                  /* <= GGQ_MAX_NDIM */
    uint64_t shape[GGQ_MAX_NDIM];
    int64_t strides[GGQ_MAX_NDIM];
    int64_t offset;
};

/* `Rearranging::new(&dst_layout, &src_layout, unit)?.launch(dst, src)` of the crate mem-rearrange 0.1.0
 * — the data movement of `concat` / `split` (merge.rs:288-357: merge-linear / split-linear of attn_qkv,
 * ffn_gate_up, ffn_gate_up_exps) and of `permute_qk` (permute_qk.rs:46-69): for every index of the
 * common shape, `unit` bytes go from src + src.offset + sum(i_k * src.strides[k]) to the same
 * expression over dst.  Bytes of `dst` the layout does not address are left untouched.  Device
 * pointers, enqueued on `stream` without synchronising.  The ranges addressed through `dst` and `src`
 * must not overlap.  Returns GGQ_ERR_LENGTH_MISMATCH when ndim or shape differ (the crate's
 * `SchemeError::ShapeMismatch`), GGQ_ERR_INVALID for ndim > GGQ_MAX_NDIM, unit == 0 or a zero dst
 * stride on a dim of extent > 1. */
int ggq_rearrange_device(void *dst, const struct ggq_layout *dst_layout, const void *src,
                         const struct ggq_layout *src_layout, size_t unit, void *stream);

/* The same on host memory, synchronous: the byte span `src_layout` addresses is copied to the GPU,
 * rearranged there and the span `dst_layout` addresses is copied back (a span the layout covers
 * only partly is uploaded first, so bytes between the addressed runs keep their values). */
int ggq_rearrange(void *dst, const struct ggq_layout *dst_layout, const void *src,
                  const struct ggq_layout *src_layout, size_t unit);

/* ---- whole-file conversion --------------------------------------------------------------------- */

struct ggq_convert_stats {
    uint64_t n_tensors, n_cast_tensors, cast_elems, bytes_in, bytes_out;
    double seconds_plan, seconds_convert, seconds_sync;
    int n_devices;
    int n_out_files;
    uint64_t n_rearranged_tensors; /* tensors evaluated device-resident (rows moved by the rearrange kernel) */
    /* per-stage accounting of the streaming pipeline, summed over the worker threads (so each can exceed the wall
     * clock `seconds_convert`; divide by n_workers for a per-thread average): */
    int n_workers;
    double worker_seconds_read;     /* pread of input bytes into pinned staging */
    double worker_seconds_write;    /* pwrite of output bytes from pinned staging */
    double worker_seconds_gpu_wait; /* blocked on the GPU: H2D + kernels + D2H of a chunk not finished yet */
    uint64_t h2d_bytes, d2h_bytes;  /* bytes that crossed PCIe in each direction */
    int n_direct_inputs;            /* input files actually opened with O_DIRECT (ggq_convert_options.direct_io) */
};

/* `OutputArgs` of xtask (xtask/src/utils/output.rs:8-53); zero means "unlimited" / "off". */
struct ggq_convert_options {
    int n_devices;          /* <= 0: every visible GPU */
    uint64_t max_tensors;   /* -t: max tensors per output shard */
    uint64_t max_bytes;     /* -s: max bytes per output shard (the reference parses "4G", "512M", ...) */
    int no_tensor_first;    /* --no-tensor-first: shard 1 carries only the metadata */
    int no_data;            /* --no-data: write header, KVs and tensor infos only */
    int direct_io;          /* read the tensors that stream through a cast with O_DIRECT: 4 KiB-aligned preads land
                             * straight in the pinned staging buffers and the H2D copy starts at the tensor's offset
                             * inside them — no page-cache copy, for inputs larger than RAM (ggus/src/file.rs:66-145
                             * maps the file instead).  Falls back to buffered reads where the file system refuses
                             * O_DIRECT.  Output bytes are identical either way. */
};

/* `xtask convert FILE --steps "cast:linear:q8_0 embd:q8_0 -> cast:linear:f32 ..."` —
 * xtask/src/convert.rs:24-58 → utils/mod.rs:36-59 → operator/cast.rs:28-90 → utils/write.rs:6-100,
 * for the steps that touch tensor bytes: `cast:<rules>` (operator/cast.rs), `merge-linear`,
 * `split-linear` | `!merge-linear` (operator/merge.rs:22-83) and `permute-qk` (operator/permute_qk.rs:11-44);
 * any other step returns GGQ_ERR_UNSUPPORTED.
 * Tensor-name → target-type rules are cast.rs:28-71 (architectures llama | gpt2 | qwen2 | clip); all
 * steps of a tensor run as one device-resident chain; tensors are spread over `n_devices` GPUs
 * (<= 0: all visible) with no inter-GPU traffic.  The output file is byte-identical to what the
 * reference writer emits: header, `general.alignment`, the other KVs in input order (minus
 * `split.*`), tensor infos, alignment-padded data.  `stats` may be NULL. */
int ggq_convert_gguf(const char *in_path, const char *out_path, const char *steps, int n_devices,
                     struct ggq_convert_stats *stats);

/* The same with the rest of the reference's pipeline around it: several input shards are merged as
 * `Content::new` does (xtask/src/utils/read.rs:5-62: alignment = max, `general.alignment` and `split.*`
 * dropped, duplicate keys / tensor names rejected) and the output is split into shards by the
 * reference's planner (xtask/src/utils/write.rs:23-51 over the byte accounting of
 * ggus/src/write/simulator.rs:75-96).  Shard i of N > 1 is written to
 * "<out_path without .gguf>-0000i-of-0000N.gguf" (ggus/src/name/shard.rs:30-39); only shard 1 carries
 * the metadata KVs (write.rs:72-81).  `opts` may be NULL (defaults: all GPUs, one shard). */
int ggq_convert_gguf_ex(const char *const *in_paths, size_t n_in, const char *out_path, const char *steps,
                        const struct ggq_convert_options *opts, struct ggq_convert_stats *stats);
const char *ggq_convert_last_error(void);

/* ---- memory helpers ------------------------------------------------------------------------ */

/* Page-locked host memory (stands in for cast.rs:158-161 `MmapMut::map_anon` when the caller
 * wants DMA-able buffers).  The host slice API detects pinned pointers and skips its bounce
 * buffers for them. */
void *ggq_host_alloc(size_t bytes);
void ggq_host_free(void *p);

/* Releases the idle stream pipelines (device staging buffers, pinned bounce buffers, streams) the host
 * entry points keep pooled between calls, the device memory cached by the stream-ordered allocator, and the 8-byte
 * ticket counters the Q3_K quantizer keeps per (calling thread, device, stream).  Pipelines in use are not touched and
 * everything is re-created on demand; do not call it while another thread has a `_device` quantize call in flight on a
 * stream (the counter of that launch would be freed under it). */
void ggq_shutdown(void);

/* Number of kernel launches issued by this library since load (all threads); for harnesses. */
uint64_t ggq_launch_count(void);

/* Library version string, e.g. "ggq-b200 0.1 (sm_100a)". */
const char *ggq_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GGQ_H */
