// ggml_quants.hpp — C++ host-side mirror of the reference's `ggml-quants` trait surface, over the
// C ABI of libggq.so (include/ggq.h).  The reference is Rust and there is no Rust toolchain in this
// image, so the host side above the C ABI is written in C++ (the Rust shim a maintainer would add is
// in INTEGRATION.md).  Names, argument meaning and error behaviour follow the reference:
//
//   trait DataBlock { const ID; const COUNT; const ZEROS }            ggml-quants/src/lib.rs:11-21
//   trait Quantize<T, N> { fn quantize(&[T;N]) -> Self; fn dequantize(&self) -> [T;N] }   lib.rs:53-59
//   trait QuantExt<T, N> { fn quantize_slice(dst, src); fn dequantize_slice(dst, src) }    lib.rs:98-104
//   enum QuantizeError { Indivisible, LengthMismatch }                 lib.rs:107-113
//   block structs (repr(C))                                            ggml-quants/src/structs/*.rs
//
// All arithmetic runs on the GPU (there is no CPU path); the per-block `quantize` / `dequantize`
// are thin wrappers over the slice calls with one block.
#pragma once

#include <array>
#include <initializer_list>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../../include/ggq.h"
#include "array_layout.hpp"

namespace ggml_quants {

// `half::f16` / `half::bf16`: bit containers; conversions happen on the device.
struct f16 { uint16_t bits; };
struct bf16 { uint16_t bits; };

enum class QuantizeError { Indivisible = GGQ_ERR_INDIVISIBLE, LengthMismatch = GGQ_ERR_LENGTH_MISMATCH };

// `Result<(), QuantizeError>`
class Result {
   public:
    explicit Result(int code) : code_(code) {}
    bool is_ok() const { return code_ == GGQ_OK; }
    bool is_err() const { return code_ != GGQ_OK; }
    int code() const { return code_; }
    // Err(e) for the two reference errors; runtime (CUDA) failures have no Rust counterpart and throw.
    QuantizeError unwrap_err() const {
        if (code_ == GGQ_ERR_INDIVISIBLE || code_ == GGQ_ERR_LENGTH_MISMATCH) return static_cast<QuantizeError>(code_);
        throw std::logic_error("unwrap_err on " + std::string(code_ == GGQ_OK ? "Ok" : ggq_last_error()));
    }
    // `.unwrap()` as at xtask/src/utils/operator/cast.rs:146,154 — panics (throws) on any error.
    void unwrap() const {
        if (code_ != GGQ_OK) throw std::runtime_error(std::string("called unwrap() on an Err value: ") + ggq_last_error());
    }

   private:
    int code_;
};

template <class T> struct FloatSide;
template <> struct FloatSide<float> { static constexpr uint32_t ID = GGQ_F32; };
template <> struct FloatSide<f16> { static constexpr uint32_t ID = GGQ_F16; };
template <> struct FloatSide<bf16> { static constexpr uint32_t ID = GGQ_BF16; };

// ---- block structs: field order and sizes of ggml-quants/src/structs/*.rs (repr(C)) ----------------
struct DeltaMin { f16 delta, min; };  // structs.rs:54-61, repr(C, align(4))

#define GGQ_BLOCK(NAME, TYPE_ID, N, BODY)                      \
    struct NAME {                                              \
        BODY static constexpr uint32_t ID = TYPE_ID;           \
        static constexpr size_t COUNT = N;                     \
        static NAME ZEROS() { NAME z{}; return z; }            \
    }
GGQ_BLOCK(Q4_0, GGQ_Q4_0, 32, f16 delta; uint8_t quants[16];);                                   // q4_0.rs:6-12
GGQ_BLOCK(Q4_1, GGQ_Q4_1, 32, alignas(4) DeltaMin delta_min; uint8_t quants[16];);               // q4_1.rs:6-12
GGQ_BLOCK(Q5_0, GGQ_Q5_0, 32, f16 delta; uint8_t qh[4]; uint8_t ql[16];);                        // q5_0.rs:6-14
GGQ_BLOCK(Q5_1, GGQ_Q5_1, 32, alignas(4) DeltaMin delta_min; uint8_t qh[4]; uint8_t ql[16];);    // q5_1.rs:6-14
GGQ_BLOCK(Q8_0, GGQ_Q8_0, 32, f16 delta; int8_t quants[32];);                                    // q8_0.rs:6-12
GGQ_BLOCK(Q8_1, GGQ_Q8_1, 32, alignas(4) f16 delta; f16 sum; int8_t quants[32];);                // q8_1.rs:8-16
GGQ_BLOCK(Q2K, GGQ_Q2K, 256, uint8_t scales[16]; uint8_t qs[64]; alignas(4) DeltaMin delta_min;); // q2_k.rs:5-13
GGQ_BLOCK(Q3K, GGQ_Q3K, 256, uint8_t hmask[32]; uint8_t qs[64]; uint8_t scales[12]; f16 delta;);  // q3_k.rs:5-15
GGQ_BLOCK(Q4K, GGQ_Q4K, 256, alignas(4) DeltaMin delta_min; uint8_t scales[12]; uint8_t qs[128];); // q4_k.rs:5-13
GGQ_BLOCK(Q5K, GGQ_Q5K, 256, f16 delta; f16 min; uint8_t scales[12]; uint8_t qh[32]; uint8_t qs[128];);  // q5_k.rs:6-18
GGQ_BLOCK(Q6K, GGQ_Q6K, 256, uint8_t ql[128]; uint8_t qh[64]; int8_t scales[16]; f16 delta;);     // q6_k.rs:6-16
GGQ_BLOCK(Q8K, GGQ_Q8K, 256, f16 delta; int8_t quants[256]; int16_t sums[16];);                   // q8_k.rs:7-15
#undef GGQ_BLOCK
static_assert(sizeof(Q4_0) == 18 && sizeof(Q4_1) == 20 && sizeof(Q5_0) == 22 && sizeof(Q5_1) == 24, "legacy layouts");
static_assert(sizeof(Q8_0) == 34 && sizeof(Q8_1) == 36 && sizeof(Q8K) == 290, "8-bit layouts");
static_assert(sizeof(Q2K) == 84 && sizeof(Q3K) == 110 && sizeof(Q4K) == 144 && sizeof(Q5K) == 176 && sizeof(Q6K) == 210, "K layouts");

// f16 / bf16 as 1-element blocks (structs/half.rs:5-6)
template <class Blk> struct BlockInfo { static constexpr uint32_t ID = Blk::ID; static constexpr size_t COUNT = Blk::COUNT; };
template <> struct BlockInfo<f16> { static constexpr uint32_t ID = GGQ_F16; static constexpr size_t COUNT = 1; };
template <> struct BlockInfo<bf16> { static constexpr uint32_t ID = GGQ_BF16; static constexpr size_t COUNT = 1; };

// ---- QuantExt<T, N> for Blk (lib.rs:116-148) ---------------------------------------------------------
template <class Blk, class T> struct QuantExt {
    static constexpr size_t N = BlockInfo<Blk>::COUNT;
    // fn quantize_slice(dst: &mut [Self], src: &[T]) -> Result<(), QuantizeError>
    static Result quantize_slice(Blk *dst, size_t dst_len, const T *src, size_t src_len) {
        return Result(ggq_quantize_slice(BlockInfo<Blk>::ID, FloatSide<T>::ID, dst, dst_len, src, src_len));
    }
    // fn dequantize_slice(dst: &mut [T], src: &[Self]) -> Result<(), QuantizeError>
    static Result dequantize_slice(T *dst, size_t dst_len, const Blk *src, size_t src_len) {
        return Result(ggq_dequantize_slice(BlockInfo<Blk>::ID, FloatSide<T>::ID, dst, dst_len, src, src_len));
    }
};

// ---- several slice calls as one (no counterpart in the reference: what the per-shard writer loop of
// ggus/src/write/file_writer.rs:121-134 becomes when it hands the library all of a shard's tensors) ---------------
template <class Blk, class T> ggq_slice_job quantize_job(Blk *dst, size_t dst_len, const T *src, size_t src_len) {
    return ggq_slice_job{BlockInfo<Blk>::ID, FloatSide<T>::ID, 1, dst, dst_len, src, src_len};
}
template <class Blk, class T> ggq_slice_job dequantize_job(T *dst, size_t dst_len, const Blk *src, size_t src_len) {
    return ggq_slice_job{BlockInfo<Blk>::ID, FloatSide<T>::ID, 0, dst, dst_len, src, src_len};
}
// host pointers, synchronous, one streamed pipeline (split over GPUs after ggq_set_shard_devices on this thread)
inline Result slices(const ggq_slice_job *jobs, size_t n) { return Result(ggq_slices(jobs, n)); }
// device pointers, enqueued on `stream`; consecutive dequantize jobs of one float side share a single grid
inline Result slices_device(const ggq_slice_job *jobs, size_t n, void *stream) { return Result(ggq_slices_device(jobs, n, stream)); }
// how `slices` would be split over n_devices GPUs (pure host function)
inline size_t plan_shards(const ggq_slice_job *jobs, size_t n, int n_devices, ggq_shard_piece *out, size_t cap) {
    return ggq_plan_shards(jobs, n, n_devices, out, cap);
}

// ---- Quantize<T, N> for Blk (lib.rs:53-90): one block at a time --------------------------------------
template <class Blk, class T> struct Quantize {
    static constexpr size_t N = BlockInfo<Blk>::COUNT;
    static Blk quantize(const std::array<T, N> &data) {
        Blk b;
        QuantExt<Blk, T>::quantize_slice(&b, 1, data.data(), N).unwrap();
        return b;
    }
    static std::array<T, N> dequantize(const Blk &b) {
        std::array<T, N> out;
        QuantExt<Blk, T>::dequantize_slice(out.data(), N, &b, 1).unwrap();
        return out;
    }
};

// ---- the caller of the slice API: `cast(row, data, from, to)` (xtask/src/utils/operator/cast.rs:93-138) ----
// `types` lists the GGmlType ids from the source type to the destination type (two entries for a
// plain cast, more for a chained `--steps "a -> b -> c"`); intermediates stay on the device.
inline Result cast(std::initializer_list<uint32_t> types, void *dst, const void *src, size_t n_elems) {
    return Result(ggq_cast(types.begin(), (int)types.size(), dst, src, n_elems));
}
// `GGmlType::size().elements_to_bytes` for a flat element count (ggus/src/tensor.rs:83-96)
inline size_t type_nbytes(uint32_t type, size_t n_elems) { return ggq_type_nbytes(type, n_elems); }

// `Rearranging::new(&dst, &src, unit)?.launch(dst_ptr, src_ptr)` (crate mem-rearrange; call sites
// xtask/src/utils/operator/merge.rs:311-313, 344-350 and permute_qk.rs:61-66) on host memory, run on the GPU.
inline Result rearrange(void *dst, const ndl::ArrayLayout &dst_layout, const void *src, const ndl::ArrayLayout &src_layout, size_t unit) {
    const ggq_layout d = dst_layout.c(), s = src_layout.c();
    return Result(ggq_rearrange(dst, &d, src, &s, unit));
}

// `xtask convert FILE -x STEPS` for `cast:`, `merge-linear`, `split-linear`, `permute-qk` steps (xtask/src/convert.rs:24-58); throws on failure.
inline ggq_convert_stats convert(const std::string &file, const std::string &out, const std::string &steps, int n_devices = 0) {
    ggq_convert_stats st{};
    const int rc = ggq_convert_gguf(file.c_str(), out.c_str(), steps.c_str(), n_devices, &st);
    if (rc != GGQ_OK) throw std::runtime_error(std::string("convert: ") + ggq_convert_last_error());
    return st;
}

}  // namespace ggml_quants
