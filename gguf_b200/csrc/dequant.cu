// dequant.cu — packed ggml blocks -> f32 / f16 / bf16, sm_100a.
//
// Replaces the per-block `Quantize::dequantize` bodies behind `QuantExt::dequantize_slice`
// (/root/reference/ggml-quants/src/lib.rs:135-147; per-type bodies cited at each decoder in
// dequant_kernel.cuh).
//
// Shape of the kernel (HBM-bound streaming codec, no tensor cores):
//   * persistent CTAs, static round-robin over tiles of TILE_ELEMS elements (size-adaptive, see below);
//   * the packed bytes of a tile are one contiguous range -> one 1-D bulk async copy (TMA engine,
//     `cp.async.bulk`, SASS UBLKCP) per tile into a STAGES-deep shared-memory ring guarded by
//     mbarriers, issued STAGES tiles ahead by one thread — no LSU instructions or registers spent on loads;
//   * every thread decodes "units" out of shared memory and writes 16-byte vectors, so each warp
//     store covers full 32-byte sectors of the output (which is 64-88 % of all traffic).
// Every multiply in every decoder is exact in f32, so `fma(q, d, m)` equals the
// reference's separate mul + add bit for bit; the only roundings are the final add and the narrow.
#include "dequant_kernel.cuh"
#include "ggq_kernels.h"

namespace ggq {

// Shipped configurations, chosen with tools/dq_sweep.cu (profiles/r01_dq_sweep_pdl*.txt; repeatable to
// +-0.3 %).  Tensors of >= DQ_BIG_ELEMS elements take the BIG config (bigger tiles, fewer stages, 512-
// thread CTAs: less per-tile overhead once ramp/tail no longer matter), smaller ones the SMALL config
// (more, smaller tiles: better balance on 8-10 us launches).
struct DqCfgDefault { static constexpr int TILE = 8192, STAGES = 3, THREADS = 256, MINB = 3; };
struct DqCfg32k2x512 { static constexpr int TILE = 32768, STAGES = 2, THREADS = 512, MINB = 1; };
struct DqCfg16k3x512 { static constexpr int TILE = 16384, STAGES = 3, THREADS = 512, MINB = 1; };
struct DqCfg16k2x512 { static constexpr int TILE = 16384, STAGES = 2, THREADS = 512, MINB = 1; };
struct DqCfg16k3x256 { static constexpr int TILE = 16384, STAGES = 3, THREADS = 256, MINB = 3; };
struct DqCfg8k2x256 { static constexpr int TILE = 8192, STAGES = 2, THREADS = 256, MINB = 3; };
// 16-bit output (f16 / bf16) and the default for every float side
template <uint32_t T, class FT> struct DqBig : DqCfgDefault {};
template <uint32_t T, class FT> struct DqSmall : DqCfgDefault {};
template <class FT> struct DqBig<T_Q8_0, FT> : DqCfg32k2x512 {};
template <class FT> struct DqBig<T_Q8_1, FT> : DqCfg32k2x512 {};
template <class FT> struct DqBig<T_Q8K, FT> : DqCfg32k2x512 {};
template <class FT> struct DqSmall<T_Q8_0, FT> : DqCfg8k2x256 {};
template <class FT> struct DqSmall<T_Q8_1, FT> : DqCfg8k2x256 {};
template <class FT> struct DqSmall<T_Q8K, FT> : DqCfg8k2x256 {};
template <class FT> struct DqBig<T_Q4K, FT> : DqCfg16k3x512 {};
template <class FT> struct DqBig<T_Q5K, FT> : DqCfg32k2x512 {};
template <class FT> struct DqBig<T_Q3K, FT> : DqCfg16k3x256 {};
template <class FT> struct DqBig<T_Q2K, FT> : DqCfg16k2x512 {};
template <class FT> struct DqBig<T_Q6K, FT> : DqCfg16k3x256 {};
template <class FT> struct DqSmall<T_Q6K, FT> : DqCfg16k3x256 {};
// f32 output writes twice the bytes per element: its own optimum (profiles/r01_dq_sweep_f32.txt, +2..15 %)
template <> struct DqBig<T_Q4_0, F32> : DqCfg32k2x512 {};
template <> struct DqBig<T_Q4_1, F32> : DqCfg32k2x512 {};
template <> struct DqBig<T_Q5_0, F32> : DqCfg32k2x512 {};
template <> struct DqBig<T_Q5_1, F32> : DqCfg32k2x512 {};
template <> struct DqBig<T_Q2K, F32> : DqCfg32k2x512 {};
template <> struct DqBig<T_Q3K, F32> : DqCfg32k2x512 {};
template <> struct DqBig<T_Q4K, F32> : DqCfg32k2x512 {};
template <> struct DqBig<T_Q5K, F32> : DqCfg16k2x512 {};
template <> struct DqBig<T_Q6K, F32> : DqCfg16k2x512 {};
template <> struct DqSmall<T_Q8_0, F32> : DqCfg32k2x512 {};
template <> struct DqSmall<T_Q8_1, F32> : DqCfg32k2x512 {};
template <> struct DqSmall<T_Q8K, F32> : DqCfg32k2x512 {};
template <> struct DqSmall<T_Q3K, F32> : DqCfg16k3x512 {};
template <> struct DqSmall<T_Q4K, F32> : DqCfg16k3x512 {};
template <> struct DqSmall<T_Q5K, F32> : DqCfg16k3x512 {};
template <> struct DqSmall<T_Q6K, F32> : DqCfg32k2x512 {};
constexpr size_t DQ_BIG_ELEMS = size_t(32) << 20;
constexpr int DQ_MODE = 0, DQ_SP = 0;

template <uint32_t T, class FT, class CFG>
static cudaError_t launch_dequant_cfg(const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev) {
    using TR = BlockTraits<T>;
    constexpr int TILE_BLOCKS = CFG::TILE / TR::ELEMS;
    constexpr int SMEM = dequant_smem_bytes<T, CFG::TILE, CFG::STAGES, DQ_MODE>();
    auto kern = dequant_kernel<T, FT, CFG::TILE, CFG::STAGES, CFG::THREADS, CFG::MINB, DQ_MODE, DQ_SP>;
    static std::atomic<int> occ_cache[MAX_DEVICES];  // per instantiation, per device
    int ctas_per_sm = 0;
    cudaError_t e = cached_occupancy(kern, CFG::THREADS, SMEM, dev.device, occ_cache, &ctas_per_sm);
    if (e != cudaSuccess) return e;
    const size_t ntiles = (nblocks + TILE_BLOCKS - 1) / TILE_BLOCKS;
    size_t grid = (size_t)dev.sm_count * ctas_per_sm;
    if (grid > ntiles) grid = ntiles;
    return launch_pdl(kern, (unsigned)grid, CFG::THREADS, SMEM, stream, static_cast<const uint8_t *>(src),
                      static_cast<typename FT::raw *>(dst), nblocks);
}

template <uint32_t T, class FT>
static cudaError_t launch_dequant(const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev) {
    if (nblocks * (size_t)BlockTraits<T>::ELEMS >= DQ_BIG_ELEMS) return launch_dequant_cfg<T, FT, DqBig<T, FT>>(src, dst, nblocks, stream, dev);
    return launch_dequant_cfg<T, FT, DqSmall<T, FT>>(src, dst, nblocks, stream, dev);
}

template <uint32_t T>
static cudaError_t launch_dequant_fdt(uint32_t fdt, const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev) {
    switch (fdt) {
        case T_F32: return launch_dequant<T, F32>(src, dst, nblocks, stream, dev);
        case T_F16: return launch_dequant<T, F16>(src, dst, nblocks, stream, dev);
        case T_BF16: return launch_dequant<T, BF16>(src, dst, nblocks, stream, dev);
    }
    return cudaErrorInvalidValue;
}

cudaError_t dequant_blocks(uint32_t type, uint32_t fdt, const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev) {
    if (nblocks == 0) return cudaSuccess;
    switch (type) {
        case T_Q4_0: return launch_dequant_fdt<T_Q4_0>(fdt, src, dst, nblocks, stream, dev);
        case T_Q4_1: return launch_dequant_fdt<T_Q4_1>(fdt, src, dst, nblocks, stream, dev);
        case T_Q5_0: return launch_dequant_fdt<T_Q5_0>(fdt, src, dst, nblocks, stream, dev);
        case T_Q5_1: return launch_dequant_fdt<T_Q5_1>(fdt, src, dst, nblocks, stream, dev);
        case T_Q8_0: return launch_dequant_fdt<T_Q8_0>(fdt, src, dst, nblocks, stream, dev);
        case T_Q8_1: return launch_dequant_fdt<T_Q8_1>(fdt, src, dst, nblocks, stream, dev);
        case T_Q2K: return launch_dequant_fdt<T_Q2K>(fdt, src, dst, nblocks, stream, dev);
        case T_Q3K: return launch_dequant_fdt<T_Q3K>(fdt, src, dst, nblocks, stream, dev);
        case T_Q4K: return launch_dequant_fdt<T_Q4K>(fdt, src, dst, nblocks, stream, dev);
        case T_Q5K: return launch_dequant_fdt<T_Q5K>(fdt, src, dst, nblocks, stream, dev);
        case T_Q6K: return launch_dequant_fdt<T_Q6K>(fdt, src, dst, nblocks, stream, dev);
        case T_Q8K: return launch_dequant_fdt<T_Q8K>(fdt, src, dst, nblocks, stream, dev);
    }
    return cudaErrorInvalidValue;
}

}  // namespace ggq
