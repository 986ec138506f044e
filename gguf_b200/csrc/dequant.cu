// dequant.cu — packed ggml blocks -> f32 / f16 / bf16, sm_100a.
//
// Replaces the per-block `Quantize::dequantize` bodies behind `QuantExt::dequantize_slice`
// (/root/reference/ggml-quants/src/lib.rs:135-147; per-type bodies cited at each decoder in
// dequant_kernel.cuh).
//
// Shape of the kernel (HBM-bound streaming codec, no tensor cores):
//   * persistent CTAs, static round-robin over tiles of TILE_ELEMS elements;
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

// shipped configuration (chosen with tools/dq_sweep.cu)
constexpr int DQ_THREADS = 256;
constexpr int DQ_STAGES = 3;
constexpr int DQ_MINB = 3;
constexpr int DQ_MODE = 0;
constexpr int DQ_SP = 0;
template <uint32_t T> struct DqTile { static constexpr int ELEMS = 8192; };
template <> struct DqTile<T_Q6K> { static constexpr int ELEMS = 16384; };

template <uint32_t T, class FT>
static cudaError_t launch_dequant(const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev) {
    using TR = BlockTraits<T>;
    constexpr int TILE = DqTile<T>::ELEMS;
    constexpr int TILE_BLOCKS = TILE / TR::ELEMS;
    constexpr int SMEM = dequant_smem_bytes<T, TILE, DQ_STAGES, DQ_MODE>();
    auto kern = dequant_kernel<T, FT, TILE, DQ_STAGES, DQ_THREADS, DQ_MINB, DQ_MODE, DQ_SP>;
    static int occ_cache[MAX_DEVICES];  // per (T, FT) instantiation, per device
    int ctas_per_sm = 0;
    cudaError_t e = cached_occupancy(kern, DQ_THREADS, SMEM, dev.device, occ_cache, &ctas_per_sm);
    if (e != cudaSuccess) return e;
    const size_t ntiles = (nblocks + TILE_BLOCKS - 1) / TILE_BLOCKS;
    size_t grid = DQ_MODE == 0 ? (size_t)dev.sm_count * ctas_per_sm : ntiles;
    if (grid > ntiles) grid = ntiles;
    return launch_pdl(kern, (unsigned)grid, DQ_THREADS, SMEM, stream, static_cast<const uint8_t *>(src),
                      static_cast<typename FT::raw *>(dst), nblocks);
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
