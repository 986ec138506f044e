// quant_k.cu — shipped configurations and launchers of the K-quant quantize kernels (quant_k_kernel.cuh).
// Compiled with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false (see the header).
#include "quant_k_kernel.cuh"

namespace ggq {

// Shipped configuration per type (tools/kq_sweep.cu; profiles/r02_kq_sweep*.txt).  The register file is split per SM
// sub-partition (4 x 16 K registers), so occupancy moves in steps of one warp per scheduler: 4 at <= 128 registers, 5 at
// <= 96, 6 at <= 80, 7 at <= 72, 8 at <= 64.  Q4K / Q5K need 128 (x, weights and candidate codes of a 32-element
// sub-block: moving the codes or the weights to shared memory, recomputing the weights, or one-warp CTAs at 96-120
// registers were all 4-60 % slower: the loop itself keeps ~112 registers live and spills cost more than the fifth warp
// buys).  What ships beyond round 1: the packed affine (AF 1: -1..-1.6 %), and for Q2K one-warp CTAs at 64 registers
// (8 warps per scheduler instead of 7, and no 4-warp granularity in the tail: -5 %).
template <uint32_t T> struct KqShipped { using type = KqCfg<4, KQuant<T>::REGS, 0, 2, 2>; };
// CL 1: the search's clamp to [0, nmax] is one VIMNMX.RELU per element instead of two FMNMX (quant_k_kernel.cuh, clamp0_relu).
template <> struct KqShipped<T_Q4K> { using type = KqCfg<4, 128, 0, 2, 2, 1, 0, 0, 1>; };
template <> struct KqShipped<T_Q5K> { using type = KqCfg<4, 128, 0, 2, 2, 1, 0, 0, 1>; };
template <> struct KqShipped<T_Q2K> { using type = KqCfg<1, 64, 0, 2, 2, 1, 0, 0, 1>; };
// Q6K's hottest pipe is the XU (FRND: 62 % busy against 49 % for the FP32 pipe): rounding every other pair with the two
// magic-number adds instead balances the two (-1 %); all pairs on the FP32 pipe is +3 %.
template <> struct KqShipped<T_Q6K> { using type = KqCfg<4, 96, 0, 2, 2, 0, 2>; };

template <uint32_t T, class FT>
static cudaError_t launch_quant_k(const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev) {
    using CFG = typename KqShipped<T>::type;
    constexpr int SBW = 32 / (256 / KQuant<T>::SUB);
    auto kern = quant_k_kernel<T, FT, CFG>;
    static std::atomic<int> occ_cache[MAX_DEVICES];
    int ctas_per_sm = 0;
    cudaError_t e = cached_occupancy(kern, CFG::THREADS, 0, dev.device, occ_cache, &ctas_per_sm);
    if (e != cudaSuccess) return e;
    const size_t ngroups = (nblocks + SBW - 1) / SBW;
    const size_t want = (ngroups + CFG::WARPS - 1) / CFG::WARPS;
    size_t grid = (size_t)dev.sm_count * ctas_per_sm;
    if (grid > want) grid = want;
    kern<<<(unsigned)grid, CFG::THREADS, 0, stream>>>(static_cast<const typename FT::raw *>(src), static_cast<uint8_t *>(dst), nblocks, 1.0f);
    return cudaGetLastError();
}
template <uint32_t T>
static cudaError_t launch_quant_k_fdt(uint32_t fdt, const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev) {
    switch (fdt) {
        case T_F32: return launch_quant_k<T, F32>(src, dst, nblocks, stream, dev);
        case T_F16: return launch_quant_k<T, F16>(src, dst, nblocks, stream, dev);
        case T_BF16: return launch_quant_k<T, BF16>(src, dst, nblocks, stream, dev);
    }
    return cudaErrorInvalidValue;
}

cudaError_t quant_blocks_k(uint32_t type, uint32_t fdt, const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev) {
    if (nblocks == 0) return cudaSuccess;
    switch (type) {
        case T_Q2K: return launch_quant_k_fdt<T_Q2K>(fdt, src, dst, nblocks, stream, dev);
        case T_Q3K: return launch_quant_k_fdt<T_Q3K>(fdt, src, dst, nblocks, stream, dev);
        case T_Q4K: return launch_quant_k_fdt<T_Q4K>(fdt, src, dst, nblocks, stream, dev);
        case T_Q5K: return launch_quant_k_fdt<T_Q5K>(fdt, src, dst, nblocks, stream, dev);
        case T_Q6K: return launch_quant_k_fdt<T_Q6K>(fdt, src, dst, nblocks, stream, dev);
    }
    return cudaErrorInvalidValue;
}

}  // namespace ggq
