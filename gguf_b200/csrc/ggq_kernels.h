// ggq_kernels.h — internal launch entry points shared by the kernel TUs and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <stddef.h>
#include <stdint.h>

namespace ggq {

constexpr int MAX_DEVICES = 32;

struct DevInfo {
    int device;    // CUDA ordinal the launch targets (current device of the caller)
    int sm_count;  // multiProcessorCount
};

// packed blocks -> float side.  `type` is a block type (legacy / K / Q8K).
cudaError_t dequant_blocks(uint32_t type, uint32_t fdt, const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev);
// the same for several tensors of one float side in ONE grid (descriptor table; see dequant.cu)
struct DequantJob {
    uint32_t type;
    const void *src;
    void *dst;
    size_t nblocks;
};
cudaError_t dequant_blocks_batch(uint32_t fdt, const DequantJob *jobs, size_t n, cudaStream_t stream, DevInfo dev, uint64_t *launches);
// float side -> packed blocks, legacy 32-element blocks and Q8K.
cudaError_t quant_blocks_legacy(uint32_t type, uint32_t fdt, const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev);
// float side -> packed blocks, K-quants (Q2K..Q6K).
cudaError_t quant_blocks_k(uint32_t type, uint32_t fdt, const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev);
// frees the ticket counters quant_blocks_k keeps per (thread, device, stream); called by ggq_shutdown()
void quant_k_release_work();
// element casts between f32 / f16 / bf16 (the 1-element "blocks" of structs/half.rs).
cudaError_t cast_elems(uint32_t src_dt, uint32_t dst_dt, const void *src, void *dst, size_t n, cudaStream_t stream, DevInfo dev);

// Strided byte-run copy (rearrange.cu): byte strides / offset, at most 4 dims, same shape on both sides.
struct StridedLayout {
    int ndim;
    uint64_t shape[4];
    int64_t strides[4];
    int64_t offset;
};
cudaError_t rearrange_strided(void *dst_base, const StridedLayout &dst, const void *src_base, const StridedLayout &src, size_t unit,
                              cudaStream_t stream, uint64_t *launches);

// occupancy cache helper: resident CTAs per SM for `kern`, after raising its dynamic smem limit.
template <class K>
static inline cudaError_t cached_occupancy(K kern, int threads, int smem, int device, std::atomic<int> *cache, int *out) {
    int v = cache[device].load(std::memory_order_relaxed);  // racing first callers compute the same value
    if (v == 0) {
        cudaError_t e = cudaSuccess;
        if (smem > 0) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        int n = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem);
        if (e != cudaSuccess) return e;
        v = n > 0 ? n : 1;
        cache[device].store(v, std::memory_order_relaxed);
    }
    *out = v;
    return cudaSuccess;
}

// Launch with the programmatic-stream-serialization attribute (PDL); kernels call pdl_wait() first.
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, unsigned block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace ggq
