// quant_k.cu — shipped configurations and launchers of the K-quant quantize kernels (quant_k_kernel.cuh).
// Compiled with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false (see the header).
#include "quant_k_kernel.cuh"

#include <mutex>
#include <vector>

namespace ggq {

// Shipped configuration per type (tools/kq_sweep.cu; profiles/r02_kq_sweep*.txt).  The register file is split per SM
// sub-partition (4 x 16 K registers), so occupancy moves in steps of one warp per scheduler: 4 at <= 128 registers, 5 at
// <= 96, 6 at <= 80, 7 at <= 72, 8 at <= 64.  Q4K / Q5K need 128 (x, weights and candidate codes of a 32-element
// sub-block: moving the codes or the weights to shared memory, recomputing the weights, or 96-120 registers were all
// 4-60 % slower: the loop itself keeps ~112 registers live and spills cost more than the fifth warp buys).
//   AF 1  the search's scale * l + min as FMUL2 + FFMA2 by an opaque 1.0 (-1..-1.6 %)
//   CL 1  the search's clamp to [0, nmax] as one VIMNMX.RELU per element instead of two FMNMX (-3.5..-4 %)
//   WARPS 1 (Q2K, Q4K, Q5K): warps are independent (own staging, own scratch, __syncwarp only), and with one pass per warp
//         (below) a one-warp CTA frees its slot the moment its pass ends; also compiles without the 16-byte spill
//         (Q4K 761 -> 754 us, Q5K 610 -> 603; Q2K: 8 warps per scheduler at 64 registers instead of 7)
template <uint32_t T> struct KqShipped { using type = KqCfg<4, KQuant<T>::REGS, 0, 2, 2>; };
template <> struct KqShipped<T_Q4K> { using type = KqCfg<1, 128, 0, 2, 2, 1, 0, 0, 1>; };
template <> struct KqShipped<T_Q5K> { using type = KqCfg<1, 128, 0, 2, 2, 1, 0, 0, 1>; };
template <> struct KqShipped<T_Q2K> { using type = KqCfg<1, 64, 0, 2, 2, 1, 0, 0, 1>; };
// Q6K's hottest pipe is the XU (FRND: 62 % busy against 49 % for the FP32 pipe): rounding every other pair with the two
// magic-number adds instead balances the two (-1 %); all pairs on the FP32 pipe is +3 %.
// CS 1: the clamp to [-32, 31] only for the candidates that can reach a bound (make_qx_quants16): 444 -> 424 us.
template <> struct KqShipped<T_Q6K> { using type = KqCfg<4, 96, 0, 2, 2, 0, 2, 0, 0, 0, 1>; };

// ---- how the warp passes reach the warps (tools/kq_sweep.cu, profiles/r02_kq_sweep_grid_f16.txt) ---------------------
// A warp pass (32 sub-block searches) takes 8-35 us.  A persistent grid with a fixed stride was the slowest way to hand
// the passes out: Q4K 839 us per 58.7 M elements against 760 us for the same binary launched as ONE PASS PER WARP
// (the hardware's CTA scheduler hands a new CTA to whichever slot frees first; the cold load of a new CTA hides behind
// the other 15-31 resident warps), Q5K 682 -> 610, Q2K 670 -> 616, Q6K 490 -> 444; every step towards longer-lived CTAs
// (2, 3, 6 passes per warp) gives part of it back.  Q3K's passes are the shortest (two super-blocks) and the most
// data-dependent (its refinement loops): one-shot CTAs lose the prefetch of the next pass (466 us against 457 persistent);
// what wins there is a persistent grid whose passes are TICKETS of a launch-wide counter: 432 us.
enum class KqGrid { OneShot, Tickets };
template <uint32_t T> struct KqGridOf { static constexpr KqGrid value = KqGrid::OneShot; };
template <> struct KqGridOf<T_Q3K> { static constexpr KqGrid value = KqGrid::Tickets; };

// The ticket counter of a launch: 8 bytes of device memory owned by (calling thread, device, stream) and zeroed in stream
// order before every launch.  Two launches that share a slot were issued by one thread into one stream, so the second
// one's memset and kernel run after the first kernel has finished; launches from other threads or into other streams have
// their own slot.  Slots live until ggq_shutdown() (quant_k_release_work), which also invalidates every thread's cache.
namespace {
struct WorkSlot {
    int device;
    cudaStream_t stream;
    unsigned long long *ptr;
};
std::mutex g_work_mu;
std::vector<unsigned long long *> g_work_all;   // every slot ever allocated, for quant_k_release_work()
std::atomic<uint64_t> g_work_epoch{1};
thread_local std::vector<WorkSlot> t_work;
thread_local uint64_t t_work_epoch = 0;

cudaError_t work_slot(int device, cudaStream_t stream, unsigned long long **out) {
    const uint64_t epoch = g_work_epoch.load(std::memory_order_acquire);
    if (t_work_epoch != epoch) {  // first use on this thread, or the slots were released since
        t_work.clear();
        t_work_epoch = epoch;
    }
    for (const WorkSlot &w : t_work)
        if (w.device == device && w.stream == stream) { *out = w.ptr; return cudaSuccess; }
    unsigned long long *p = nullptr;
    const cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p), sizeof(unsigned long long));
    if (e != cudaSuccess) return e;
    {
        std::lock_guard<std::mutex> lk(g_work_mu);
        g_work_all.push_back(p);
    }
    t_work.push_back({device, stream, p});
    *out = p;
    return cudaSuccess;
}
}  // namespace

void quant_k_release_work() {
    std::vector<unsigned long long *> all;
    {
        std::lock_guard<std::mutex> lk(g_work_mu);
        all.swap(g_work_all);
        g_work_epoch.fetch_add(1, std::memory_order_acq_rel);
    }
    for (unsigned long long *p : all) cudaFree(p);   // cudaFree waits for the device: no launch can still be using it
}

template <uint32_t T, class FT>
static cudaError_t launch_quant_k(const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev) {
    using CFG = typename KqShipped<T>::type;
    constexpr int SBW = 32 / (256 / KQuant<T>::SUB);
    auto kern = quant_k_kernel<T, FT, CFG>;
    const size_t ngroups = (nblocks + SBW - 1) / SBW;
    const size_t want = (ngroups + CFG::WARPS - 1) / CFG::WARPS;   // one pass per warp
    size_t grid = want;
    unsigned long long *work = nullptr;
    if constexpr (KqGridOf<T>::value == KqGrid::Tickets) {
        static std::atomic<int> occ_cache[MAX_DEVICES];
        int ctas_per_sm = 0;
        cudaError_t e = cached_occupancy(kern, CFG::THREADS, 0, dev.device, occ_cache, &ctas_per_sm);
        if (e != cudaSuccess) return e;
        const size_t resident = (size_t)dev.sm_count * ctas_per_sm;
        if (want > resident) {  // more than one pass per warp: a persistent grid, passes beyond each warp's first by ticket
            grid = resident;
            if ((e = work_slot(dev.device, stream, &work)) != cudaSuccess) return e;
            if ((e = cudaMemsetAsync(work, 0, sizeof(unsigned long long), stream)) != cudaSuccess) return e;
        }
    }
    if (grid > 0x7FFFFFFFu) return cudaErrorInvalidValue;   // 2^31 CTAs of >= 512 elements: beyond any addressable tensor
    kern<<<(unsigned)grid, CFG::THREADS, 0, stream>>>(static_cast<const typename FT::raw *>(src), static_cast<uint8_t *>(dst), nblocks, 1.0f, work);
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
