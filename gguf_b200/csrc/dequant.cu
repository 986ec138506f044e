// dequant.cu — packed ggml blocks -> f32 / f16 / bf16, sm_100a.
//
// Replaces the per-block `Quantize::dequantize` bodies behind `QuantExt::dequantize_slice`
// (/root/reference/ggml-quants/src/lib.rs:135-147; per-type bodies cited at each decoder in
// dequant_kernel.cuh).
//
// Shape of the kernel (HBM-bound streaming codec, no tensor cores):
//   * tiles of TILE_ELEMS elements; big tensors run one tile per short-lived CTA, small ones persistent CTAs
//     with a ring of stages (size-adaptive, see the configuration table below);
//   * the packed bytes of a tile are one contiguous range -> one 1-D bulk async copy (TMA engine,
//     `cp.async.bulk`, SASS UBLKCP) per tile into shared memory, completion on an mbarrier, issued by one
//     thread — no LSU instructions or registers spent on loads;
//   * every thread decodes "units" out of shared memory and writes 16-byte vectors, so each warp
//     store covers full 32-byte sectors of the output (which is 64-88 % of all traffic).
// Every multiply in every decoder is exact in f32, so `fma(q, d, m)` equals the
// reference's separate mul + add bit for bit; the only roundings are the final add and the narrow.
#include <cstdlib>

#include "dequant_kernel.cuh"
#include "ggq_kernels.h"

namespace ggq {

// Shipped configurations, chosen with tools/dq_sweep.cu (profiles/r01_dq_sweep_modes*.txt; repeatable to
// +-0.5 %).  Kernel shapes (dequant_kernel.cuh):
//   ONE  (MODE 1)  one tile per CTA, 128-thread CTAs, the register cap forced down (MINB) so that 8-12 CTAs are
//                  resident per SM: while one CTA waits for its bulk copy its neighbours decode and store, and the
//                  hardware block scheduler keeps the window of addresses in flight contiguous.  A persistent
//                  grid-stride fill tops out at 5.3-5.6 TB/s on this GPU, the same fill as one 16 KB tile per CTA
//                  reaches 6.3 TB/s; the decoders follow: 6.3-6.6 TB/s (96-101 % of the measured copy peak) on
//                  58.7 M elements against 5.7-5.9 TB/s for the ring.  Needs a few waves of CTAs to overlap.
//   MANY (MODE 3)  the same with 2-4 smaller consecutive tiles per CTA, all requested up front: the instruction-
//                  heavier Q2K..Q5K decoders gain another 1.5-2.5 points on big tensors.
//   RING (MODE 0)  persistent CTAs with a STAGES-deep bulk-copy ring: still the better shape for small tensors
//                  (less than ~2 waves of tiles), where ONE's CTAs would all load and then all store in step.  Small
//                  tiles (4096 elements) and small CTAs (128 threads, 8 per SM) balance these 3-9 us launches best.
template <int TILE_, int STAGES_, int THREADS_, int MINB_, int MODE_> struct DqCfg {
    static constexpr int TILE = TILE_, STAGES = STAGES_, THREADS = THREADS_, MINB = MINB_, MODE = MODE_;
};
using Ring8k3 = DqCfg<8192, 3, 256, 3, 0>;
using Ring16k3 = DqCfg<16384, 3, 256, 3, 0>;
using Ring4k3x128 = DqCfg<4096, 3, 128, 8, 0>;   // small CTAs, 8 per SM: +3..5 % on 8-24 Mi-element tensors (r01_dq_sweep_ring_small.txt)
using Ring4k3x256 = DqCfg<4096, 3, 256, 4, 0>;   // best below ~6 Mi elements
using Ring8k3x128 = DqCfg<8192, 3, 128, 6, 0>;
using One16k8 = DqCfg<16384, 1, 128, 8, 1>;
using One8k8 = DqCfg<8192, 1, 128, 8, 1>;
using Two8k10 = DqCfg<8192, 2, 128, 10, 3>;      // MODE 3: two consecutive tiles per short-lived CTA, both requested up front
using Four4k8 = DqCfg<4096, 4, 128, 8, 3>;
using One4k10 = DqCfg<4096, 1, 128, 10, 1>;

constexpr size_t Mi = size_t(1) << 20;
template <uint32_t T> constexpr bool is_q8_family() { return T == T_Q8_0 || T == T_Q8_1 || T == T_Q8K; }
template <uint32_t T> constexpr bool is_k2345() { return T == T_Q2K || T == T_Q3K || T == T_Q4K || T == T_Q5K; }

// 16-bit output, big tensors
template <uint32_t T> struct DqOneBig { using type = One16k8; };
// the instruction-heavier K decoders gain another 1-2.5 % from several smaller tiles per CTA (r01_dq_sweep_mode3.txt)
template <> struct DqOneBig<T_Q2K> { using type = Four4k8; };
template <> struct DqOneBig<T_Q3K> { using type = Two8k10; };
template <> struct DqOneBig<T_Q4K> { using type = Two8k10; };
template <> struct DqOneBig<T_Q5K> { using type = Two8k10; };
// small tensors that stay on the ring (16-bit output, 6 Mi elements and up)
template <uint32_t T> struct DqRingSmall { using type = Ring4k3x128; };
template <> struct DqRingSmall<T_Q6K> { using type = Ring16k3; };
template <> struct DqRingSmall<T_Q8_0> { using type = Ring8k3x128; };
template <> struct DqRingSmall<T_Q8_1> { using type = Ring8k3x128; };
template <> struct DqRingSmall<T_Q8K> { using type = Ring8k3x128; };

template <uint32_t T, class FT, class CFG>
static cudaError_t launch_dequant_cfg(const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev) {
    using TR = BlockTraits<T>;
    constexpr int TILE_BLOCKS = CFG::TILE / TR::ELEMS;
    constexpr int SMEM = dequant_smem_bytes<T, CFG::TILE, CFG::STAGES, CFG::MODE>();
    auto kern = dequant_kernel<T, FT, CFG::TILE, CFG::STAGES, CFG::THREADS, CFG::MINB, CFG::MODE, 0>;
    static std::atomic<int> occ_cache[MAX_DEVICES];  // per instantiation, per device
    int ctas_per_sm = 0;
    cudaError_t e = cached_occupancy(kern, CFG::THREADS, SMEM, dev.device, occ_cache, &ctas_per_sm);
    if (e != cudaSuccess) return e;
    const size_t ntiles = (nblocks + TILE_BLOCKS - 1) / TILE_BLOCKS;
    // RING: one persistent CTA per slot.  ONE: one CTA per tile, or per STAGES consecutive tiles (MODE 3)
    const size_t want = CFG::MODE == 3 ? (ntiles + CFG::STAGES - 1) / CFG::STAGES : ntiles;
    size_t grid = CFG::MODE == 0 ? (size_t)dev.sm_count * ctas_per_sm : (size_t)0x7FFFFFFF;
    if (grid > want) grid = want;
    return launch_pdl(kern, (unsigned)grid, CFG::THREADS, SMEM, stream, static_cast<const uint8_t *>(src),
                      static_cast<typename FT::raw *>(dst), nblocks);
}

template <uint32_t T, class FT>
static cudaError_t launch_dequant(const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev) {
    const size_t n = nblocks * (size_t)BlockTraits<T>::ELEMS;
    if constexpr (FT::SIZE == 4) {  // f32 output: twice the bytes per element, ONE wins from ~6 Mi elements up
        if (n >= 12 * Mi) return launch_dequant_cfg<T, FT, One8k8>(src, dst, nblocks, stream, dev);
        if (n >= 6 * Mi || is_q8_family<T>()) return launch_dequant_cfg<T, FT, One4k10>(src, dst, nblocks, stream, dev);
        return launch_dequant_cfg<T, FT, Ring8k3>(src, dst, nblocks, stream, dev);
    } else {
        if (n >= (is_k2345<T>() ? 32 : 24) * Mi) return launch_dequant_cfg<T, FT, typename DqOneBig<T>::type>(src, dst, nblocks, stream, dev);
        if (n >= 6 * Mi) return launch_dequant_cfg<T, FT, typename DqRingSmall<T>::type>(src, dst, nblocks, stream, dev);
        if constexpr (T == T_Q6K) return launch_dequant_cfg<T, FT, Ring4k3x128>(src, dst, nblocks, stream, dev);
        else return launch_dequant_cfg<T, FT, Ring4k3x256>(src, dst, nblocks, stream, dev);
    }
}

// ---- batched launch: the tiles of up to DQ_BATCH_MAX tensors (any mix of block types, one float side) in ONE grid ----
// A 4096x4096 tensor is a 7-9 us launch of which ramp and tail are a quarter; back to back with PDL the
// attention-shaped tensors of a step still ran at 79-84 % of the copy peak.  Here every CTA takes one tile of
// whichever tensor its block index falls into (descriptor table in the kernel parameters, a CTA-uniform switch
// over the block type), so a step's tensors share one ramp and one tail and the hardware block scheduler streams
// through all of them like through one big tensor (shape ONE above: short-lived 128-thread CTAs, 8 per SM, one
// bulk copy per tile).
constexpr int DQ_BATCH_MAX = 16;
struct DqBatchJob {
    const uint8_t *src;
    void *dst;
    unsigned long long nblocks;
    unsigned int tile0;  // first tile (= block index) of this job in the grid
    uint32_t type;
};
struct DqBatch {
    DqBatchJob job[DQ_BATCH_MAX];
    int n;
};

// NT consecutive tiles of one job per CTA, all bulk copies requested up front (the MANY shape above; NT = 1 is ONE)
template <uint32_t T, class FT, int TILE_ELEMS, int THREADS, int NT>
__device__ __forceinline__ void dequant_tiles(const uint8_t *__restrict__ src, typename FT::raw *__restrict__ dst, size_t nblocks, size_t t0,
                                              uint64_t *bars, uint8_t *stages, int stage_bytes) {
    using TR = BlockTraits<T>;
    constexpr int TILE_BLOCKS = TILE_ELEMS / TR::ELEMS;
    constexpr int TILE_BYTES = TILE_BLOCKS * TR::BYTES;
    constexpr int UNITS = Decoder<T>::template units<FT::V>();
    static_assert(TILE_BYTES % 16 == 0, "tile must be a whole number of 16-byte chunks");
    const int tid = threadIdx.x;
    const size_t full_tiles = nblocks / TILE_BLOCKS;
    const size_t ntiles = full_tiles + (nblocks % TILE_BLOCKS ? 1 : 0);
    const bool vec = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
    const bool src_fast = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NT; i++)
            if (t0 + i < full_tiles && src_fast) {
                mbar_expect_tx(&bars[i], TILE_BYTES);
                bulk_g2s(stages + (size_t)i * stage_bytes, src + (t0 + i) * (size_t)TILE_BYTES, TILE_BYTES, &bars[i]);
            }
    }
#pragma unroll
    for (int i = 0; i < NT; i++) {
        const size_t t = t0 + i;
        if (t >= ntiles) break;
        uint8_t *stage = stages + (size_t)i * stage_bytes;
        const int nb = t < full_tiles ? TILE_BLOCKS : (int)(nblocks % TILE_BLOCKS);
        if (t < full_tiles && src_fast) {
            mbar_wait(&bars[i], 0);
        } else {
            cta_copy_g2s(stage, src + t * (size_t)TILE_BYTES, (uint32_t)nb * TR::BYTES, tid, THREADS);
            __syncthreads();
        }
        typename FT::raw *out = dst + t * (size_t)TILE_ELEMS;
        bool bad = false;  // see dequant_kernel
        if (nb == TILE_BLOCKS) {
#pragma unroll 2
            for (int u = tid; u < TILE_BLOCKS * UNITS; u += THREADS)
                bad |= Decoder<T>::template run<FT, 0>(stage + (u / UNITS) * TR::BYTES, u % UNITS, out + (size_t)(u / UNITS) * TR::ELEMS, vec);
        } else {
            for (int u = tid; u < nb * UNITS; u += THREADS)
                bad |= Decoder<T>::template run<FT, 0>(stage + (u / UNITS) * TR::BYTES, u % UNITS, out + (size_t)(u / UNITS) * TR::ELEMS, vec);
        }
#ifndef GGQ_AB_NO_EXACT_PASS
        if (bad) fix_units_exact<T, FT, THREADS>(stage, out, nb * UNITS, tid);
#endif
    }
}

template <class FT, int TILE_ELEMS, int THREADS, int MINB, int NT>
__global__ void __launch_bounds__(THREADS, MINB) dequant_batch_kernel(const __grid_constant__ DqBatch batch) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    uint8_t *stages = smem + 128;
    constexpr int STAGE_BYTES = (TILE_ELEMS / 256 * 290 + 15) & ~15;  // the widest tile: Q8K
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NT; i++) mbar_init(&bars[i], 1);
        fence_barrier_init();
    }
    pdl_launch_dependents();
    __syncthreads();
    pdl_wait();
    int j = 0;  // the job this CTA's tiles belong to (CTA-uniform)
#pragma unroll
    for (int k = 1; k < DQ_BATCH_MAX; k++)
        if (k < batch.n && blockIdx.x >= batch.job[k].tile0) j = k;
    const DqBatchJob &jb = batch.job[j];
    const size_t t0 = (size_t)(blockIdx.x - jb.tile0) * NT;
    typename FT::raw *dst = static_cast<typename FT::raw *>(jb.dst);
#define GGQ_DQ_CASE(T) \
    case T: dequant_tiles<T, FT, TILE_ELEMS, THREADS, NT>(jb.src, dst, (size_t)jb.nblocks, t0, bars, stages, STAGE_BYTES); break;
    switch (jb.type) {
        GGQ_DQ_CASE(T_Q4_0) GGQ_DQ_CASE(T_Q4_1) GGQ_DQ_CASE(T_Q5_0) GGQ_DQ_CASE(T_Q5_1) GGQ_DQ_CASE(T_Q8_0) GGQ_DQ_CASE(T_Q8_1)
        GGQ_DQ_CASE(T_Q2K) GGQ_DQ_CASE(T_Q3K) GGQ_DQ_CASE(T_Q4K) GGQ_DQ_CASE(T_Q5K) GGQ_DQ_CASE(T_Q6K) GGQ_DQ_CASE(T_Q8K)
    }
#undef GGQ_DQ_CASE
}

static int block_elems_of(uint32_t t) { return (t >= T_Q2K && t <= T_Q8K) ? 256 : 32; }
static int block_bytes_of(uint32_t t) {
    switch (t) {
        case T_Q4_0: return 18; case T_Q4_1: return 20; case T_Q5_0: return 22; case T_Q5_1: return 24; case T_Q8_0: return 34; case T_Q8_1: return 36;
        case T_Q2K: return 84; case T_Q3K: return 110; case T_Q4K: return 144; case T_Q5K: return 176; case T_Q6K: return 210; case T_Q8K: return 290;
    }
    return 0;
}

template <class FT, int TILE, int THREADS, int MINB, int NT>
static cudaError_t launch_dequant_batch(const DequantJob *jobs, size_t n, cudaStream_t stream, DevInfo dev, uint64_t *launches) {
    constexpr int SMEM = 128 + NT * ((TILE / 256 * 290 + 15) & ~15) + 16;
    auto kern = dequant_batch_kernel<FT, TILE, THREADS, MINB, NT>;
    static std::atomic<int> occ_cache[MAX_DEVICES];
    int ctas_per_sm = 0;
    cudaError_t e = cached_occupancy(kern, THREADS, SMEM, dev.device, occ_cache, &ctas_per_sm);  // also raises the smem limit
    if (e != cudaSuccess) return e;
    size_t i = 0;
    while (i < n) {
        DqBatch b{};
        unsigned long long ctas = 0;
        for (; i < n && b.n < DQ_BATCH_MAX; i++) {
            if (jobs[i].nblocks == 0) continue;
            if (!block_bytes_of(jobs[i].type)) return cudaErrorInvalidValue;
            const unsigned long long tb = TILE / block_elems_of(jobs[i].type);
            const unsigned long long nt = ((jobs[i].nblocks + tb - 1) / tb + NT - 1) / NT;  // CTAs of this job
            if (ctas + nt > 0x7FFFFFFFull) {  // grid.x limit: close this batch (a single job beyond it is split by the caller)
                if (b.n == 0) return cudaErrorInvalidValue;
                break;
            }
            b.job[b.n++] = {static_cast<const uint8_t *>(jobs[i].src), jobs[i].dst, (unsigned long long)jobs[i].nblocks, (unsigned int)ctas, jobs[i].type};
            ctas += nt;
        }
        if (b.n == 0) continue;
        e = launch_pdl(kern, (unsigned)ctas, THREADS, SMEM, stream, b);
        if (e != cudaSuccess) return e;
        if (launches) ++*launches;
    }
    return cudaSuccess;
}

cudaError_t dequant_blocks_batch(uint32_t fdt, const DequantJob *jobs, size_t n, cudaStream_t stream, DevInfo dev, uint64_t *launches) {
    switch (fdt) {
        case T_F32: return launch_dequant_batch<F32, 8192, 128, 8, 1>(jobs, n, stream, dev, launches);
#ifdef GGQ_BATCH_EXPERIMENT  /* tools/gpu_round.sh batch_cfg: other launch shapes of the same kernel, f16 only */
        case T_F16: {
            static const int cfg = getenv("GGQ_BATCH_CFG") ? atoi(getenv("GGQ_BATCH_CFG")) : 0;
            switch (cfg) {
                case 1: return launch_dequant_batch<F16, 8192, 128, 8, 1>(jobs, n, stream, dev, launches);
                case 2: return launch_dequant_batch<F16, 8192, 128, 8, 2>(jobs, n, stream, dev, launches);
                case 3: return launch_dequant_batch<F16, 8192, 128, 10, 2>(jobs, n, stream, dev, launches);
                case 4: return launch_dequant_batch<F16, 4096, 128, 8, 4>(jobs, n, stream, dev, launches);
                case 5: return launch_dequant_batch<F16, 16384, 256, 4, 1>(jobs, n, stream, dev, launches);
                case 6: return launch_dequant_batch<F16, 16384, 128, 6, 2>(jobs, n, stream, dev, launches);
                case 7: return launch_dequant_batch<F16, 32768, 256, 3, 1>(jobs, n, stream, dev, launches);
                default: return launch_dequant_batch<F16, 16384, 128, 8, 1>(jobs, n, stream, dev, launches);
            }
        }
#else
        case T_F16: return launch_dequant_batch<F16, 16384, 128, 8, 1>(jobs, n, stream, dev, launches);
#endif
        case T_BF16: return launch_dequant_batch<BF16, 16384, 128, 8, 1>(jobs, n, stream, dev, launches);
    }
    return cudaErrorInvalidValue;
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
