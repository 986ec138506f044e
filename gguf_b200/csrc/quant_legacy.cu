// quant_legacy.cu — f32 / f16 / bf16 -> Q4_0 Q4_1 Q5_0 Q5_1 Q8_0 Q8_1 (32-element blocks) and Q8K
// (256-element block, reference layout), sm_100a.  Compiled with -fmad=false: the reference's Rust
// never fuses `x * recip + 8.5`, so neither may this file.
//
// Replaces the per-block `Quantize::quantize` bodies behind `QuantExt::quantize_slice`
// (/root/reference/ggml-quants/src/lib.rs:121-133; per-type bodies cited at each encoder).
//
// Shape of the kernel (HBM-bound, input-dominated):
//   * each thread owns 8 consecutive input elements (one 16-byte load for f16/bf16, two for f32) of
//     QROWS independent rows, all loads issued before any use;
//   * 4 lanes own a 32-element block (32 lanes own a Q8K super-block); the per-block folds of
//     structs.rs:91-107 become xor-shuffle reductions whose tie-break is "lowest index wins", which
//     is exactly what the reference's left-to-right strict-compare folds produce;
//   * delta / recip are computed redundantly per lane with IEEE div.rn (identical bits in every lane);
//   * packed blocks are assembled in a shared-memory tile and leave the SM as ONE 1-D bulk async
//     store per tile (`cp.async.bulk.global.shared::cta`, SASS UBLKCP), double buffered.
#include "ggq_common.cuh"
#include "ggq_kernels.h"

namespace ggq {

constexpr int Q_THREADS = 256;
constexpr int Q_ROWS = 4;                                  // independent 8-element chunks per thread
constexpr int Q_TILE_ELEMS = Q_THREADS * 8 * Q_ROWS;       // 8192

constexpr unsigned FULL = 0xFFFFFFFFu;

// ---- folds of structs.rs:91-107 over 8 lane-local elements, then across `LANES` lanes ----------
// max_abs: fold acc.max(|x|) from 0, NaN ignored.
template <int LANES> __device__ __forceinline__ float block_max_abs(const float *x) {
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) { const float a = fabsf(x[i]); acc = a > acc ? a : acc; }
#pragma unroll
    for (int m = 1; m < LANES; m <<= 1) { const float o = __shfl_xor_sync(FULL, acc, m); acc = o > acc ? o : acc; }
    return acc;
}
// max_by_abs: the FIRST x with strictly greatest |x| (sign kept); NaN never selected.
template <int LANES> __device__ __forceinline__ float block_max_by_abs(const float *x, int lane) {
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) acc = fabsf(x[i]) > fabsf(acc) ? x[i] : acc;
#pragma unroll
    for (int m = 1; m < LANES; m <<= 1) {
        const float o = __shfl_xor_sync(FULL, acc, m);
        // the partner with the lower lane id covers lower element indices and wins ties
        const bool lower = (lane & m) == 0;
        const float first = lower ? acc : o, second = lower ? o : acc;
        acc = fabsf(second) > fabsf(first) ? second : first;
    }
    return acc;
}
// min_max: strict-compare folds from (f32::MAX, f32::MIN); first seen wins ties (keeps the sign of
// the first zero, as the reference's x86 lowering does — see oracle/ggq_oracle.c min_max()).
template <int LANES> __device__ __forceinline__ void block_min_max(const float *x, int lane, float &mn, float &mx) {
    float lo = 3.40282347e+38f, hi = -3.40282347e+38f;
#pragma unroll
    for (int i = 0; i < 8; i++) { lo = x[i] < lo ? x[i] : lo; hi = x[i] > hi ? x[i] : hi; }
#pragma unroll
    for (int m = 1; m < LANES; m <<= 1) {
        const float ol = __shfl_xor_sync(FULL, lo, m), oh = __shfl_xor_sync(FULL, hi, m);
        const bool lower = (lane & m) == 0;
        const float fl = lower ? lo : ol, sl = lower ? ol : lo;
        const float fh = lower ? hi : oh, sh = lower ? oh : hi;
        lo = sl < fl ? sl : fl;
        hi = sh > fh ? sh : fh;
    }
    mn = lo;
    mx = hi;
}

// Rust `v as u8` for v already known <= 255 on the high side by the caller's clamp order
__device__ __forceinline__ uint32_t as_u8(float v) { return min(__float2uint_rz(v), 255u); }  // NaN -> 0, neg -> 0
// Rust `v.round() as i8`: half away from zero, saturating, NaN -> 0
__device__ __forceinline__ int round_as_i8(float v) {
    // floor(|v| + 0.5) == trunc(RZ(|v| + 0.5)); RZ keeps the sum from rounding up across an integer
    const float r = __fadd_rz(v, copysignf(0.5f, v));
    return max(-128, min(127, __float2int_rz(r)));
}

// ---- encoders: `x` = this lane's 8 elements, `j` = lane index inside the block group ------------
template <uint32_t T> struct Encoder;

// pack 8 nibble codes c[k] (k-th element of this lane) into two words, one code per byte
__device__ __forceinline__ void bytes8(const uint32_t *c, uint32_t &w0, uint32_t &w1) {
    w0 = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
    w1 = c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24);
}
// 4-bit payload shared by Q4_x / Q5_x: lanes j=0,1 hold elements 0..15 (low nibbles), lanes 2,3 hold
// 16..31 (high nibbles) of the same 16 bytes.  After the exchange lane j writes 4 of the bytes.
template <int QOFF, int ALIGN> __device__ __forceinline__ void store_nibbles(uint8_t *blk, int j, uint32_t w0, uint32_t w1) {
    const uint32_t o0 = __shfl_xor_sync(FULL, w0, 2), o1 = __shfl_xor_sync(FULL, w1, 2);
    if (j < 2) sts32<ALIGN>(blk + QOFF + 8 * j, w0 | (o0 << 4));           // bytes 8j .. 8j+3
    else       sts32<ALIGN>(blk + QOFF + 8 * (j - 2) + 4, o1 | (w1 << 4)); // bytes 8(j-2)+4 .. +7
}

// q4_0.rs:23-44
template <> struct Encoder<T_Q4_0> {
    static constexpr int LANES = 4;
    static __device__ __forceinline__ void run(const float *x, int j, int lane, uint8_t *blk) {
        const float mx = block_max_by_abs<4>(x, lane);
        uint32_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint16_t d16 = 0;
        if (mx != 0.0f) {
            const float d = __fdiv_rn(mx, -8.0f), r = __fdiv_rn(1.0f, d);
            d16 = f2h(d);
#pragma unroll
            for (int k = 0; k < 8; k++) c[k] = as_u8(fminf(__fadd_rn(__fmul_rn(x[k], r), 8.5f), 15.0f));
        }
        uint32_t w0, w1;
        bytes8(c, w0, w1);
        store_nibbles<2, 2>(blk, j, w0, w1);
        if (j == 0) *reinterpret_cast<uint16_t *>(blk) = d16;
    }
};

// q4_1.rs:23-47
template <> struct Encoder<T_Q4_1> {
    static constexpr int LANES = 4;
    static __device__ __forceinline__ void run(const float *x, int j, int lane, uint8_t *blk) {
        float mn, mx;
        block_min_max<4>(x, lane, mn, mx);
        uint32_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint16_t d16 = 0;
        if (mn != mx) {
            const float d = __fdiv_rn(__fsub_rn(mx, mn), 15.0f), r = __fdiv_rn(1.0f, d);
            d16 = f2h(d);
#pragma unroll
            for (int k = 0; k < 8; k++) c[k] = min(as_u8(__fadd_rn(__fmul_rn(__fsub_rn(x[k], mn), r), 0.5f)), 15u);
        }
        uint32_t w0, w1;
        bytes8(c, w0, w1);
        store_nibbles<4, 4>(blk, j, w0, w1);
        if (j == 0) *reinterpret_cast<uint32_t *>(blk) = (uint32_t)d16 | ((uint32_t)f2h(mn) << 16);
    }
};

// q5_0.rs:26-58  (qh bit i = bit 4 of element i's code, for all 32 elements)
template <> struct Encoder<T_Q5_0> {
    static constexpr int LANES = 4;
    static __device__ __forceinline__ void run(const float *x, int j, int lane, uint8_t *blk) {
        const float mx = block_max_by_abs<4>(x, lane);
        uint32_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint16_t d16 = 0;
        if (mx != 0.0f) {
            const float d = __fdiv_rn(mx, -16.0f), r = __fdiv_rn(1.0f, d);
            d16 = f2h(d);
#pragma unroll
            for (int k = 0; k < 8; k++) c[k] = min(as_u8(__fadd_rn(__fmul_rn(x[k], r), 16.5f)), 31u);
        }
        uint32_t hb = 0, n[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { hb |= (c[k] >> 4) << k; n[k] = c[k] & 15u; }
        uint32_t w0, w1;
        bytes8(n, w0, w1);
        store_nibbles<6, 2>(blk, j, w0, w1);
        blk[2 + j] = (uint8_t)hb;
        if (j == 0) *reinterpret_cast<uint16_t *>(blk) = d16;
    }
};

// q5_1.rs:26-62
template <> struct Encoder<T_Q5_1> {
    static constexpr int LANES = 4;
    static __device__ __forceinline__ void run(const float *x, int j, int lane, uint8_t *blk) {
        float mn, mx;
        block_min_max<4>(x, lane, mn, mx);
        uint32_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint16_t d16 = 0;
        if (mn != mx) {
            const float d = __fdiv_rn(__fsub_rn(mx, mn), 31.0f), r = __fdiv_rn(1.0f, d);
            d16 = f2h(d);
#pragma unroll
            for (int k = 0; k < 8; k++) c[k] = min(as_u8(__fadd_rn(__fmul_rn(__fsub_rn(x[k], mn), r), 0.5f)), 31u);
        }
        uint32_t hb = 0, n[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { hb |= (c[k] >> 4) << k; n[k] = c[k] & 15u; }
        uint32_t w0, w1;
        bytes8(n, w0, w1);
        store_nibbles<8, 4>(blk, j, w0, w1);
        blk[4 + j] = (uint8_t)hb;
        if (j == 0) *reinterpret_cast<uint32_t *>(blk) = (uint32_t)d16 | ((uint32_t)f2h(mn) << 16);
    }
};

// q8_0.rs:23-41 and q8_1.rs:28-55 (Q8_1 adds sum = f16(Σq as f32 * delta), delta unrounded)
template <uint32_t T, int QOFF, bool WITH_SUM> struct Encoder8 {
    static constexpr int LANES = 4;
    static __device__ __forceinline__ void run(const float *x, int j, int /*lane*/, uint8_t *blk) {
        const float amax = block_max_abs<4>(x);
        int q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        float d = 0.0f;
        uint16_t d16 = 0;
        if (amax != 0.0f) {
            d = __fdiv_rn(amax, 127.0f);
            const float r = __fdiv_rn(1.0f, d);
            d16 = f2h(d);
#pragma unroll
            for (int k = 0; k < 8; k++) q[k] = round_as_i8(__fmul_rn(x[k], r));
        }
        const uint32_t w0 = (q[0] & 0xFF) | ((q[1] & 0xFF) << 8) | ((q[2] & 0xFF) << 16) | ((uint32_t)q[3] << 24);
        const uint32_t w1 = (q[4] & 0xFF) | ((q[5] & 0xFF) << 8) | ((q[6] & 0xFF) << 16) | ((uint32_t)q[7] << 24);
        constexpr int AL = (QOFF % 4 == 0 && BlockTraits<T>::BYTES % 4 == 0) ? 4 : 2;
        sts32<AL>(blk + QOFF + 8 * j, w0);
        sts32<AL>(blk + QOFF + 8 * j + 4, w1);
        if constexpr (WITH_SUM) {
            int s = q[0] + q[1] + q[2] + q[3] + q[4] + q[5] + q[6] + q[7];
            s += __shfl_xor_sync(FULL, s, 1);
            s += __shfl_xor_sync(FULL, s, 2);
            if (j == 0) {
                const uint16_t s16 = amax != 0.0f ? f2h(__fmul_rn((float)s, d)) : (uint16_t)0;
                *reinterpret_cast<uint32_t *>(blk) = (uint32_t)d16 | ((uint32_t)s16 << 16);
            }
        } else {
            if (j == 0) *reinterpret_cast<uint16_t *>(blk) = d16;
        }
    }
};
template <> struct Encoder<T_Q8_0> : Encoder8<T_Q8_0, 2, false> {};
template <> struct Encoder<T_Q8_1> : Encoder8<T_Q8_1, 4, true> {};

// q8_k.rs:27-54 — reference layout {delta: f16, quants: [i8;256], sums: [i16;16]}, 290 bytes
template <> struct Encoder<T_Q8K> {
    static constexpr int LANES = 32;
    static __device__ __forceinline__ void run(const float *x, int j, int lane, uint8_t *blk) {
        const float mx = block_max_by_abs<32>(x, lane);
        int q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint16_t d16 = 0;
        if (mx != 0.0f) {
            const float d = __fdiv_rn(mx, -127.0f), r = __fdiv_rn(1.0f, d);
            d16 = f2h(d);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                // (x*recip).round().min(127.) as i8 : NaN.round() is NaN, NaN.min(127.) is 127
                const float p = __fmul_rn(x[k], r);
                q[k] = (p != p) ? 127 : min(round_as_i8(p), 127);
            }
        }
        const uint32_t w0 = (q[0] & 0xFF) | ((q[1] & 0xFF) << 8) | ((q[2] & 0xFF) << 16) | ((uint32_t)q[3] << 24);
        const uint32_t w1 = (q[4] & 0xFF) | ((q[5] & 0xFF) << 8) | ((q[6] & 0xFF) << 16) | ((uint32_t)q[7] << 24);
        sts32<2>(blk + 2 + 8 * j, w0);
        sts32<2>(blk + 2 + 8 * j + 4, w1);
        int s = q[0] + q[1] + q[2] + q[3] + q[4] + q[5] + q[6] + q[7];
        s += __shfl_xor_sync(FULL, s, 1);
        if ((j & 1) == 0) *reinterpret_cast<uint16_t *>(blk + 258 + (j >> 1) * 2) = (uint16_t)(int16_t)s;
        if (j == 0) *reinterpret_cast<uint16_t *>(blk) = d16;
    }
};

// ---------------------------------------------------------------------------------------------
template <uint32_t T, class FT>
__global__ void __launch_bounds__(Q_THREADS, 3)
quant_legacy_kernel(const typename FT::raw *__restrict__ src, uint8_t *__restrict__ dst, size_t nblocks) {
    using TR = BlockTraits<T>;
    using E = Encoder<T>;
    constexpr int TILE_BLOCKS = Q_TILE_ELEMS / TR::ELEMS;
    constexpr int TILE_BYTES = TILE_BLOCKS * TR::BYTES;
    static_assert(TILE_BYTES % 16 == 0, "tile must be a whole number of 16-byte chunks");
    __shared__ __align__(128) uint8_t stage[2][TILE_BYTES];

    const int tid = threadIdx.x, lane = tid & 31;
    const int j = tid % E::LANES;
    const size_t ntiles = (nblocks + TILE_BLOCKS - 1) / TILE_BLOCKS;
    const bool vec_in = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
    const bool bulk_out = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;

    int it = 0;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const size_t blk0 = t * (size_t)TILE_BLOCKS;
        const int nb = (int)min((size_t)TILE_BLOCKS, nblocks - blk0);
        const typename FT::raw *in = src + blk0 * TR::ELEMS;
        uint8_t *st = stage[it & 1];

        float x[Q_ROWS][8];
#pragma unroll
        for (int r = 0; r < Q_ROWS; r++) {
            const int e0 = (r * Q_THREADS + tid) * 8;
            if (e0 / TR::ELEMS < nb) {
                load8<FT>(in + e0, x[r], vec_in);
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++) x[r][k] = 0.0f;
            }
        }
        // stage[it&1] was handed to the bulk store two iterations ago; thread 0 confirmed that
        // store had finished READING shared memory before last iteration's barrier (see below).
#pragma unroll
        for (int r = 0; r < Q_ROWS; r++) {
            const int b = ((r * Q_THREADS + tid) * 8) / TR::ELEMS;
            E::run(x[r], j, lane, st + b * TR::BYTES);   // shuffles inside: all lanes participate
        }
        // generic-proxy writes -> async-proxy (bulk store) reads
        fence_proxy_async_smem();
        if (tid == 0) bulk_wait_read<0>();  // every earlier bulk store has released its stage
        __syncthreads();
        uint8_t *out = dst + blk0 * TR::BYTES;
        if (bulk_out && nb == TILE_BLOCKS) {
            if (tid == 0) { bulk_s2g(out, st, TILE_BYTES); bulk_commit(); }
        } else {
            cta_copy_s2g(out, st, (uint32_t)nb * TR::BYTES, tid, Q_THREADS);
            __syncthreads();
        }
    }
    if (tid == 0) bulk_wait_all<0>();
}

template <uint32_t T, class FT>
static cudaError_t launch_quant(const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev) {
    using TR = BlockTraits<T>;
    constexpr int TILE_BLOCKS = Q_TILE_ELEMS / TR::ELEMS;
    auto kern = quant_legacy_kernel<T, FT>;
    static int occ_cache[MAX_DEVICES];
    int ctas_per_sm = 0;
    cudaError_t e = cached_occupancy(kern, Q_THREADS, 0, dev.device, occ_cache, &ctas_per_sm);
    if (e != cudaSuccess) return e;
    const size_t ntiles = (nblocks + TILE_BLOCKS - 1) / TILE_BLOCKS;
    size_t grid = (size_t)dev.sm_count * ctas_per_sm;
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, Q_THREADS, 0, stream>>>(static_cast<const typename FT::raw *>(src), static_cast<uint8_t *>(dst), nblocks);
    return cudaGetLastError();
}

template <uint32_t T>
static cudaError_t launch_quant_fdt(uint32_t fdt, const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev) {
    switch (fdt) {
        case T_F32: return launch_quant<T, F32>(src, dst, nblocks, stream, dev);
        case T_F16: return launch_quant<T, F16>(src, dst, nblocks, stream, dev);
        case T_BF16: return launch_quant<T, BF16>(src, dst, nblocks, stream, dev);
    }
    return cudaErrorInvalidValue;
}

cudaError_t quant_blocks_legacy(uint32_t type, uint32_t fdt, const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev) {
    if (nblocks == 0) return cudaSuccess;
    switch (type) {
        case T_Q4_0: return launch_quant_fdt<T_Q4_0>(fdt, src, dst, nblocks, stream, dev);
        case T_Q4_1: return launch_quant_fdt<T_Q4_1>(fdt, src, dst, nblocks, stream, dev);
        case T_Q5_0: return launch_quant_fdt<T_Q5_0>(fdt, src, dst, nblocks, stream, dev);
        case T_Q5_1: return launch_quant_fdt<T_Q5_1>(fdt, src, dst, nblocks, stream, dev);
        case T_Q8_0: return launch_quant_fdt<T_Q8_0>(fdt, src, dst, nblocks, stream, dev);
        case T_Q8_1: return launch_quant_fdt<T_Q8_1>(fdt, src, dst, nblocks, stream, dev);
        case T_Q8K: return launch_quant_fdt<T_Q8K>(fdt, src, dst, nblocks, stream, dev);
    }
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------------------------
// f32 / f16 / bf16 element casts — structs/half.rs:8-38 through the adapters of lib.rs:62-90:
// every cast goes through f32 (widen exact, narrow RNE), NaNs are quieted with the payload kept.
// ---------------------------------------------------------------------------------------------
template <class FT> __device__ __forceinline__ float widen_exact(typename FT::raw v);
template <> __device__ __forceinline__ float widen_exact<F32>(float v) { return v; }
template <> __device__ __forceinline__ float widen_exact<F16>(uint16_t v) { return h2f_exact(v); }
template <> __device__ __forceinline__ float widen_exact<BF16>(uint16_t v) { return bf2f_exact(v); }
template <class FT> __device__ __forceinline__ typename FT::raw narrow_exact(float f);
template <> __device__ __forceinline__ float narrow_exact<F32>(float f) { return f; }
template <> __device__ __forceinline__ uint16_t narrow_exact<F16>(float f) { return f2h_exact(f); }
template <> __device__ __forceinline__ uint16_t narrow_exact<BF16>(float f) { return f2bf_exact(f); }

template <class ST, class DT>
__global__ void __launch_bounds__(256) cast_kernel(const typename ST::raw *__restrict__ src, typename DT::raw *__restrict__ dst, size_t n) {
    using SR = typename ST::raw;
    using DR = typename DT::raw;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
    const size_t nchunks = n / 8;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (vec) {
        for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += stride) {
            alignas(16) SR in[8];
            if constexpr (sizeof(SR) == 4) {
                *reinterpret_cast<uint4 *>(in) = __ldg(reinterpret_cast<const uint4 *>(src + c * 8));
                *reinterpret_cast<uint4 *>(in + 4) = __ldg(reinterpret_cast<const uint4 *>(src + c * 8) + 1);
            } else {
                *reinterpret_cast<uint4 *>(in) = __ldg(reinterpret_cast<const uint4 *>(src + c * 8));
            }
            alignas(16) DR out[8];
#pragma unroll
            for (int k = 0; k < 8; k++) out[k] = narrow_exact<DT>(widen_exact<ST>(in[k]));
            if constexpr (sizeof(DR) == 4) {
                *reinterpret_cast<uint4 *>(dst + c * 8) = *reinterpret_cast<uint4 *>(out);
                *reinterpret_cast<uint4 *>(dst + c * 8 + 4) = *reinterpret_cast<uint4 *>(out + 4);
            } else {
                *reinterpret_cast<uint4 *>(dst + c * 8) = *reinterpret_cast<uint4 *>(out);
            }
        }
        for (size_t i = nchunks * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            dst[i] = narrow_exact<DT>(widen_exact<ST>(src[i]));
    } else {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            dst[i] = narrow_exact<DT>(widen_exact<ST>(src[i]));
    }
}

template <class ST, class DT>
static cudaError_t launch_cast(const void *src, void *dst, size_t n, cudaStream_t stream, DevInfo dev) {
    size_t want = (n / 8 + 255) / 256;
    size_t grid = (size_t)dev.sm_count * 8;
    if (grid > want) grid = want ? want : 1;
    cast_kernel<ST, DT><<<(unsigned)grid, 256, 0, stream>>>(static_cast<const typename ST::raw *>(src), static_cast<typename DT::raw *>(dst), n);
    return cudaGetLastError();
}
template <class ST>
static cudaError_t launch_cast_dst(uint32_t dst_dt, const void *src, void *dst, size_t n, cudaStream_t stream, DevInfo dev) {
    switch (dst_dt) {
        case T_F32: return launch_cast<ST, F32>(src, dst, n, stream, dev);
        case T_F16: return launch_cast<ST, F16>(src, dst, n, stream, dev);
        case T_BF16: return launch_cast<ST, BF16>(src, dst, n, stream, dev);
    }
    return cudaErrorInvalidValue;
}
cudaError_t cast_elems(uint32_t src_dt, uint32_t dst_dt, const void *src, void *dst, size_t n, cudaStream_t stream, DevInfo dev) {
    if (n == 0) return cudaSuccess;
    switch (src_dt) {
        case T_F32: return launch_cast_dst<F32>(dst_dt, src, dst, n, stream, dev);
        case T_F16: return launch_cast_dst<F16>(dst_dt, src, dst, n, stream, dev);
        case T_BF16: return launch_cast_dst<BF16>(dst_dt, src, dst, n, stream, dev);
    }
    return cudaErrorInvalidValue;
}

}  // namespace ggq
