// dequant.cu — packed ggml blocks -> f32 / f16 / bf16, sm_100a.
//
// Replaces the per-block `Quantize::dequantize` bodies behind `QuantExt::dequantize_slice`
// (/root/reference/ggml-quants/src/lib.rs:135-147; per-type bodies cited at each decoder).
//
// Shape of the kernel (HBM-bound streaming codec, no tensor cores):
//   * persistent CTAs, static round-robin over tiles of TILE_ELEMS elements;
//   * the packed bytes of a tile are one contiguous range -> one 1-D bulk async copy (TMA engine,
//     `cp.async.bulk`, SASS UBLKCP) per tile into an NSTAGE shared-memory ring guarded by mbarriers,
//     issued NSTAGE tiles ahead by one thread — no LSU instructions or registers spent on loads;
//   * every thread decodes "units" out of shared memory and writes 16-byte vectors, so each warp
//     store covers full 32-byte sectors of the output (which is 64-88 % of all traffic).
// Every multiply in every decoder is exact in f32, so `fma(q, d, m)` equals the
// reference's separate mul + add bit for bit; the only roundings are the final add and the narrow.
#include "ggq_common.cuh"
#include "ggq_kernels.h"

namespace ggq {

constexpr int DQ_THREADS = 256;
constexpr int DQ_STAGES = 3;

template <uint32_t T> struct DqTile { static constexpr int ELEMS = 16384; };
template <> struct DqTile<T_Q8_0> { static constexpr int ELEMS = 8192; };
template <> struct DqTile<T_Q8_1> { static constexpr int ELEMS = 8192; };
template <> struct DqTile<T_Q8K> { static constexpr int ELEMS = 8192; };

// V bytes from shared memory at 2-byte alignment, as bytes[]
template <int V, int ALIGN> __device__ __forceinline__ void lds_bytes(const uint8_t *p, uint32_t *b) {
#pragma unroll
    for (int w = 0; w < V / 4; w++) {
        uint32_t x = lds32<ALIGN>(p + 4 * w);
        b[4 * w] = x & 0xFF; b[4 * w + 1] = (x >> 8) & 0xFF; b[4 * w + 2] = (x >> 16) & 0xFF; b[4 * w + 3] = x >> 24;
    }
}

// Decoder<T>::UNITS(V) units per block; unit `u` of block `b` writes its outputs under `out`
// (pointer to the block's first output element).
template <uint32_t T> struct Decoder;

// ---- Q4_0: q4_0.rs:46-57   y[i] = ((b&15) - 8) * d ; y[i+16] = ((b>>4) - 8) * d ----------------
template <> struct Decoder<T_Q4_0> {
    template <int V> static __host__ __device__ constexpr int units() { return 16 / V; }
    template <class FT> static __device__ __forceinline__ void run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V;
        const float d = h2f((uint16_t)lds16(b));
        uint32_t q[V];
        lds_bytes<V, 2>(b + 2 + u * V, q);
        float lo[V], hi[V];
#pragma unroll
        for (int k = 0; k < V; k++) {
            lo[k] = __fmul_rn(u2f_biased(q[k] & 15u, 8.0f), d);
            hi[k] = __fmul_rn(u2f_biased(q[k] >> 4, 8.0f), d);
        }
        emit<FT>(out + u * V, lo, vec);
        emit<FT>(out + 16 + u * V, hi, vec);
    }
};

// ---- Q4_1: q4_1.rs:49-60   y = q * d + m -------------------------------------------------------
template <> struct Decoder<T_Q4_1> {
    template <int V> static __host__ __device__ constexpr int units() { return 16 / V; }
    template <class FT> static __device__ __forceinline__ void run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V;
        const uint32_t dm = lds32<4>(b);
        const float d = h2f((uint16_t)(dm & 0xFFFF)), m = h2f((uint16_t)(dm >> 16));
        uint32_t q[V];
        lds_bytes<V, 4>(b + 4 + u * V, q);
        float lo[V], hi[V];
#pragma unroll
        for (int k = 0; k < V; k++) {
            lo[k] = __fmaf_rn(u2f_biased(q[k] & 15u, 0.0f), d, m);
            hi[k] = __fmaf_rn(u2f_biased(q[k] >> 4, 0.0f), d, m);
        }
        emit<FT>(out + u * V, lo, vec);
        emit<FT>(out + 16 + u * V, hi, vec);
    }
};

// ---- Q5_0: q5_0.rs:60-73   5th bit of element i is bit i of qh ---------------------------------
template <> struct Decoder<T_Q5_0> {
    template <int V> static __host__ __device__ constexpr int units() { return 16 / V; }
    template <class FT> static __device__ __forceinline__ void run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V;
        const float d = h2f((uint16_t)lds16(b));
        const uint32_t qh = lds32<2>(b + 2);
        uint32_t q[V];
        lds_bytes<V, 2>(b + 6 + u * V, q);
        const uint32_t hl = qh >> (u * V), hh = qh >> (16 + u * V);
        float lo[V], hi[V];
#pragma unroll
        for (int k = 0; k < V; k++) {
            lo[k] = __fmul_rn(u2f_biased((q[k] & 15u) | (((hl >> k) & 1u) << 4), 16.0f), d);
            hi[k] = __fmul_rn(u2f_biased((q[k] >> 4) | (((hh >> k) & 1u) << 4), 16.0f), d);
        }
        emit<FT>(out + u * V, lo, vec);
        emit<FT>(out + 16 + u * V, hi, vec);
    }
};

// ---- Q5_1: q5_1.rs:64-77 -----------------------------------------------------------------------
template <> struct Decoder<T_Q5_1> {
    template <int V> static __host__ __device__ constexpr int units() { return 16 / V; }
    template <class FT> static __device__ __forceinline__ void run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V;
        const uint32_t dm = lds32<4>(b);
        const float d = h2f((uint16_t)(dm & 0xFFFF)), m = h2f((uint16_t)(dm >> 16));
        const uint32_t qh = lds32<4>(b + 4);
        uint32_t q[V];
        lds_bytes<V, 4>(b + 8 + u * V, q);
        const uint32_t hl = qh >> (u * V), hh = qh >> (16 + u * V);
        float lo[V], hi[V];
#pragma unroll
        for (int k = 0; k < V; k++) {
            lo[k] = __fmaf_rn(u2f_biased((q[k] & 15u) | (((hl >> k) & 1u) << 4), 0.0f), d, m);
            hi[k] = __fmaf_rn(u2f_biased((q[k] >> 4) | (((hh >> k) & 1u) << 4), 0.0f), d, m);
        }
        emit<FT>(out + u * V, lo, vec);
        emit<FT>(out + 16 + u * V, hi, vec);
    }
};

// ---- Q8_0 / Q8_1 / Q8K: q8_0.rs:43-47, q8_1.rs:57-61, q8_k.rs:56-60   y = q * d ----------------
template <uint32_t T, int QOFF> struct Decoder8 {
    template <int V> static __host__ __device__ constexpr int units() { return BlockTraits<T>::ELEMS / V; }
    template <class FT> static __device__ __forceinline__ void run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V;
        const float d = h2f((uint16_t)lds16(b));
        uint32_t q[V];
        lds_bytes<V, (QOFF % 4 == 0 && BlockTraits<T>::BYTES % 4 == 0) ? 4 : 2>(b + QOFF + u * V, q);
        float y[V];
#pragma unroll
        for (int k = 0; k < V; k++) y[k] = __fmul_rn(s8_to_f(q[k]), d);
        emit<FT>(out + u * V, y, vec);
    }
};
template <> struct Decoder<T_Q8_0> : Decoder8<T_Q8_0, 2> {};
template <> struct Decoder<T_Q8_1> : Decoder8<T_Q8_1, 4> {};
template <> struct Decoder<T_Q8K> : Decoder8<T_Q8K, 2> {};

// ---- K-quants: layouts structs/q{2..6}_k.rs; arithmetic = upstream ggml dequantize_row_qN_K ----
// 6-bit (scale, min) pair j of the 12-byte Q4K/Q5K table held as three words
__device__ __forceinline__ void scale_min_k4(int j, uint32_t s0, uint32_t s1, uint32_t s2, uint32_t &sc, uint32_t &m) {
    // bytes 0..3 = s0, 4..7 = s1, 8..11 = s2
    const int jj = j & 3;
    const uint32_t a = (s0 >> (8 * jj)) & 0xFF;  // s[jj]
    const uint32_t bq = (s1 >> (8 * jj)) & 0xFF; // s[jj+4]
    const uint32_t c = (s2 >> (8 * jj)) & 0xFF;  // s[jj+8]
    if (j < 4) {
        sc = a & 63u;
        m = bq & 63u;
    } else {
        sc = (c & 0xFu) | ((a >> 6) << 4);
        m = (c >> 4) | ((bq >> 6) << 4);
    }
}

template <> struct Decoder<T_Q4K> {
    template <int V> static __host__ __device__ constexpr int units() { return 128 / V; }
    template <class FT> static __device__ __forceinline__ void run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V, CP = 32 / V;
        const int p = u / CP, c = u % CP;
        const uint4 hdr = *reinterpret_cast<const uint4 *>(b);  // delta, min, scales[12]
        const float d = h2f((uint16_t)(hdr.x & 0xFFFF)), dmin = h2f((uint16_t)(hdr.x >> 16));
        uint32_t sc1, m1, sc2, m2;
        scale_min_k4(2 * p, hdr.y, hdr.z, hdr.w, sc1, m1);
        scale_min_k4(2 * p + 1, hdr.y, hdr.z, hdr.w, sc2, m2);
        const float d1 = __fmul_rn(d, (float)sc1), mm1 = __fmul_rn(dmin, (float)m1);
        const float d2 = __fmul_rn(d, (float)sc2), mm2 = __fmul_rn(dmin, (float)m2);
        uint32_t q[V];
        lds_bytes<V, 4>(b + 16 + 32 * p + c * V, q);
        float lo[V], hi[V];
#pragma unroll
        for (int k = 0; k < V; k++) {
            lo[k] = __fmaf_rn(d1, u2f_biased(q[k] & 15u, 0.0f), -mm1);
            hi[k] = __fmaf_rn(d2, u2f_biased(q[k] >> 4, 0.0f), -mm2);
        }
        emit<FT>(out + 64 * p + c * V, lo, vec);
        emit<FT>(out + 64 * p + 32 + c * V, hi, vec);
    }
};

template <> struct Decoder<T_Q5K> {
    template <int V> static __host__ __device__ constexpr int units() { return 128 / V; }
    template <class FT> static __device__ __forceinline__ void run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V, CP = 32 / V;
        const int p = u / CP, c = u % CP;
        const uint4 hdr = *reinterpret_cast<const uint4 *>(b);
        const float d = h2f((uint16_t)(hdr.x & 0xFFFF)), dmin = h2f((uint16_t)(hdr.x >> 16));
        uint32_t sc1, m1, sc2, m2;
        scale_min_k4(2 * p, hdr.y, hdr.z, hdr.w, sc1, m1);
        scale_min_k4(2 * p + 1, hdr.y, hdr.z, hdr.w, sc2, m2);
        const float d1 = __fmul_rn(d, (float)sc1), mm1 = __fmul_rn(dmin, (float)m1);
        const float d2 = __fmul_rn(d, (float)sc2), mm2 = __fmul_rn(dmin, (float)m2);
        uint32_t q[V], h[V];
        lds_bytes<V, 4>(b + 48 + 32 * p + c * V, q);
        lds_bytes<V, 4>(b + 16 + c * V, h);
        float lo[V], hi[V];
#pragma unroll
        for (int k = 0; k < V; k++) {
            const uint32_t ql = (q[k] & 15u) | (((h[k] >> (2 * p)) & 1u) << 4);
            const uint32_t qu = (q[k] >> 4) | (((h[k] >> (2 * p + 1)) & 1u) << 4);
            lo[k] = __fmaf_rn(d1, u2f_biased(ql, 0.0f), -mm1);
            hi[k] = __fmaf_rn(d2, u2f_biased(qu, 0.0f), -mm2);
        }
        emit<FT>(out + 64 * p + c * V, lo, vec);
        emit<FT>(out + 64 * p + 32 + c * V, hi, vec);
    }
};

template <> struct Decoder<T_Q6K> {
    template <int V> static __host__ __device__ constexpr int units() { return 64 / V; }
    template <class FT> static __device__ __forceinline__ void run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V, CP = 32 / V;
        const int n = u / CP, l0 = (u % CP) * V;
        const float d = h2f((uint16_t)lds16(b + 208));
        uint32_t qa[V], qb[V], qh[V];
        lds_bytes<V, 2>(b + 64 * n + l0, qa);
        lds_bytes<V, 2>(b + 64 * n + 32 + l0, qb);
        lds_bytes<V, 2>(b + 128 + 32 * n + l0, qh);
        const uint8_t *sc = b + 192 + 8 * n + l0 / 16;
        const float s1 = __fmul_rn(d, s8_to_f(sc[0])), s2 = __fmul_rn(d, s8_to_f(sc[2]));
        const float s3 = __fmul_rn(d, s8_to_f(sc[4])), s4 = __fmul_rn(d, s8_to_f(sc[6]));
        float y1[V], y2[V], y3[V], y4[V];
#pragma unroll
        for (int k = 0; k < V; k++) {
            y1[k] = __fmul_rn(s1, u2f_biased((qa[k] & 15u) | (((qh[k] >> 0) & 3u) << 4), 32.0f));
            y2[k] = __fmul_rn(s2, u2f_biased((qb[k] & 15u) | (((qh[k] >> 2) & 3u) << 4), 32.0f));
            y3[k] = __fmul_rn(s3, u2f_biased((qa[k] >> 4) | (((qh[k] >> 4) & 3u) << 4), 32.0f));
            y4[k] = __fmul_rn(s4, u2f_biased((qb[k] >> 4) | (((qh[k] >> 6) & 3u) << 4), 32.0f));
        }
        typename FT::raw *o = out + 128 * n + l0;
        emit<FT>(o, y1, vec);
        emit<FT>(o + 32, y2, vec);
        emit<FT>(o + 64, y3, vec);
        emit<FT>(o + 96, y4, vec);
    }
};

template <> struct Decoder<T_Q2K> {
    template <int V> static __host__ __device__ constexpr int units() { return 64 / V; }
    template <class FT> static __device__ __forceinline__ void run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V, CP = 32 / V;
        const int n = u / CP, l0 = (u % CP) * V;
        const uint32_t dm = lds32<4>(b + 80);
        const float d = h2f((uint16_t)(dm & 0xFFFF)), dmin = h2f((uint16_t)(dm >> 16));
        uint32_t q[V];
        lds_bytes<V, 4>(b + 16 + 32 * n + l0, q);
        const uint8_t *sc = b + 8 * n + l0 / 16;
        typename FT::raw *o = out + 128 * n + l0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t s = sc[2 * k];
            const float dl = __fmul_rn(d, (float)(s & 15u)), ml = __fmul_rn(dmin, (float)(s >> 4));
            float y[V];
#pragma unroll
            for (int i = 0; i < V; i++) y[i] = __fmaf_rn(dl, u2f_biased((q[i] >> (2 * k)) & 3u, 0.0f), -ml);
            emit<FT>(o + 32 * k, y, vec);
        }
    }
};

template <> struct Decoder<T_Q3K> {
    template <int V> static __host__ __device__ constexpr int units() { return 64 / V; }
    template <class FT> static __device__ __forceinline__ void run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V, CP = 32 / V;
        const int n = u / CP, l0 = (u % CP) * V;
        const float d = h2f((uint16_t)lds16(b + 108));
        uint32_t q[V], hm[V];
        lds_bytes<V, 2>(b + 32 + 32 * n + l0, q);
        lds_bytes<V, 2>(b + l0, hm);
        const uint8_t *scales = b + 96;
        typename FT::raw *o = out + 128 * n + l0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int j = 8 * n + 2 * k + l0 / 16;
            const uint32_t lo4 = j < 8 ? (scales[j] & 15u) : (scales[j - 8] >> 4);
            const uint32_t hi2 = (scales[8 + (j & 3)] >> (2 * (j >> 2))) & 3u;
            const float dl = __fmul_rn(d, u2f_biased(lo4 | (hi2 << 4), 32.0f));
            const int bit = 4 * n + k;
            float y[V];
#pragma unroll
            for (int i = 0; i < V; i++) {
                // ((q >> 2k) & 3) - (hbit ? 0 : 4)  ==  (two-bit | hbit << 2) - 4
                const uint32_t v = ((q[i] >> (2 * k)) & 3u) | (((hm[i] >> bit) & 1u) << 2);
                y[i] = __fmul_rn(dl, u2f_biased(v, 4.0f));
            }
            emit<FT>(o + 32 * k, y, vec);
        }
    }
};

// ---------------------------------------------------------------------------------------------
template <uint32_t T, class FT>
__global__ void __launch_bounds__(DQ_THREADS, 3)
dequant_kernel(const uint8_t *__restrict__ src, typename FT::raw *__restrict__ dst, size_t nblocks) {
    using TR = BlockTraits<T>;
    constexpr int TILE_BLOCKS = DqTile<T>::ELEMS / TR::ELEMS;
    constexpr int TILE_BYTES = TILE_BLOCKS * TR::BYTES;
    constexpr int UNITS = Decoder<T>::template units<FT::V>();
    static_assert(TILE_BYTES % 16 == 0, "tile must be a whole number of 16-byte chunks");

    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    uint8_t *stages = smem + 128;

    const int tid = threadIdx.x;
    const size_t full_tiles = nblocks / TILE_BLOCKS;
    const int rem_blocks = (int)(nblocks % TILE_BLOCKS);
    const size_t ntiles = full_tiles + (rem_blocks ? 1 : 0);
    const bool src_fast = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
    const bool vec = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < DQ_STAGES; s++) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncthreads();

    auto issue = [&](size_t i) {  // thread 0 only
        const size_t t = blockIdx.x + i * (size_t)gridDim.x;
        if (t < full_tiles && src_fast) {
            const int s = (int)(i % DQ_STAGES);
            mbar_expect_tx(&bars[s], TILE_BYTES);
            bulk_g2s(stages + (size_t)s * TILE_BYTES, src + t * (size_t)TILE_BYTES, TILE_BYTES, &bars[s]);
        }
    };
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < DQ_STAGES; i++) issue(i);
    }

    size_t i = 0;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
        const int s = (int)(i % DQ_STAGES);
        uint8_t *stage = stages + (size_t)s * TILE_BYTES;
        const bool bulk = (t < full_tiles) && src_fast;
        const int nb = (t < full_tiles) ? TILE_BLOCKS : rem_blocks;
        if (bulk) {
            mbar_wait(&bars[s], (uint32_t)((i / DQ_STAGES) & 1));
        } else {
            cta_copy_g2s(stage, src + t * (size_t)TILE_BYTES, (uint32_t)nb * TR::BYTES, tid, DQ_THREADS);
            __syncthreads();
        }
        typename FT::raw *out = dst + t * (size_t)DqTile<T>::ELEMS;
        const int nunits = nb * UNITS;
        if (nb == TILE_BLOCKS) {
#pragma unroll 2
            for (int u = tid; u < TILE_BLOCKS * UNITS; u += DQ_THREADS)
                Decoder<T>::template run<FT>(stage + (u / UNITS) * TR::BYTES, u % UNITS, out + (size_t)(u / UNITS) * TR::ELEMS, vec);
        } else {
            for (int u = tid; u < nunits; u += DQ_THREADS)
                Decoder<T>::template run<FT>(stage + (u / UNITS) * TR::BYTES, u % UNITS, out + (size_t)(u / UNITS) * TR::ELEMS, vec);
        }
        __syncthreads();  // stage s fully consumed
        if (tid == 0) issue(i + DQ_STAGES);
    }
}

template <uint32_t T, class FT>
static cudaError_t launch_dequant(const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev) {
    using TR = BlockTraits<T>;
    constexpr int TILE_BLOCKS = DqTile<T>::ELEMS / TR::ELEMS;
    constexpr int SMEM = 128 + DQ_STAGES * TILE_BLOCKS * TR::BYTES;
    auto kern = dequant_kernel<T, FT>;
    static int occ_cache[MAX_DEVICES];  // per (T, FT) instantiation, per device
    int ctas_per_sm = 0;
    cudaError_t e = cached_occupancy(kern, DQ_THREADS, SMEM, dev.device, occ_cache, &ctas_per_sm);
    if (e != cudaSuccess) return e;
    const size_t ntiles = (nblocks + TILE_BLOCKS - 1) / TILE_BLOCKS;
    size_t grid = (size_t)dev.sm_count * ctas_per_sm;
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, DQ_THREADS, SMEM, stream>>>(static_cast<const uint8_t *>(src),
                                                       static_cast<typename FT::raw *>(dst), nblocks);
    return cudaGetLastError();
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
