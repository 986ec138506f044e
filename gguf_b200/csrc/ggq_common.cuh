// ggq_common.cuh — shared device helpers for the sm_100a block codec kernels.
//
// Block layouts follow /root/reference/ggml-quants/src/structs/*.rs (byte offsets in SURVEY.md §2.2).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ggq {

// GGmlType discriminants (ggus/src/tensor.rs:15-50)
enum : uint32_t {
    T_F32 = 0, T_F16 = 1, T_Q4_0 = 2, T_Q4_1 = 3, T_Q5_0 = 6, T_Q5_1 = 7, T_Q8_0 = 8, T_Q8_1 = 9,
    T_Q2K = 10, T_Q3K = 11, T_Q4K = 12, T_Q5K = 13, T_Q6K = 14, T_Q8K = 15, T_BF16 = 30
};

template <uint32_t T> struct BlockTraits;
#define GGQ_TRAITS(T, E, B) \
    template <> struct BlockTraits<T> { static constexpr int ELEMS = E; static constexpr int BYTES = B; }
GGQ_TRAITS(T_Q4_0, 32, 18);
GGQ_TRAITS(T_Q4_1, 32, 20);
GGQ_TRAITS(T_Q5_0, 32, 22);
GGQ_TRAITS(T_Q5_1, 32, 24);
GGQ_TRAITS(T_Q8_0, 32, 34);
GGQ_TRAITS(T_Q8_1, 32, 36);
GGQ_TRAITS(T_Q2K, 256, 84);
GGQ_TRAITS(T_Q3K, 256, 110);
GGQ_TRAITS(T_Q4K, 256, 144);
GGQ_TRAITS(T_Q5K, 256, 176);
GGQ_TRAITS(T_Q6K, 256, 210);
GGQ_TRAITS(T_Q8K, 256, 290);
#undef GGQ_TRAITS

// ---------------------------------------------------------------------------------------------
// float <-> f16 / bf16 with the `half` 2.6.0 crate's bit-level semantics (RNE narrow, exact widen,
// NaN quieted keeping the top payload bits).  The hardware converts agree on every non-NaN input;
// the *_exact variants patch NaN payloads and are used by the pure cast kernels.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }

__device__ __forceinline__ float h2f_exact(uint16_t h) {
    float f = h2f(h);
    if ((h & 0x7FFFu) > 0x7C00u)
        f = __uint_as_float(((uint32_t)(h & 0x8000u) << 16) | 0x7FC00000u | ((uint32_t)(h & 0x03FFu) << 13));
    return f;
}
__device__ __forceinline__ float bf2f_exact(uint16_t h) {
    uint32_t i = h;
    if ((i & 0x7FFFu) > 0x7F80u) i |= 0x0040u;
    return __uint_as_float(i << 16);
}
__device__ __forceinline__ uint16_t f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }
__device__ __forceinline__ uint16_t f2h_exact(float f) {
    uint16_t r = f2h(f);
    uint32_t x = __float_as_uint(f);
    if ((x & 0x7FFFFFFFu) > 0x7F800000u)
        r = (uint16_t)(((x & 0x80000000u) >> 16) | 0x7C00u | 0x0200u | ((x & 0x007FFFFFu) >> 13));
    return r;
}
__device__ __forceinline__ uint16_t f2bf_exact(float f) {
    uint32_t x = __float_as_uint(f);
    if ((x & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((x >> 16) | 0x0040u);
    return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}

// exact float(n - bias) for a small unsigned n (n < 2^23), without an I2F conversion
__device__ __forceinline__ float u2f_biased(uint32_t n, float bias) {
    return __fsub_rn(__uint_as_float(0x4B000000u | n), 8388608.0f + bias);
}
// exact float((int8)b)
__device__ __forceinline__ float s8_to_f(uint32_t b) { return u2f_biased((b & 0xFFu) ^ 0x80u, 128.0f); }

// ---------------------------------------------------------------------------------------------
// float-side element types
// ---------------------------------------------------------------------------------------------
struct F32 { using raw = float;    static constexpr int SIZE = 4; static constexpr int V = 4; };
struct F16 { using raw = uint16_t; static constexpr int SIZE = 2; static constexpr int V = 8; };
struct BF16 { using raw = uint16_t; static constexpr int SIZE = 2; static constexpr int V = 8; };

// Store FT::V consecutive results (one 16-byte vector when `vec`, element stores otherwise).
template <class FT> __device__ __forceinline__ void emit(typename FT::raw *p, const float *v, bool vec);
template <> __device__ __forceinline__ void emit<F32>(float *p, const float *v, bool vec) {
    if (vec) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) p[i] = v[i];
    }
}
template <> __device__ __forceinline__ void emit<F16>(uint16_t *p, const float *v, bool vec) {
    __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
    __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
    uint4 w = make_uint4(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1),
                         *reinterpret_cast<uint32_t *>(&h2), *reinterpret_cast<uint32_t *>(&h3));
    if (vec) {
        *reinterpret_cast<uint4 *>(p) = w;
    } else {
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 4; i++) { p[2 * i] = (uint16_t)(ww[i] & 0xFFFFu); p[2 * i + 1] = (uint16_t)(ww[i] >> 16); }
    }
}
template <> __device__ __forceinline__ void emit<BF16>(uint16_t *p, const float *v, bool vec) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 w = make_uint4(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1),
                         *reinterpret_cast<uint32_t *>(&h2), *reinterpret_cast<uint32_t *>(&h3));
    if (vec) {
        *reinterpret_cast<uint4 *>(p) = w;
    } else {
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 4; i++) { p[2 * i] = (uint16_t)(ww[i] & 0xFFFFu); p[2 * i + 1] = (uint16_t)(ww[i] >> 16); }
    }
}

// Load 8 consecutive float-side elements, widened to f32.
template <class FT> __device__ __forceinline__ void load8(const typename FT::raw *p, float *x, bool vec);
template <> __device__ __forceinline__ void load8<F32>(const float *p, float *x, bool vec) {
    if (vec) {
        float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = p[i];
    }
}
__device__ __forceinline__ void load8_u16(const uint16_t *p, uint32_t *w, bool vec) {
    if (vec) {
        uint4 a = __ldg(reinterpret_cast<const uint4 *>(p));
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) w[i] = (uint32_t)p[2 * i] | ((uint32_t)p[2 * i + 1] << 16);
    }
}
template <> __device__ __forceinline__ void load8<F16>(const uint16_t *p, float *x, bool vec) {
    uint32_t w[4];
    load8_u16(p, w, vec);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float2 f = __half22float2(*reinterpret_cast<__half2 *>(&w[i]));
        x[2 * i] = f.x; x[2 * i + 1] = f.y;
    }
}
template <> __device__ __forceinline__ void load8<BF16>(const uint16_t *p, float *x, bool vec) {
    uint32_t w[4];
    load8_u16(p, w, vec);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        x[2 * i] = __uint_as_float(w[i] << 16);
        x[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
}

// ---------------------------------------------------------------------------------------------
// shared-memory byte access at 2-byte alignment (legacy blocks are 18/22/34 bytes wide)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lds16(const uint8_t *p) { return *reinterpret_cast<const uint16_t *>(p); }
// 4 bytes at an address that is ALIGN-aligned (ALIGN in {2,4,...})
template <int ALIGN> __device__ __forceinline__ uint32_t lds32(const uint8_t *p) {
    if constexpr (ALIGN >= 4) return *reinterpret_cast<const uint32_t *>(p);
    else return lds16(p) | (lds16(p + 2) << 16);
}
// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copies (TMA engine; SASS UBLKCP)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy; src/dst 16-byte aligned, bytes % 16 == 0; completes on `bar`
__device__ __forceinline__ void bulk_g2s(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// 16-byte asynchronous global -> shared copies (SASS LDGSTS), per-thread group completion
__device__ __forceinline__ void cp_async16(void *sdst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Programmatic dependent launch (PDL): `pdl_launch_dependents` lets the next kernel in the stream start
// scheduling its CTAs as ours retire; `pdl_wait` blocks until the previous kernel in the stream has
// completed and flushed.  Every kernel calls pdl_wait before its first global access, so stream-order
// semantics are unchanged; only launch latency and the prologue overlap the predecessor's tail.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Byte-exact cooperative copy global -> shared for tiles the bulk path cannot take (tail tile,
// source not 16-byte aligned).  Never reads outside [g, g+n).
__device__ __forceinline__ void cta_copy_g2s(uint8_t *s, const uint8_t *g, uint32_t n, int tid, int nthreads) {
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
        const uint32_t nv = n >> 4;
        for (uint32_t i = tid; i < nv; i += nthreads) reinterpret_cast<uint4 *>(s)[i] = __ldg(reinterpret_cast<const uint4 *>(g) + i);
        for (uint32_t i = (nv << 4) + tid; i < n; i += nthreads) s[i] = g[i];
    } else if ((reinterpret_cast<uintptr_t>(g) & 1u) == 0) {
        const uint32_t nv = n >> 1;
        for (uint32_t i = tid; i < nv; i += nthreads) reinterpret_cast<uint16_t *>(s)[i] = reinterpret_cast<const uint16_t *>(g)[i];
        if ((n & 1u) && tid == 0) s[n - 1] = g[n - 1];
    } else {
        for (uint32_t i = tid; i < n; i += nthreads) s[i] = g[i];
    }
}
// Byte-exact cooperative copy shared -> global.
__device__ __forceinline__ void cta_copy_s2g(uint8_t *g, const uint8_t *s, uint32_t n, int tid, int nthreads) {
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
        const uint32_t nv = n >> 4;
        for (uint32_t i = tid; i < nv; i += nthreads) reinterpret_cast<uint4 *>(g)[i] = reinterpret_cast<const uint4 *>(s)[i];
        for (uint32_t i = (nv << 4) + tid; i < n; i += nthreads) g[i] = s[i];
    } else if ((reinterpret_cast<uintptr_t>(g) & 1u) == 0) {
        const uint32_t nv = n >> 1;
        for (uint32_t i = tid; i < nv; i += nthreads) reinterpret_cast<uint16_t *>(g)[i] = reinterpret_cast<const uint16_t *>(s)[i];
        if ((n & 1u) && tid == 0) g[n - 1] = s[n - 1];
    } else {
        for (uint32_t i = tid; i < n; i += nthreads) g[i] = s[i];
    }
}

}  // namespace ggq
