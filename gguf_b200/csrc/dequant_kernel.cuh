// dequant_kernel.cuh — decoders + the tiled dequantize kernel template (see dequant.cu for the
// design notes).  Kept in a header so tools/dq_sweep.cu can instantiate other tile / stage / thread
// configurations of exactly the code that ships.
#pragma once
#include "ggq_common.cuh"

namespace ggq {

// store policy for the output vectors: 0 = default (write-back), 1 = streaming (.cs, evict-first)
template <class FT, int SP> __device__ __forceinline__ void emit_p(typename FT::raw *p, const float *v, bool vec) {
    if constexpr (SP == 0) {
        emit<FT>(p, v, vec);
    } else {
        if (!vec) { emit<FT>(p, v, false); return; }
        if constexpr (FT::SIZE == 4) {
            __stcs(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
        } else {
            alignas(16) typename FT::raw tmp[8];
            emit<FT>(tmp, v, true);
            __stcs(reinterpret_cast<uint4 *>(p), *reinterpret_cast<uint4 *>(tmp));
        }
    }
}



// ---- word-level helpers -----------------------------------------------------------------------
// NW 32-bit words from shared memory at ALIGN-byte alignment
template <int NW, int ALIGN> __device__ __forceinline__ void lds_words(const uint8_t *p, uint32_t (&w)[NW]) {
    if constexpr (ALIGN >= 4) {
#pragma unroll
        for (int i = 0; i < NW; i++) w[i] = *reinterpret_cast<const uint32_t *>(p + 4 * i);
    } else {
        // 2-byte aligned (18/22/34/110/210/290-byte blocks): NW+1 aligned words and a funnel shift
        // instead of 2*NW halfword loads.  May touch up to 2 bytes before / 4 bytes after the range:
        // always inside the stage ring, which is padded by 16 bytes (dequant_smem_bytes).
        const uint32_t a = smem_u32(p);
        const uint32_t sh = (a & 2u) << 3;
        const uint32_t *q = reinterpret_cast<const uint32_t *>(p - (a & 2u));
        uint32_t t[NW + 1];
#pragma unroll
        for (int i = 0; i <= NW; i++) t[i] = q[i];
#pragma unroll
        for (int i = 0; i < NW; i++) w[i] = __funnelshift_r(t[i], t[i + 1], sh);
    }
}
// exact float(byte k of w) - bias: PRMT drops the byte into the mantissa of 2^23, one FADD removes it
__device__ __forceinline__ float bytef(uint32_t w, int k, float bias) {
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u + k)), 8388608.0f + bias);
}
// bits b0..b3 of x -> bit 0 of bytes 0..3
__device__ __forceinline__ uint32_t spread4(uint32_t x) { return ((x & 0xFu) * 0x00204081u) & 0x01010101u; }

// `codes` holds one unsigned code per byte (FT::V codes in FT::V/4 words).
// y = (code - bias) * d, f32 arithmetic then narrow (the reference's op order)
template <class FT, int SP> __device__ __forceinline__ void emit_scaled(typename FT::raw *out, const uint32_t (&codes)[FT::V / 4], float bias, float d, bool vec) {
    float y[FT::V];
#pragma unroll
    for (int k = 0; k < FT::V; k++) y[k] = __fmul_rn(bytef(codes[k >> 2], k & 3, bias), d);
    emit_p<FT, SP>(out, y, vec);
}
// y = code * d + m  (product exact in f32 -> one fused rounding == the reference's mul then add)
template <class FT, int SP> __device__ __forceinline__ void emit_affine(typename FT::raw *out, const uint32_t (&codes)[FT::V / 4], float d, float m, bool vec) {
    float y[FT::V];
#pragma unroll
    for (int k = 0; k < FT::V; k++) y[k] = __fmaf_rn(bytef(codes[k >> 2], k & 3, 0.0f), d, m);
    emit_p<FT, SP>(out, y, vec);
}
// f16 output of y = (code - BIAS) * d where d is itself an f16 value: (code - BIAS) is exact in f16 and
// the f32 product the reference narrows is exact, so ONE half multiply (RNE, subnormals kept) gives
// the same bits as f32-multiply-then-narrow for every non-NaN result.  0x6400 | code == 1024 + code.
template <int BIAS, int SP> __device__ __forceinline__ void emit_scaled_h(uint16_t *out, const uint32_t (&codes)[2], uint32_t dbits, bool vec) {
    const uint32_t d2u = dbits | (dbits << 16);
    const __half2 d2 = *reinterpret_cast<const __half2 *>(&d2u);
    const __half2 off = __floats2half2_rn(1024.0f + BIAS, 1024.0f + BIAS);
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const uint32_t a = __byte_perm(codes[i], 0x64646464u, 0x4140u), c = __byte_perm(codes[i], 0x64646464u, 0x4342u);
        const __half2 ra = __hmul2(__hsub2(*reinterpret_cast<const __half2 *>(&a), off), d2);
        const __half2 rc = __hmul2(__hsub2(*reinterpret_cast<const __half2 *>(&c), off), d2);
        r[2 * i] = *reinterpret_cast<const uint32_t *>(&ra);
        r[2 * i + 1] = *reinterpret_cast<const uint32_t *>(&rc);
    }
    if (vec) {
        const uint4 v = make_uint4(r[0], r[1], r[2], r[3]);
        if constexpr (SP == 0) *reinterpret_cast<uint4 *>(out) = v;
        else __stcs(reinterpret_cast<uint4 *>(out), v);
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) { out[2 * i] = (uint16_t)(r[i] & 0xFFFFu); out[2 * i + 1] = (uint16_t)(r[i] >> 16); }
    }
}
template <class FT> struct IsF16 { static constexpr bool value = false; };
template <> struct IsF16<F16> { static constexpr bool value = true; };

// (code - BIAS) * d with d an f16 field: half fast path for f16 output, f32 path otherwise
template <class FT, int SP, int BIAS> __device__ __forceinline__ void emit_sym(typename FT::raw *out, const uint32_t (&codes)[FT::V / 4], uint32_t dbits, bool vec) {
    if constexpr (IsF16<FT>::value) emit_scaled_h<BIAS, SP>(out, codes, dbits, vec);
    else emit_scaled<FT, SP>(out, codes, (float)BIAS, h2f((uint16_t)dbits), vec);
}

// ---- blocks whose f16 scale fields are NaN or infinite -------------------------------------------------------------
// With finite fields no decoder can produce a NaN (every product and sum stays far inside f32's range), and the fast
// paths below are bit-exact.  A NaN / infinite field is garbage input, but the reference still defines the result: its
// arithmetic runs on x86 SSE, where an operation with a NaN operand returns that operand quieted (payload kept; the
// first operand when both are NaN) and an invalid operation (inf * 0, inf - inf) returns the negative quiet NaN
// 0xFFC00000; `half`'s narrowing then keeps sign and top payload bits.  The GPU would return 0x7FFFFFFF for all of
// these.  So a unit whose block has such a field leaves the fast path (one test of bits already in a register) and
// evaluates the reference's expression element by element, operation by operation, with these rules.
__device__ __forceinline__ bool f16_nonfinite(uint32_t h) { return (h & 0x7C00u) == 0x7C00u; }
__device__ __forceinline__ float x86_result(float r, float a, float b) {
    if (r == r) return r;
    if (a != a) return __uint_as_float(__float_as_uint(a) | 0x00400000u);
    if (b != b) return __uint_as_float(__float_as_uint(b) | 0x00400000u);
    return __uint_as_float(0xFFC00000u);
}
__device__ __forceinline__ float xmul(float a, float b) { return x86_result(__fmul_rn(a, b), a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return x86_result(__fadd_rn(a, b), a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return x86_result(__fsub_rn(a, b), a, b); }
__device__ __forceinline__ float fld16(const uint8_t *p) { return h2f_exact((uint16_t)lds16(p)); }
__device__ __forceinline__ uint32_t fld32u(const uint8_t *p) { return lds16(p) | (lds16(p + 2) << 16); }
// 6-bit (scale, min) pair j of Q4K / Q5K from the 12 bytes at s (upstream get_scale_min_k4)
__device__ __forceinline__ void scale_min_bytes(int j, const uint8_t *s, uint32_t &sc, uint32_t &m) {
    if (j < 4) {
        sc = s[j] & 63u;
        m = s[j + 4] & 63u;
    } else {
        sc = (s[j + 4] & 0xFu) | ((uint32_t)(s[j - 4] >> 6) << 4);
        m = (s[j + 4] >> 4) | ((uint32_t)(s[j] >> 6) << 4);
    }
}

// element i of block b, the reference's expression (structs/*.rs; upstream dequantize_row_qN_K for the K-quants)
template <uint32_t T> __device__ __forceinline__ float ref_elem(const uint8_t *b, int i);
template <> __device__ __forceinline__ float ref_elem<T_Q4_0>(const uint8_t *b, int i) {
    const uint32_t q = b[2 + (i & 15)];
    return xmul((float)((int)(i < 16 ? (q & 0xFu) : (q >> 4)) - 8), fld16(b));
}
template <> __device__ __forceinline__ float ref_elem<T_Q4_1>(const uint8_t *b, int i) {
    const uint32_t q = b[4 + (i & 15)];
    return xadd(xmul((float)(i < 16 ? (q & 0xFu) : (q >> 4)), fld16(b)), fld16(b + 2));
}
__device__ __forceinline__ int code5(const uint8_t *ql, uint32_t qh, int i) {
    const uint32_t q = ql[i & 15];
    return (int)((i < 16 ? (q & 0xFu) : (q >> 4)) | (((qh >> i) & 1u) << 4));
}
template <> __device__ __forceinline__ float ref_elem<T_Q5_0>(const uint8_t *b, int i) {
    return xmul((float)(code5(b + 6, fld32u(b + 2), i) - 16), fld16(b));
}
template <> __device__ __forceinline__ float ref_elem<T_Q5_1>(const uint8_t *b, int i) {
    return xadd(xmul((float)code5(b + 8, fld32u(b + 4), i), fld16(b)), fld16(b + 2));
}
template <> __device__ __forceinline__ float ref_elem<T_Q8_0>(const uint8_t *b, int i) { return xmul((float)(int8_t)b[2 + i], fld16(b)); }
template <> __device__ __forceinline__ float ref_elem<T_Q8_1>(const uint8_t *b, int i) { return xmul((float)(int8_t)b[4 + i], fld16(b)); }
template <> __device__ __forceinline__ float ref_elem<T_Q8K>(const uint8_t *b, int i) { return xmul((float)(int8_t)b[2 + i], fld16(b)); }
template <> __device__ __forceinline__ float ref_elem<T_Q2K>(const uint8_t *b, int i) {
    const int n = i >> 7, j = (i & 127) >> 5, l = i & 31;
    const uint32_t s = b[8 * n + 2 * j + (l >> 4)];
    const float dl = xmul(fld16(b + 80), (float)(s & 0xFu)), ml = xmul(fld16(b + 82), (float)(s >> 4));
    return xsub(xmul(dl, (float)((b[16 + 32 * n + l] >> (2 * j)) & 3u)), ml);
}
template <> __device__ __forceinline__ float ref_elem<T_Q3K>(const uint8_t *b, int i) {
    const int n = i >> 7, j = (i & 127) >> 5, l = i & 31, is = 8 * n + 2 * j + (l >> 4);
    const uint8_t *sc = b + 96;
    const uint32_t lo4 = is < 8 ? (sc[is] & 15u) : (uint32_t)(sc[is - 8] >> 4);
    const uint32_t hi2 = (sc[8 + (is & 3)] >> (2 * (is >> 2))) & 3u;
    const float dl = xmul(fld16(b + 108), (float)((int)(lo4 | (hi2 << 4)) - 32));
    const int q = (int)((b[32 + 32 * n + l] >> (2 * j)) & 3u) - ((b[l] & (1u << (4 * n + j))) ? 0 : 4);
    return xmul(dl, (float)q);
}
template <> __device__ __forceinline__ float ref_elem<T_Q4K>(const uint8_t *b, int i) {
    const int g = i >> 6, hi = (i >> 5) & 1, l = i & 31;
    uint32_t sc, m;
    scale_min_bytes(2 * g + hi, b + 4, sc, m);
    const float d1 = xmul(fld16(b), (float)sc), m1 = xmul(fld16(b + 2), (float)m);
    const uint32_t q = b[16 + 32 * g + l];
    return xsub(xmul(d1, (float)(hi ? (q >> 4) : (q & 0xFu))), m1);
}
template <> __device__ __forceinline__ float ref_elem<T_Q5K>(const uint8_t *b, int i) {
    const int g = i >> 6, hi = (i >> 5) & 1, l = i & 31;
    uint32_t sc, m;
    scale_min_bytes(2 * g + hi, b + 4, sc, m);
    const float d1 = xmul(fld16(b), (float)sc), m1 = xmul(fld16(b + 2), (float)m);
    const uint32_t q = b[48 + 32 * g + l];
    const uint32_t c = (hi ? (q >> 4) : (q & 0xFu)) + ((b[16 + l] & (1u << (2 * g + hi))) ? 16u : 0u);
    return xsub(xmul(d1, (float)c), m1);
}
template <> __device__ __forceinline__ float ref_elem<T_Q6K>(const uint8_t *b, int i) {
    const int n = i >> 7, grp = (i & 127) >> 5, l = i & 31;
    const uint32_t ql = b[64 * n + l + ((grp & 1) ? 32 : 0)], qh = b[128 + 32 * n + l];
    const int q = (int)(((grp & 2) ? (ql >> 4) : (ql & 0xFu)) | (((qh >> (2 * grp)) & 3u) << 4)) - 32;
    const float sc = (float)(int8_t)b[192 + 8 * n + (l >> 4) + 2 * grp];
    return xmul(xmul(fld16(b + 208), sc), (float)q);
}
template <class FT> __device__ __forceinline__ void store_exact(typename FT::raw *p, float y);
template <> __device__ __forceinline__ void store_exact<F32>(float *p, float y) { *p = y; }
template <> __device__ __forceinline__ void store_exact<F16>(uint16_t *p, float y) { *p = f2h_exact(y); }
template <> __device__ __forceinline__ void store_exact<BF16>(uint16_t *p, float y) { *p = f2bf_exact(y); }
// `nruns` runs of V consecutive elements, `stride` elements apart, starting at element s0 of the block: what one unit
// of a fast decoder writes.  Out of line: it must not cost the fast path registers.
template <uint32_t T, class FT>
__device__ __forceinline__ void decode_exact_runs(const uint8_t *b, typename FT::raw *out, int s0, int stride, int nruns) {
    for (int r = 0; r < nruns; r++)
        for (int k = 0; k < FT::V; k++) {
            const int i = s0 + r * stride + k;
            store_exact<FT>(out + i, ref_elem<T>(b, i));
        }
}

// Decoder<T>::units<V>() units per block; unit `u` of block `b` writes its outputs under `out`
// (pointer to the block's first output element).
template <uint32_t T> struct Decoder;

// ---- Q4_0: q4_0.rs:46-57   y[i] = ((b&15) - 8) * d ; y[i+16] = ((b>>4) - 8) * d ----------------
template <> struct Decoder<T_Q4_0> {
    template <int V> static __host__ __device__ constexpr int units() { return 16 / V; }
    template <class FT, int SP> static __device__ __forceinline__ bool run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V, NW = V / 4;
        const uint32_t dbits = lds16(b);
        uint32_t w[NW], lo[NW], hi[NW];
        lds_words<NW, 2>(b + 2 + u * V, w);
#pragma unroll
        for (int i = 0; i < NW; i++) { lo[i] = w[i] & 0x0F0F0F0Fu; hi[i] = (w[i] >> 4) & 0x0F0F0F0Fu; }
        emit_sym<FT, SP, 8>(out + u * V, lo, dbits, vec);
        emit_sym<FT, SP, 8>(out + 16 + u * V, hi, dbits, vec);
        return f16_nonfinite(dbits);
    }
};

// ---- Q4_1: q4_1.rs:49-60   y = q * d + m -------------------------------------------------------
template <> struct Decoder<T_Q4_1> {
    template <int V> static __host__ __device__ constexpr int units() { return 16 / V; }
    template <class FT, int SP> static __device__ __forceinline__ bool run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V, NW = V / 4;
        const uint32_t dm = lds32<4>(b);
        const float d = h2f((uint16_t)(dm & 0xFFFF)), m = h2f((uint16_t)(dm >> 16));
        uint32_t w[NW], lo[NW], hi[NW];
        lds_words<NW, 4>(b + 4 + u * V, w);
#pragma unroll
        for (int i = 0; i < NW; i++) { lo[i] = w[i] & 0x0F0F0F0Fu; hi[i] = (w[i] >> 4) & 0x0F0F0F0Fu; }
        emit_affine<FT, SP>(out + u * V, lo, d, m, vec);
        emit_affine<FT, SP>(out + 16 + u * V, hi, d, m, vec);
        return f16_nonfinite(dm) || f16_nonfinite(dm >> 16);
    }
};

// ---- Q5_0: q5_0.rs:60-73   5th bit of element i is bit i of qh ---------------------------------
template <> struct Decoder<T_Q5_0> {
    template <int V> static __host__ __device__ constexpr int units() { return 16 / V; }
    template <class FT, int SP> static __device__ __forceinline__ bool run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V, NW = V / 4;
        const uint32_t dbits = lds16(b);
        const uint32_t qh = lds32<2>(b + 2);
        uint32_t w[NW], lo[NW], hi[NW];
        lds_words<NW, 2>(b + 6 + u * V, w);
#pragma unroll
        for (int i = 0; i < NW; i++) {
            lo[i] = (w[i] & 0x0F0F0F0Fu) | (spread4(qh >> (u * V + 4 * i)) << 4);
            hi[i] = ((w[i] >> 4) & 0x0F0F0F0Fu) | (spread4(qh >> (16 + u * V + 4 * i)) << 4);
        }
        emit_sym<FT, SP, 16>(out + u * V, lo, dbits, vec);
        emit_sym<FT, SP, 16>(out + 16 + u * V, hi, dbits, vec);
        return f16_nonfinite(dbits);
    }
};

// ---- Q5_1: q5_1.rs:64-77 -----------------------------------------------------------------------
template <> struct Decoder<T_Q5_1> {
    template <int V> static __host__ __device__ constexpr int units() { return 16 / V; }
    template <class FT, int SP> static __device__ __forceinline__ bool run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V, NW = V / 4;
        const uint32_t dm = lds32<4>(b);
        const float d = h2f((uint16_t)(dm & 0xFFFF)), m = h2f((uint16_t)(dm >> 16));
        const uint32_t qh = lds32<4>(b + 4);
        uint32_t w[NW], lo[NW], hi[NW];
        lds_words<NW, 4>(b + 8 + u * V, w);
#pragma unroll
        for (int i = 0; i < NW; i++) {
            lo[i] = (w[i] & 0x0F0F0F0Fu) | (spread4(qh >> (u * V + 4 * i)) << 4);
            hi[i] = ((w[i] >> 4) & 0x0F0F0F0Fu) | (spread4(qh >> (16 + u * V + 4 * i)) << 4);
        }
        emit_affine<FT, SP>(out + u * V, lo, d, m, vec);
        emit_affine<FT, SP>(out + 16 + u * V, hi, d, m, vec);
        return f16_nonfinite(dm) || f16_nonfinite(dm >> 16);
    }
};

// ---- Q8_0 / Q8_1 / Q8K: q8_0.rs:43-47, q8_1.rs:57-61, q8_k.rs:56-60   y = q * d ----------------
template <uint32_t T, int QOFF> struct Decoder8 {
    template <int V> static __host__ __device__ constexpr int units() { return BlockTraits<T>::ELEMS / V; }
    template <class FT, int SP> static __device__ __forceinline__ bool run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V, NW = V / 4;
        const uint32_t dbits = lds16(b);
        uint32_t w[NW];
        lds_words<NW, (QOFF % 4 == 0 && BlockTraits<T>::BYTES % 4 == 0) ? 4 : 2>(b + QOFF + u * V, w);
#pragma unroll
        for (int i = 0; i < NW; i++) w[i] ^= 0x80808080u;  // int8 -> biased unsigned (q + 128)
        emit_sym<FT, SP, 128>(out + u * V, w, dbits, vec);
        return f16_nonfinite(dbits);
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
    template <class FT, int SP> static __device__ __forceinline__ bool run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V, NW = V / 4, CP = 32 / V;
        const int p = u / CP, c = u % CP;
        const uint4 hdr = *reinterpret_cast<const uint4 *>(b);  // delta, min, scales[12]
        const float d = h2f((uint16_t)(hdr.x & 0xFFFF)), dmin = h2f((uint16_t)(hdr.x >> 16));
        uint32_t sc1, m1, sc2, m2;
        scale_min_k4(2 * p, hdr.y, hdr.z, hdr.w, sc1, m1);
        scale_min_k4(2 * p + 1, hdr.y, hdr.z, hdr.w, sc2, m2);
        const float d1 = __fmul_rn(d, u2f_biased(sc1, 0.0f)), mm1 = __fmul_rn(dmin, u2f_biased(m1, 0.0f));
        const float d2 = __fmul_rn(d, u2f_biased(sc2, 0.0f)), mm2 = __fmul_rn(dmin, u2f_biased(m2, 0.0f));
        uint32_t w[NW], lo[NW], hi[NW];
        lds_words<NW, 4>(b + 16 + 32 * p + c * V, w);
#pragma unroll
        for (int i = 0; i < NW; i++) { lo[i] = w[i] & 0x0F0F0F0Fu; hi[i] = (w[i] >> 4) & 0x0F0F0F0Fu; }
        emit_affine<FT, SP>(out + 64 * p + c * V, lo, d1, -mm1, vec);
        emit_affine<FT, SP>(out + 64 * p + 32 + c * V, hi, d2, -mm2, vec);
        return f16_nonfinite(hdr.x) || f16_nonfinite(hdr.x >> 16);
    }
};

template <> struct Decoder<T_Q5K> {
    template <int V> static __host__ __device__ constexpr int units() { return 128 / V; }
    template <class FT, int SP> static __device__ __forceinline__ bool run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V, NW = V / 4, CP = 32 / V;
        const int p = u / CP, c = u % CP;
        const uint4 hdr = *reinterpret_cast<const uint4 *>(b);
        const float d = h2f((uint16_t)(hdr.x & 0xFFFF)), dmin = h2f((uint16_t)(hdr.x >> 16));
        uint32_t sc1, m1, sc2, m2;
        scale_min_k4(2 * p, hdr.y, hdr.z, hdr.w, sc1, m1);
        scale_min_k4(2 * p + 1, hdr.y, hdr.z, hdr.w, sc2, m2);
        const float d1 = __fmul_rn(d, u2f_biased(sc1, 0.0f)), mm1 = __fmul_rn(dmin, u2f_biased(m1, 0.0f));
        const float d2 = __fmul_rn(d, u2f_biased(sc2, 0.0f)), mm2 = __fmul_rn(dmin, u2f_biased(m2, 0.0f));
        uint32_t w[NW], h[NW], lo[NW], hi[NW];
        lds_words<NW, 4>(b + 48 + 32 * p + c * V, w);
        lds_words<NW, 4>(b + 16 + c * V, h);
#pragma unroll
        for (int i = 0; i < NW; i++) {
            lo[i] = (w[i] & 0x0F0F0F0Fu) | (((h[i] >> (2 * p)) & 0x01010101u) << 4);
            hi[i] = ((w[i] >> 4) & 0x0F0F0F0Fu) | (((h[i] >> (2 * p + 1)) & 0x01010101u) << 4);
        }
        emit_affine<FT, SP>(out + 64 * p + c * V, lo, d1, -mm1, vec);
        emit_affine<FT, SP>(out + 64 * p + 32 + c * V, hi, d2, -mm2, vec);
        return f16_nonfinite(hdr.x) || f16_nonfinite(hdr.x >> 16);
    }
};

template <> struct Decoder<T_Q6K> {
    template <int V> static __host__ __device__ constexpr int units() { return 64 / V; }
    template <class FT, int SP> static __device__ __forceinline__ bool run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V, NW = V / 4, CP = 32 / V;
        const int n = u / CP, l0 = (u % CP) * V;
        const uint32_t dbits = lds16(b + 208);
        const float d = h2f((uint16_t)dbits);
        uint32_t qa[NW], qb[NW], qh[NW], c1[NW], c2[NW], c3[NW], c4[NW];
        lds_words<NW, 2>(b + 64 * n + l0, qa);
        lds_words<NW, 2>(b + 64 * n + 32 + l0, qb);
        lds_words<NW, 2>(b + 128 + 32 * n + l0, qh);
        const uint8_t *sc = b + 192 + 8 * n + l0 / 16;
        const float s1 = __fmul_rn(d, s8_to_f(sc[0])), s2 = __fmul_rn(d, s8_to_f(sc[2]));
        const float s3 = __fmul_rn(d, s8_to_f(sc[4])), s4 = __fmul_rn(d, s8_to_f(sc[6]));
#pragma unroll
        for (int i = 0; i < NW; i++) {
            c1[i] = (qa[i] & 0x0F0F0F0Fu) | ((qh[i] & 0x03030303u) << 4);
            c2[i] = (qb[i] & 0x0F0F0F0Fu) | (((qh[i] >> 2) & 0x03030303u) << 4);
            c3[i] = ((qa[i] >> 4) & 0x0F0F0F0Fu) | (((qh[i] >> 4) & 0x03030303u) << 4);
            c4[i] = ((qb[i] >> 4) & 0x0F0F0F0Fu) | (((qh[i] >> 6) & 0x03030303u) << 4);
        }
        typename FT::raw *o = out + 128 * n + l0;
        emit_scaled<FT, SP>(o, c1, 32.0f, s1, vec);
        emit_scaled<FT, SP>(o + 32, c2, 32.0f, s2, vec);
        emit_scaled<FT, SP>(o + 64, c3, 32.0f, s3, vec);
        emit_scaled<FT, SP>(o + 96, c4, 32.0f, s4, vec);
        return f16_nonfinite(dbits);
    }
};

template <> struct Decoder<T_Q2K> {
    template <int V> static __host__ __device__ constexpr int units() { return 64 / V; }
    template <class FT, int SP> static __device__ __forceinline__ bool run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V, NW = V / 4, CP = 32 / V;
        const int n = u / CP, l0 = (u % CP) * V;
        const uint32_t dm = lds32<4>(b + 80);
        const float d = h2f((uint16_t)(dm & 0xFFFF)), dmin = h2f((uint16_t)(dm >> 16));
        uint32_t w[NW];
        lds_words<NW, 4>(b + 16 + 32 * n + l0, w);
        const uint8_t *sc = b + 8 * n + l0 / 16;
        typename FT::raw *o = out + 128 * n + l0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t s = sc[2 * k];
            const float dl = __fmul_rn(d, u2f_biased(s & 15u, 0.0f)), ml = __fmul_rn(dmin, u2f_biased(s >> 4, 0.0f));
            uint32_t c[NW];
#pragma unroll
            for (int i = 0; i < NW; i++) c[i] = (w[i] >> (2 * k)) & 0x03030303u;
            emit_affine<FT, SP>(o + 32 * k, c, dl, -ml, vec);
        }
        return f16_nonfinite(dm) || f16_nonfinite(dm >> 16);
    }
};

template <> struct Decoder<T_Q3K> {
    template <int V> static __host__ __device__ constexpr int units() { return 64 / V; }
    template <class FT, int SP> static __device__ __forceinline__ bool run(const uint8_t *b, int u, typename FT::raw *out, bool vec) {
        constexpr int V = FT::V, NW = V / 4, CP = 32 / V;
        const int n = u / CP, l0 = (u % CP) * V;
        const uint32_t dbits = lds16(b + 108);
        const float d = h2f((uint16_t)dbits);
        uint32_t w[NW], hm[NW];
        lds_words<NW, 2>(b + 32 + 32 * n + l0, w);
        lds_words<NW, 2>(b + l0, hm);
        const uint8_t *scales = b + 96;
        typename FT::raw *o = out + 128 * n + l0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int j = 8 * n + 2 * k + l0 / 16;
            const uint32_t lo4 = j < 8 ? (scales[j] & 15u) : (scales[j - 8] >> 4);
            const uint32_t hi2 = (scales[8 + (j & 3)] >> (2 * (j >> 2))) & 3u;
            const float dl = __fmul_rn(d, u2f_biased(lo4 | (hi2 << 4), 32.0f));
            const int bit = 4 * n + k;
            uint32_t c[NW];
            // ((q >> 2k) & 3) - (hbit ? 0 : 4)  ==  (two-bit | hbit << 2) - 4
#pragma unroll
            for (int i = 0; i < NW; i++) c[i] = ((w[i] >> (2 * k)) & 0x03030303u) | (((hm[i] >> bit) & 0x01010101u) << 2);
            emit_scaled<FT, SP>(o + 32 * k, c, 4.0f, dl, vec);
        }
        return f16_nonfinite(dbits);
    }
};

// What unit u of Decoder<T> writes: `nruns` runs of V elements, `stride` apart, from element s0 of its block; and which
// f16 fields of the block decide whether the unit must be re-evaluated exactly.
template <uint32_t T> struct UnitGeom;
#define GGQ_GEOM(T, S0, STRIDE, NRUNS, BAD)                                                                          \
    template <> struct UnitGeom<T> {                                                                                 \
        static __device__ __forceinline__ void runs(int u, int V, int &s0, int &stride, int &nruns) {               \
            const int CP = 32 / V; (void)CP;                                                                         \
            s0 = S0; stride = STRIDE; nruns = NRUNS;                                                                 \
        }                                                                                                            \
        static __device__ __forceinline__ bool nonfinite(const uint8_t *b) { return BAD; }                           \
    }
GGQ_GEOM(T_Q4_0, u * V, 16, 2, f16_nonfinite(lds16(b)));
GGQ_GEOM(T_Q4_1, u * V, 16, 2, f16_nonfinite(lds16(b)) || f16_nonfinite(lds16(b + 2)));
GGQ_GEOM(T_Q5_0, u * V, 16, 2, f16_nonfinite(lds16(b)));
GGQ_GEOM(T_Q5_1, u * V, 16, 2, f16_nonfinite(lds16(b)) || f16_nonfinite(lds16(b + 2)));
GGQ_GEOM(T_Q8_0, u * V, 0, 1, f16_nonfinite(lds16(b)));
GGQ_GEOM(T_Q8_1, u * V, 0, 1, f16_nonfinite(lds16(b)));
GGQ_GEOM(T_Q8K, u * V, 0, 1, f16_nonfinite(lds16(b)));
GGQ_GEOM(T_Q4K, 64 * (u / CP) + (u % CP) * V, 32, 2, f16_nonfinite(lds16(b)) || f16_nonfinite(lds16(b + 2)));
GGQ_GEOM(T_Q5K, 64 * (u / CP) + (u % CP) * V, 32, 2, f16_nonfinite(lds16(b)) || f16_nonfinite(lds16(b + 2)));
GGQ_GEOM(T_Q6K, 128 * (u / CP) + (u % CP) * V, 32, 4, f16_nonfinite(lds16(b + 208)));
GGQ_GEOM(T_Q2K, 128 * (u / CP) + (u % CP) * V, 32, 4, f16_nonfinite(lds16(b + 80)) || f16_nonfinite(lds16(b + 82)));
GGQ_GEOM(T_Q3K, 128 * (u / CP) + (u % CP) * V, 32, 4, f16_nonfinite(lds16(b + 108)));
#undef GGQ_GEOM

// Second pass over this thread's units of a tile, taken only when the fast pass met a block with a NaN / infinite scale
// field: those units are evaluated again, exactly (see above), over what the fast path wrote.  A thread rewrites only
// its own outputs, so no synchronisation is needed; out of line so the fast loop pays one OR per unit and nothing else.
template <uint32_t T, class FT, int THREADS>
__device__ __noinline__ void fix_units_exact(const uint8_t *stage, typename FT::raw *out, int nunits, int tid) {
    using TR = BlockTraits<T>;
    constexpr int UNITS = Decoder<T>::template units<FT::V>();
    for (int u = tid; u < nunits; u += THREADS) {
        const uint8_t *b = stage + (u / UNITS) * TR::BYTES;
        if (!UnitGeom<T>::nonfinite(b)) continue;
        int s0, stride, nruns;
        UnitGeom<T>::runs(u % UNITS, FT::V, s0, stride, nruns);
        decode_exact_runs<T, FT>(b, out + (size_t)(u / UNITS) * TR::ELEMS, s0, stride, nruns);
    }
}

// MODE 0: persistent CTAs, static round-robin over tiles, STAGES-deep bulk-copy (TMA) ring.
// MODE 1: one tile per CTA (grid = ntiles), the tile arrives by one bulk copy: the hardware block scheduler
//         overlaps the load of one CTA with the stores of its neighbours and balances the tail.
// MODE 2: one tile per CTA, staged by cooperative 16-byte loads (no mbarrier, no async proxy).
// MODE 3: STAGES consecutive tiles per short-lived CTA (grid = ceil(ntiles / STAGES)); all their bulk copies are
//         issued up front, so the CTA decodes tile k while tiles k+1.. are still in flight.
template <uint32_t T, class FT, int TILE_ELEMS, int STAGES, int THREADS, int MINB, int MODE, int SP>
__global__ void __launch_bounds__(THREADS, MINB)
dequant_kernel(const uint8_t *__restrict__ src, typename FT::raw *__restrict__ dst, size_t nblocks) {
    using TR = BlockTraits<T>;
    constexpr int TILE_BLOCKS = TILE_ELEMS / TR::ELEMS;
    constexpr int TILE_BYTES = TILE_BLOCKS * TR::BYTES;
    constexpr int UNITS = Decoder<T>::template units<FT::V>();
    constexpr int NST = (MODE == 0 || MODE == 3) ? STAGES : 1;
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

    if constexpr (MODE != 2) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < NST; s++) mbar_init(&bars[s], 1);
            fence_barrier_init();
        }
    }
    pdl_launch_dependents();  // the next kernel may begin scheduling as our CTAs retire
    if constexpr (MODE != 2) __syncthreads();
    pdl_wait();               // the previous kernel in the stream is complete before any global access

    auto tile_of = [&](size_t i) -> size_t {  // the i-th tile of this CTA
        return MODE == 3 ? (size_t)blockIdx.x * STAGES + i : blockIdx.x + i * (size_t)gridDim.x;
    };
    auto issue = [&](size_t i) {  // thread 0 only
        const size_t t = tile_of(i);
        if (MODE != 2 && t < full_tiles && src_fast) {
            const int s = (int)(i % NST);
            mbar_expect_tx(&bars[s], TILE_BYTES);
            bulk_g2s(stages + (size_t)s * TILE_BYTES, src + t * (size_t)TILE_BYTES, TILE_BYTES, &bars[s]);
        }
    };
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NST; i++) issue(i);
    }

    for (size_t i = 0; MODE != 3 || i < (size_t)STAGES; ++i) {
        const size_t t = tile_of(i);
        if (t >= ntiles) break;
        const int s = (int)(i % NST);
        uint8_t *stage = stages + (size_t)s * TILE_BYTES;
        const bool bulk = MODE != 2 && (t < full_tiles) && src_fast;
        const int nb = (t < full_tiles) ? TILE_BLOCKS : rem_blocks;
        if (bulk) {
            mbar_wait(&bars[s], (uint32_t)((i / NST) & 1));
        } else {
            cta_copy_g2s(stage, src + t * (size_t)TILE_BYTES, (uint32_t)nb * TR::BYTES, tid, THREADS);
            __syncthreads();
        }
        typename FT::raw *out = dst + t * (size_t)TILE_ELEMS;
        bool bad = false;  // some unit of this thread met a NaN / infinite scale field
        if (nb == TILE_BLOCKS) {
#pragma unroll 2
            for (int u = tid; u < TILE_BLOCKS * UNITS; u += THREADS)
                bad |= Decoder<T>::template run<FT, SP>(stage + (u / UNITS) * TR::BYTES, u % UNITS, out + (size_t)(u / UNITS) * TR::ELEMS, vec);
        } else {
            for (int u = tid; u < nb * UNITS; u += THREADS)
                bad |= Decoder<T>::template run<FT, SP>(stage + (u / UNITS) * TR::BYTES, u % UNITS, out + (size_t)(u / UNITS) * TR::ELEMS, vec);
        }
#ifndef GGQ_AB_NO_EXACT_PASS  /* A/B builds only (tools/gpu_round.sh ab_exact) */
        if (bad) fix_units_exact<T, FT, THREADS>(stage, out, nb * UNITS, tid);
#endif
        if constexpr (MODE == 0) {
            __syncthreads();  // stage s fully consumed
            if (tid == 0) issue(i + NST);
        } else if constexpr (MODE != 3) {
            if (t + gridDim.x < ntiles) {  // only when the grid was capped below ntiles
                __syncthreads();
                if (MODE == 1 && tid == 0) issue(i + 1);
            }
        }
    }
}

template <uint32_t T, int TILE_ELEMS, int STAGES, int MODE> constexpr int dequant_smem_bytes() {
    return 128 + ((MODE == 0 || MODE == 3) ? STAGES : 1) * (TILE_ELEMS / BlockTraits<T>::ELEMS) * BlockTraits<T>::BYTES + 16;
}

}  // namespace ggq
