/*
 * ggq_oracle.c — CPU ORACLE (test infrastructure only; see ggq_oracle.h for the parity status).
 *
 * Build: gcc -O3 -std=gnu11 -ffp-contract=off -fno-fast-math -fPIC -shared -pthread
 *        (-ffp-contract=off is REQUIRED: Rust never contracts a*b+c into an FMA.)
 *
 * Every function cites the reference file:line it restates (paths under /root/reference/).
 * All arithmetic is IEEE binary32, one rounding per written operator, evaluated left to right.
 *
 * The K-quant arithmetic (make_qkx2_quants, make_qx_quants, make_q3_quants, quantize / dequantize of Q2K..Q6K) has no
 * counterpart in the reference (`todo!()`); it restates upstream ggml's `ggml-quants.c` (llama.cpp / ggml,
 * Copyright (c) 2023-2024 The ggml authors, MIT License: "Permission is hereby granted, free of charge, to any person
 * obtaining a copy of this software and associated documentation files ... The above copyright notice and this permission
 * notice shall be included in all copies or substantial portions of the Software.") so that results are bit-identical
 * to what every GGUF consumer expects.  The loops follow upstream's order of operations by necessity.
 */
#include "ggq_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#define GGO_X86 1
#endif

/* ------------------------------------------------------------------------------------------ */
/* f32 <-> f16 / bf16: crate `half` 2.6.0 (Cargo.lock:243-246), not vendored in the reference. */
/* Published algorithm: IEEE-754 RNE narrow, overflow -> inf, NaN -> quiet NaN keeping the top   */
/* payload bits; widen is exact, NaN quieted.  Call sites: structs.rs:72-73,81,87;               */
/* structs/half.rs:16,20,32,36; lib.rs:68,72,83,88.                                              */
/* ------------------------------------------------------------------------------------------ */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

uint16_t ggo_f32_to_f16(float v) {
    uint32_t x = f2u(v);
    uint32_t sign = x & 0x80000000u, exp = x & 0x7F800000u, man = x & 0x007FFFFFu;
    if (exp == 0x7F800000u) {
        uint32_t nan_bit = man == 0 ? 0 : 0x0200u;
        return (uint16_t)((sign >> 16) | 0x7C00u | nan_bit | (man >> 13));
    }
    uint32_t hs = sign >> 16;
    int32_t he = (int32_t)(exp >> 23) - 127 + 15;
    if (he >= 0x1F) return (uint16_t)(hs | 0x7C00u);
    if (he <= 0) {
        if (14 - he > 24) return (uint16_t)hs;
        man |= 0x00800000u;
        uint32_t hm = man >> (14 - he);
        uint32_t rb = 1u << (13 - he);
        if ((man & rb) != 0 && (man & (3 * rb - 1)) != 0) hm += 1;
        return (uint16_t)(hs | hm);
    }
    uint32_t r = hs | ((uint32_t)he << 10) | (man >> 13);
    if ((man & 0x1000u) != 0 && (man & (3 * 0x1000u - 1)) != 0) r += 1;
    return (uint16_t)r;
}

float ggo_f16_to_f32(uint16_t h) {
    uint32_t i = h;
    if ((i & 0x7FFFu) == 0) return u2f(i << 16);
    uint32_t hs = i & 0x8000u, he = i & 0x7C00u, hm = i & 0x03FFu;
    if (he == 0x7C00u) {
        if (hm == 0) return u2f((hs << 16) | 0x7F800000u);
        return u2f((hs << 16) | 0x7FC00000u | (hm << 13));
    }
    uint32_t sign = hs << 16;
    int32_t ue = ((int32_t)he >> 10) - 15;
    if (he == 0) {
        /* subnormal: normalise */
        int e = 0;
        uint32_t m = hm;
        while ((m & 0x0400u) == 0) { m <<= 1; e++; }
        uint32_t exp = (uint32_t)(127 - 15 - e + 1) << 23;
        uint32_t man = (m & 0x03FFu) << 13;
        return u2f(sign | exp | man);
    }
    return u2f(sign | ((uint32_t)(ue + 127) << 23) | (hm << 13));
}

uint16_t ggo_f32_to_bf16(float v) {
    uint32_t x = f2u(v);
    if ((x & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((x >> 16) | 0x0040u);
    uint32_t rb = 0x8000u;
    if ((x & rb) != 0 && (x & (3 * rb - 1)) != 0) return (uint16_t)((x >> 16) + 1);
    return (uint16_t)(x >> 16);
}

float ggo_bf16_to_f32(uint16_t h) {
    uint32_t i = h;
    if ((i & 0x7FFFu) > 0x7F80u) return u2f((i | 0x0040u) << 16);
    return u2f(i << 16);
}

/* ------------------------------------------------------------------------------------------ */
/* Rust scalar semantics                                                                        */
/* ------------------------------------------------------------------------------------------ */
/* `v as u8`: saturating, NaN -> 0, truncation toward zero */
static inline uint8_t as_u8(float v) {
    if (!(v > 0.0f)) return 0; /* NaN, <= 0 */
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}
/* `v as i8`: saturating, NaN -> 0 */
static inline int8_t as_i8(float v) {
    if (v != v) return 0;
    if (v <= -128.0f) return -128;
    if (v >= 127.0f) return 127;
    return (int8_t)v;
}
/* f32::min(a, b): NaN operand dropped (IEEE minNum) */
static inline float rs_min(float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    return b < a ? b : a;
}
/* f32::round(): half away from zero */
static inline float rs_round(float v) { return roundf(v); }

/* structs.rs:91-94  max_abs: fold acc.max(|x|) from 0 (NaN ignored) */
static float max_abs(const float *x, int n) {
    float acc = 0.0f;
    for (int i = 0; i < n; i++) {
        float a = fabsf(x[i]);
        if (a > acc) acc = a;
    }
    return acc;
}
/* structs.rs:96-100  max_by_abs: first x with strictly larger |x| wins, sign kept */
static float max_by_abs(const float *x, int n) {
    float acc = 0.0f;
    for (int i = 0; i < n; i++)
        if (fabsf(x[i]) > fabsf(acc)) acc = x[i];
    return acc;
}
/* structs.rs:102-107  min_max: fold (min.min(x), max.max(x)) from (f32::MAX, f32::MIN).
 * NaN never replaces the accumulator.  Rust leaves min(+0,-0) unspecified; this oracle fixes it
 * the way rustc's x86-64 lowering (minss/maxss with the accumulator as the kept operand) does:
 * the accumulator is replaced only on a STRICT compare, i.e. the first-seen zero keeps its sign. */
static void min_max(const float *x, int n, float *mn, float *mx) {
    float lo = 3.40282347e+38f, hi = -3.40282347e+38f;
    for (int i = 0; i < n; i++) {
        if (x[i] < lo) lo = x[i];
        if (x[i] > hi) hi = x[i];
    }
    *mn = lo;
    *mx = hi;
}

static inline void put16(uint8_t *p, uint16_t v) { p[0] = (uint8_t)(v & 0xFF); p[1] = (uint8_t)(v >> 8); }
static inline uint16_t get16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

/* ------------------------------------------------------------------------------------------ */
/* Legacy 32-element blocks                                                                     */
/* ------------------------------------------------------------------------------------------ */

/* q4_0.rs:23-44 */
static void q4_0_quantize(const float *x, uint8_t *y) {
    float max = max_by_abs(x, 32);
    if (max == 0.0f) { memset(y, 0, 18); return; }
    float delta = max / -8.0f;
    float recip = 1.0f / delta;
    put16(y, ggo_f32_to_f16(delta));
    for (int i = 0; i < 16; i++) {
        uint8_t l = as_u8(rs_min(x[i] * recip + 8.5f, 15.0f));
        uint8_t h = as_u8(rs_min(x[i + 16] * recip + 8.5f, 15.0f));
        y[2 + i] = (uint8_t)((h << 4) | l);
    }
}
/* q4_0.rs:46-57 */
static void q4_0_dequantize(const uint8_t *y, float *x) {
    float delta = ggo_f16_to_f32(get16(y));
    for (int i = 0; i < 16; i++) {
        uint8_t b = y[2 + i];
        x[i] = (float)((int32_t)(b & 0xF) - 8) * delta;
        x[i + 16] = (float)((int32_t)(b >> 4) - 8) * delta;
    }
}

/* q4_1.rs:23-47 */
static void q4_1_quantize(const float *x, uint8_t *y) {
    float mn, mx;
    min_max(x, 32, &mn, &mx);
    if (mn == mx) {
        memset(y, 0, 20);
        put16(y + 2, ggo_f32_to_f16(mn));
        return;
    }
    float delta = (mx - mn) / 15.0f;
    float recip = 1.0f / delta;
    put16(y, ggo_f32_to_f16(delta));
    put16(y + 2, ggo_f32_to_f16(mn));
    for (int i = 0; i < 16; i++) {
        uint8_t l = as_u8((x[i] - mn) * recip + 0.5f);
        uint8_t h = as_u8((x[i + 16] - mn) * recip + 0.5f);
        if (l > 15) l = 15;
        if (h > 15) h = 15;
        y[4 + i] = (uint8_t)((h << 4) | l);
    }
}
/* q4_1.rs:49-60 */
static void q4_1_dequantize(const uint8_t *y, float *x) {
    float delta = ggo_f16_to_f32(get16(y)), mn = ggo_f16_to_f32(get16(y + 2));
    for (int i = 0; i < 16; i++) {
        uint8_t b = y[4 + i];
        x[i] = (float)(b & 0xF) * delta + mn;
        x[i + 16] = (float)(b >> 4) * delta + mn;
    }
}

/* shared qh/ql packing: q5_0.rs:43-51, q5_1.rs:46-54 */
static void pack5(const uint8_t *q, uint8_t *qh4, uint8_t *ql16) {
    uint32_t qh = 0;
    for (int i = 0; i < 16; i++) {
        uint8_t l = q[i], h = q[i + 16];
        qh |= (((uint32_t)l >> 4) & 1u) << i;
        qh |= (((uint32_t)h >> 4) & 1u) << (i + 16);
        ql16[i] = (uint8_t)(((h & 0xF) << 4) | (l & 0xF));
    }
    qh4[0] = (uint8_t)qh; qh4[1] = (uint8_t)(qh >> 8); qh4[2] = (uint8_t)(qh >> 16); qh4[3] = (uint8_t)(qh >> 24);
}
static inline uint32_t le32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* q5_0.rs:26-58 */
static void q5_0_quantize(const float *x, uint8_t *y) {
    float max = max_by_abs(x, 32);
    if (max == 0.0f) { memset(y, 0, 22); return; }
    float delta = max / -16.0f;
    float recip = 1.0f / delta;
    uint8_t q[32];
    for (int i = 0; i < 32; i++) {
        uint8_t v = as_u8(x[i] * recip + 16.5f);
        q[i] = v > 31 ? 31 : v;
    }
    put16(y, ggo_f32_to_f16(delta));
    pack5(q, y + 2, y + 6);
}
/* q5_0.rs:60-73 */
static void q5_0_dequantize(const uint8_t *y, float *x) {
    float delta = ggo_f16_to_f32(get16(y));
    uint32_t qh = le32(y + 2);
    for (int i = 0; i < 16; i++) {
        uint8_t b = y[6 + i];
        uint8_t lo = (uint8_t)((b & 0xF) | ((uint8_t)((qh >> i) << 4) & 0x10));
        uint8_t hi = (uint8_t)((b >> 4) | ((uint8_t)(qh >> (i + 12)) & 0x10));
        x[i] = (float)((int8_t)lo - 16) * delta;
        x[i + 16] = (float)((int8_t)hi - 16) * delta;
    }
}

/* q5_1.rs:26-62 */
static void q5_1_quantize(const float *x, uint8_t *y) {
    float mn, mx;
    min_max(x, 32, &mn, &mx);
    if (mn == mx) {
        memset(y, 0, 24);
        put16(y + 2, ggo_f32_to_f16(mn));
        return;
    }
    float delta = (mx - mn) / 31.0f;
    float recip = 1.0f / delta;
    uint8_t q[32];
    for (int i = 0; i < 32; i++) {
        uint8_t v = as_u8((x[i] - mn) * recip + 0.5f);
        q[i] = v > 31 ? 31 : v;
    }
    put16(y, ggo_f32_to_f16(delta));
    put16(y + 2, ggo_f32_to_f16(mn));
    pack5(q, y + 4, y + 8);
}
/* q5_1.rs:64-77 */
static void q5_1_dequantize(const uint8_t *y, float *x) {
    float delta = ggo_f16_to_f32(get16(y)), mn = ggo_f16_to_f32(get16(y + 2));
    uint32_t qh = le32(y + 4);
    for (int i = 0; i < 16; i++) {
        uint8_t b = y[8 + i];
        uint8_t lo = (uint8_t)((b & 0xF) | ((uint8_t)((qh >> i) << 4) & 0x10));
        uint8_t hi = (uint8_t)((b >> 4) | ((uint8_t)(qh >> (i + 12)) & 0x10));
        x[i] = (float)lo * delta + mn;
        x[i + 16] = (float)hi * delta + mn;
    }
}

/* q8_0.rs:23-41 */
static void q8_0_quantize(const float *x, uint8_t *y) {
    float amax = max_abs(x, 32);
    if (amax == 0.0f) { memset(y, 0, 34); return; }
    float delta = amax / 127.0f;
    float recip = 1.0f / delta;
    put16(y, ggo_f32_to_f16(delta));
    for (int i = 0; i < 32; i++) y[2 + i] = (uint8_t)as_i8(rs_round(x[i] * recip));
}
/* q8_0.rs:43-47 */
static void q8_0_dequantize(const uint8_t *y, float *x) {
    float delta = ggo_f16_to_f32(get16(y));
    for (int i = 0; i < 32; i++) x[i] = (float)(int8_t)y[2 + i] * delta;
}

/* q8_1.rs:28-55 */
static void q8_1_quantize(const float *x, uint8_t *y) {
    float amax = max_abs(x, 32);
    if (amax == 0.0f) { memset(y, 0, 36); return; }
    float delta = amax / 127.0f;
    float recip = 1.0f / delta;
    int16_t sum = 0;
    for (int i = 0; i < 32; i++) {
        int8_t q = as_i8(rs_round(x[i] * recip));
        y[4 + i] = (uint8_t)q;
        sum = (int16_t)(sum + q);
    }
    put16(y, ggo_f32_to_f16(delta));
    put16(y + 2, ggo_f32_to_f16((float)sum * delta)); /* unrounded f32 delta: q8_1.rs:52 */
}
/* q8_1.rs:57-61 */
static void q8_1_dequantize(const uint8_t *y, float *x) {
    float delta = ggo_f16_to_f32(get16(y));
    for (int i = 0; i < 32; i++) x[i] = (float)(int8_t)y[4 + i] * delta;
}

/* q8_k.rs:27-54 — reference layout: f16 delta, 290 bytes (q8_k.rs:7-15) */
static void q8_k_quantize(const float *x, uint8_t *y) {
    float max = max_by_abs(x, 256);
    if (max == 0.0f) { memset(y, 0, 290); return; }
    float delta = max / -127.0f;
    float recip = 1.0f / delta;
    int16_t sums[16];
    memset(sums, 0, sizeof sums);
    for (int i = 0; i < 256; i++) {
        int8_t q = as_i8(rs_min(rs_round(x[i] * recip), 127.0f));
        y[2 + i] = (uint8_t)q;
        sums[i / 16] = (int16_t)(sums[i / 16] + q);
    }
    put16(y, ggo_f32_to_f16(delta));
    for (int i = 0; i < 16; i++) put16(y + 258 + 2 * i, (uint16_t)sums[i]);
}
/* q8_k.rs:56-60 */
static void q8_k_dequantize(const uint8_t *y, float *x) {
    float delta = ggo_f16_to_f32(get16(y));
    for (int i = 0; i < 256; i++) x[i] = (float)(int8_t)y[2 + i] * delta;
}

/* ------------------------------------------------------------------------------------------ */
/* K-quants.  Layouts: structs/q{2,3,4,5,6}_k.rs.  Arithmetic: upstream ggml semantics          */
/* (`ggml/src/ggml-quants.c`, not a dependency of the reference; the reference has `todo!()`).  */
/* ------------------------------------------------------------------------------------------ */
#define GROUP_MAX_EPS 1e-15f

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* round-half-to-even via the 1.5*2^23 trick (valid for |v| <= 4194303) */
static inline int nearest_int(float v) {
    float t = v + 12582912.0f;
    int i;
    memcpy(&i, &t, 4);
    return (i & 0x007fffff) - 0x00400000;
}

/* 6-bit scale/min unpack for Q4K/Q5K */
static inline void get_scale_min_k4(int j, const uint8_t *q, uint8_t *d, uint8_t *m) {
    if (j < 4) {
        *d = q[j] & 63;
        *m = q[j + 4] & 63;
    } else {
        *d = (uint8_t)((q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4));
        *m = (uint8_t)((q[j + 4] >> 4) | ((q[j - 0] >> 6) << 4));
    }
}

static float make_qkx2_quants(int n, int nmax, const float *x, const float *weights, uint8_t *L,
                              float *the_min, uint8_t *Laux, float rmin, float rdelta, int nstep,
                              int use_mad) {
    float min = x[0];
    float max = x[0];
    float sum_w = weights[0];
    float sum_x = sum_w * x[0];
    for (int i = 1; i < n; ++i) {
        if (x[i] < min) min = x[i];
        if (x[i] > max) max = x[i];
        float w = weights[i];
        sum_w += w;
        sum_x += w * x[i];
    }
    if (min > 0) min = 0;
    if (max == min) {
        for (int i = 0; i < n; ++i) L[i] = 0;
        *the_min = -min;
        return 0.f;
    }
    float iscale = (float)nmax / (max - min);
    float scale = 1 / iscale;
    float best_mad = 0;
    for (int i = 0; i < n; ++i) {
        int l = nearest_int(iscale * (x[i] - min));
        L[i] = (uint8_t)imax(0, imin(nmax, l));
        float diff = scale * (float)L[i] + min - x[i];
        diff = use_mad ? fabsf(diff) : diff * diff;
        float w = weights[i];
        best_mad += w * diff;
    }
    if (nstep < 1) {
        *the_min = -min;
        return scale;
    }
    for (int is = 0; is <= nstep; ++is) {
        iscale = (rmin + rdelta * (float)is + (float)nmax) / (max - min);
        float sum_l = 0, sum_l2 = 0, sum_xl = 0;
        for (int i = 0; i < n; ++i) {
            int l = nearest_int(iscale * (x[i] - min));
            l = imax(0, imin(nmax, l));
            Laux[i] = (uint8_t)l;
            float w = weights[i];
            sum_l += w * (float)l;
            sum_l2 += w * (float)l * (float)l;
            sum_xl += w * (float)l * x[i];
        }
        float D = sum_w * sum_l2 - sum_l * sum_l;
        if (D > 0) {
            float this_scale = (sum_w * sum_xl - sum_x * sum_l) / D;
            float this_min = (sum_l2 * sum_x - sum_l * sum_xl) / D;
            if (this_min > 0) {
                this_min = 0;
                this_scale = sum_xl / sum_l2;
            }
            float mad = 0;
            for (int i = 0; i < n; ++i) {
                float diff = this_scale * (float)Laux[i] + this_min - x[i];
                diff = use_mad ? fabsf(diff) : diff * diff;
                float w = weights[i];
                mad += w * diff;
            }
            if (mad < best_mad) {
                for (int i = 0; i < n; ++i) L[i] = Laux[i];
                best_mad = mad;
                scale = this_scale;
                min = this_min;
            }
        }
    }
    *the_min = -min;
    return scale;
}

/* rmse_type == 1 (weights x^2), qw == NULL: the only form the K-quant reference rows use */
static float make_qx_quants(int n, int nmax, const float *x, int8_t *L) {
    float max = 0;
    float amax = 0;
    for (int i = 0; i < n; ++i) {
        float ax = fabsf(x[i]);
        if (ax > amax) { amax = ax; max = x[i]; }
    }
    if (amax < GROUP_MAX_EPS) {
        for (int i = 0; i < n; ++i) L[i] = 0;
        return 0.f;
    }
    float iscale = (float)(-nmax) / max;
    float sumlx = 0;
    float suml2 = 0;
    for (int i = 0; i < n; ++i) {
        int l = nearest_int(iscale * x[i]);
        l = imax(-nmax, imin(nmax - 1, l));
        L[i] = (int8_t)(l + nmax);
        float w = x[i] * x[i];
        sumlx += w * x[i] * (float)l;
        suml2 += w * (float)l * (float)l;
    }
    float scale = suml2 ? sumlx / suml2 : 0.0f;
    float best = scale * sumlx;
    for (int is = -9; is <= 9; ++is) {
        if (is == 0) continue;
        iscale = -((float)nmax + 0.1f * (float)is) / max;
        sumlx = suml2 = 0;
        for (int i = 0; i < n; ++i) {
            int l = nearest_int(iscale * x[i]);
            l = imax(-nmax, imin(nmax - 1, l));
            float w = x[i] * x[i];
            sumlx += w * x[i] * (float)l;
            suml2 += w * (float)l * (float)l;
        }
        if (suml2 > 0 && sumlx * sumlx > best * suml2) {
            for (int i = 0; i < n; ++i) {
                int l = nearest_int(iscale * x[i]);
                L[i] = (int8_t)(nmax + imax(-nmax, imin(nmax - 1, l)));
            }
            scale = sumlx / suml2;
            best = scale * sumlx;
        }
    }
    return scale;
}

/* do_rmse == true */
static float make_q3_quants(int n, int nmax, const float *x, int8_t *L) {
    float max = 0;
    float amax = 0;
    for (int i = 0; i < n; ++i) {
        float ax = fabsf(x[i]);
        if (ax > amax) { amax = ax; max = x[i]; }
    }
    if (amax < GROUP_MAX_EPS) {
        for (int i = 0; i < n; ++i) L[i] = 0;
        return 0.f;
    }
    float iscale = (float)(-nmax) / max;
    float sumlx = 0;
    float suml2 = 0;
    for (int i = 0; i < n; ++i) {
        int l = nearest_int(iscale * x[i]);
        l = imax(-nmax, imin(nmax - 1, l));
        L[i] = (int8_t)l;
        float w = x[i] * x[i];
        sumlx += w * x[i] * (float)l;
        suml2 += w * (float)l * (float)l;
    }
    for (int itry = 0; itry < 5; ++itry) {
        int n_changed = 0;
        for (int i = 0; i < n; ++i) {
            float w = x[i] * x[i];
            float slx = sumlx - w * x[i] * (float)L[i];
            if (slx > 0) {
                float sl2 = suml2 - w * (float)L[i] * (float)L[i];
                int new_l = nearest_int(x[i] * sl2 / slx);
                new_l = imax(-nmax, imin(nmax - 1, new_l));
                if (new_l != L[i]) {
                    slx += w * x[i] * (float)new_l;
                    sl2 += w * (float)new_l * (float)new_l;
                    if (sl2 > 0 && slx * slx * suml2 > sumlx * sumlx * sl2) {
                        L[i] = (int8_t)new_l;
                        sumlx = slx;
                        suml2 = sl2;
                        ++n_changed;
                    }
                }
            }
        }
        if (!n_changed) break;
    }
    for (int i = 0; i < n; ++i) L[i] = (int8_t)(L[i] + nmax);
    return sumlx / suml2;
}

/* ---- Q2K: layout q2_k.rs:5-13 {scales[16] @0, qs[64] @16, delta @80, min @82} ---- */
static void q2_k_quantize(const float *x, uint8_t *y) {
    uint8_t L[256], Laux[16];
    float weights[16], mins[16], scales[16];
    const float q4scale = 15.f;
    uint8_t *ysc = y, *yqs = y + 16;
    float max_scale = 0, max_min = 0;
    for (int j = 0; j < 16; ++j) {
        for (int l = 0; l < 16; ++l) weights[l] = fabsf(x[16 * j + l]);
        scales[j] = make_qkx2_quants(16, 3, x + 16 * j, weights, L + 16 * j, &mins[j], Laux, -0.5f, 0.1f, 15, 1);
        if (scales[j] > max_scale) max_scale = scales[j];
        if (mins[j] > max_min) max_min = mins[j];
    }
    uint16_t d16, dmin16;
    if (max_scale > 0) {
        float iscale = q4scale / max_scale;
        for (int j = 0; j < 16; ++j) ysc[j] = (uint8_t)nearest_int(iscale * scales[j]);
        d16 = ggo_f32_to_f16(max_scale / q4scale);
    } else {
        for (int j = 0; j < 16; ++j) ysc[j] = 0;
        d16 = ggo_f32_to_f16(0.f);
    }
    if (max_min > 0) {
        float iscale = q4scale / max_min;
        for (int j = 0; j < 16; ++j) ysc[j] |= (uint8_t)(nearest_int(iscale * mins[j]) << 4);
        dmin16 = ggo_f32_to_f16(max_min / q4scale);
    } else {
        dmin16 = ggo_f32_to_f16(0.f);
    }
    put16(y + 80, d16);
    put16(y + 82, dmin16);
    for (int j = 0; j < 16; ++j) {
        const float d = ggo_f16_to_f32(d16) * (float)(ysc[j] & 0xF);
        if (!d) continue;
        const float dm = ggo_f16_to_f32(dmin16) * (float)(ysc[j] >> 4);
        for (int ii = 0; ii < 16; ++ii) {
            int l = nearest_int((x[16 * j + ii] + dm) / d);
            L[16 * j + ii] = (uint8_t)imax(0, imin(3, l));
        }
    }
    for (int j = 0; j < 256; j += 128)
        for (int l = 0; l < 32; ++l)
            yqs[j / 4 + l] = (uint8_t)(L[j + l] | (L[j + l + 32] << 2) | (L[j + l + 64] << 4) | (L[j + l + 96] << 6));
}
static void q2_k_dequantize(const uint8_t *y, float *out) {
    const float d = ggo_f16_to_f32(get16(y + 80)), dmin = ggo_f16_to_f32(get16(y + 82));
    const uint8_t *sc = y, *q = y + 16;
    int is = 0;
    for (int n = 0; n < 256; n += 128) {
        int shift = 0;
        for (int j = 0; j < 4; ++j) {
            uint8_t s = sc[is++];
            float dl = d * (float)(s & 0xF), ml = dmin * (float)(s >> 4);
            for (int l = 0; l < 16; ++l) *out++ = dl * (float)((int8_t)((q[l] >> shift) & 3)) - ml;
            s = sc[is++];
            dl = d * (float)(s & 0xF);
            ml = dmin * (float)(s >> 4);
            for (int l = 0; l < 16; ++l) *out++ = dl * (float)((int8_t)((q[l + 16] >> shift) & 3)) - ml;
            shift += 2;
        }
        q += 32;
    }
}

/* ---- Q3K: layout q3_k.rs:5-15 {hmask[32] @0, qs[64] @32, scales[12] @96, delta @108} ---- */
static void q3_k_quantize(const float *x, uint8_t *y) {
    int8_t L[256];
    float scales[16];
    uint8_t *hmask = y, *qs = y + 32, *ysc = y + 96;
    float max_scale = 0, amax = 0;
    for (int j = 0; j < 16; ++j) {
        scales[j] = make_q3_quants(16, 4, x + 16 * j, L + 16 * j);
        float scale = fabsf(scales[j]);
        if (scale > amax) { amax = scale; max_scale = scales[j]; }
    }
    memset(ysc, 0, 12);
    uint16_t d16;
    if (max_scale) {
        float iscale = -32.f / max_scale;
        for (int j = 0; j < 16; ++j) {
            int8_t l = (int8_t)nearest_int(iscale * scales[j]);
            l = (int8_t)(imax(-32, imin(31, l)) + 32);
            if (j < 8) ysc[j] = (uint8_t)(l & 0xF);
            else ysc[j - 8] |= (uint8_t)((l & 0xF) << 4);
            l >>= 4;
            ysc[j % 4 + 8] |= (uint8_t)(l << (2 * (j / 4)));
        }
        d16 = ggo_f32_to_f16(1 / iscale);
    } else {
        d16 = ggo_f32_to_f16(0.f);
    }
    put16(y + 108, d16);
    for (int j = 0; j < 16; ++j) {
        int8_t sc = (int8_t)(j < 8 ? ysc[j] & 0xF : ysc[j - 8] >> 4);
        sc = (int8_t)((sc | (((ysc[8 + j % 4] >> (2 * (j / 4))) & 3) << 4)) - 32);
        float d = ggo_f16_to_f32(d16) * (float)sc;
        if (!d) continue;
        for (int ii = 0; ii < 16; ++ii) {
            int l = nearest_int(x[16 * j + ii] / d);
            l = imax(-4, imin(3, l));
            L[16 * j + ii] = (int8_t)(l + 4);
        }
    }
    memset(hmask, 0, 32);
    int m = 0;
    uint8_t hm = 1;
    for (int j = 0; j < 256; ++j) {
        if (L[j] > 3) { hmask[m] |= hm; L[j] = (int8_t)(L[j] - 4); }
        if (++m == 32) { m = 0; hm = (uint8_t)(hm << 1); }
    }
    for (int j = 0; j < 256; j += 128)
        for (int l = 0; l < 32; ++l)
            qs[j / 4 + l] = (uint8_t)(L[j + l] | (L[j + l + 32] << 2) | (L[j + l + 64] << 4) | (L[j + l + 96] << 6));
}
static void q3_k_dequantize(const uint8_t *y, float *out) {
    const uint32_t kmask1 = 0x03030303, kmask2 = 0x0f0f0f0f;
    const float d_all = ggo_f16_to_f32(get16(y + 108));
    const uint8_t *q = y + 32, *hm = y;
    uint8_t m = 1;
    uint32_t aux[4];
    memcpy(aux, y + 96, 12);
    uint32_t tmp = aux[2];
    aux[2] = ((aux[0] >> 4) & kmask2) | (((tmp >> 4) & kmask1) << 4);
    aux[3] = ((aux[1] >> 4) & kmask2) | (((tmp >> 6) & kmask1) << 4);
    aux[0] = (aux[0] & kmask2) | (((tmp >> 0) & kmask1) << 4);
    aux[1] = (aux[1] & kmask2) | (((tmp >> 2) & kmask1) << 4);
    const int8_t *scales = (const int8_t *)aux;
    int is = 0;
    for (int n = 0; n < 256; n += 128) {
        int shift = 0;
        for (int j = 0; j < 4; ++j) {
            float dl = d_all * (float)(scales[is++] - 32);
            for (int l = 0; l < 16; ++l)
                *out++ = dl * (float)((int8_t)((q[l + 0] >> shift) & 3) - ((hm[l + 0] & m) ? 0 : 4));
            dl = d_all * (float)(scales[is++] - 32);
            for (int l = 0; l < 16; ++l)
                *out++ = dl * (float)((int8_t)((q[l + 16] >> shift) & 3) - ((hm[l + 16] & m) ? 0 : 4));
            shift += 2;
            m = (uint8_t)(m << 1);
        }
        q += 32;
    }
}

/* shared by Q4K / Q5K: 6-bit packing of 8 (scale, min) pairs into 12 bytes */
static void pack_scales_k4(const float *scales, const float *mins, float inv_scale, float inv_min, uint8_t *ysc) {
    for (int j = 0; j < 8; ++j) {
        uint8_t ls = (uint8_t)nearest_int(inv_scale * scales[j]);
        uint8_t lm = (uint8_t)nearest_int(inv_min * mins[j]);
        ls = ls < 63 ? ls : 63;
        lm = lm < 63 ? lm : 63;
        if (j < 4) {
            ysc[j] = ls;
            ysc[j + 4] = lm;
        } else {
            ysc[j + 4] = (uint8_t)((ls & 0xF) | ((lm & 0xF) << 4));
            ysc[j - 4] |= (uint8_t)((ls >> 4) << 6);
            ysc[j - 0] |= (uint8_t)((lm >> 4) << 6);
        }
    }
}

/* ---- Q4K: layout q4_k.rs:5-13 {delta @0, min @2, scales[12] @4, qs[128] @16} ---- */
static void q4_k_quantize(const float *x, uint8_t *y) {
    uint8_t L[256], Laux[32];
    float weights[32], mins[8], scales[8];
    float max_scale = 0, max_min = 0;
    for (int j = 0; j < 8; ++j) {
        float sum_x2 = 0;
        for (int l = 0; l < 32; ++l) sum_x2 += x[32 * j + l] * x[32 * j + l];
        float av_x = sqrtf(sum_x2 / 32);
        for (int l = 0; l < 32; ++l) weights[l] = av_x + fabsf(x[32 * j + l]);
        scales[j] = make_qkx2_quants(32, 15, x + 32 * j, weights, L + 32 * j, &mins[j], Laux, -1.f, 0.1f, 20, 0);
        if (scales[j] > max_scale) max_scale = scales[j];
        if (mins[j] > max_min) max_min = mins[j];
    }
    float inv_scale = max_scale > 0 ? 63.f / max_scale : 0.f;
    float inv_min = max_min > 0 ? 63.f / max_min : 0.f;
    uint8_t *ysc = y + 4;
    pack_scales_k4(scales, mins, inv_scale, inv_min, ysc);
    uint16_t d16 = ggo_f32_to_f16(max_scale / 63.f), dmin16 = ggo_f32_to_f16(max_min / 63.f);
    put16(y, d16);
    put16(y + 2, dmin16);
    for (int j = 0; j < 8; ++j) {
        uint8_t sc, m;
        get_scale_min_k4(j, ysc, &sc, &m);
        const float d = ggo_f16_to_f32(d16) * (float)sc;
        if (!d) continue;
        const float dm = ggo_f16_to_f32(dmin16) * (float)m;
        for (int ii = 0; ii < 32; ++ii) {
            int l = nearest_int((x[32 * j + ii] + dm) / d);
            L[32 * j + ii] = (uint8_t)imax(0, imin(15, l));
        }
    }
    uint8_t *q = y + 16;
    for (int j = 0; j < 256; j += 64) {
        for (int l = 0; l < 32; ++l) q[l] = (uint8_t)(L[j + l] | (L[j + l + 32] << 4));
        q += 32;
    }
}
static void q4_k_dequantize(const uint8_t *y, float *out) {
    const float d = ggo_f16_to_f32(get16(y)), min = ggo_f16_to_f32(get16(y + 2));
    const uint8_t *q = y + 16;
    int is = 0;
    uint8_t sc, m;
    for (int j = 0; j < 256; j += 64) {
        get_scale_min_k4(is + 0, y + 4, &sc, &m);
        const float d1 = d * (float)sc, m1 = min * (float)m;
        get_scale_min_k4(is + 1, y + 4, &sc, &m);
        const float d2 = d * (float)sc, m2 = min * (float)m;
        for (int l = 0; l < 32; ++l) *out++ = d1 * (float)(q[l] & 0xF) - m1;
        for (int l = 0; l < 32; ++l) *out++ = d2 * (float)(q[l] >> 4) - m2;
        q += 32;
        is += 2;
    }
}

/* ---- Q5K: layout q5_k.rs:6-18 {delta @0, min @2, scales[12] @4, qh[32] @16, qs[128] @48} ---- */
static void q5_k_quantize(const float *x, uint8_t *y) {
    uint8_t L[256], Laux[32];
    float weights[32], mins[8], scales[8];
    float max_scale = 0, max_min = 0;
    for (int j = 0; j < 8; ++j) {
        float sum_x2 = 0;
        for (int l = 0; l < 32; ++l) sum_x2 += x[32 * j + l] * x[32 * j + l];
        float av_x = sqrtf(sum_x2 / 32);
        for (int l = 0; l < 32; ++l) weights[l] = av_x + fabsf(x[32 * j + l]);
        scales[j] = make_qkx2_quants(32, 31, x + 32 * j, weights, L + 32 * j, &mins[j], Laux, -0.5f, 0.1f, 15, 0);
        if (scales[j] > max_scale) max_scale = scales[j];
        if (mins[j] > max_min) max_min = mins[j];
    }
    float inv_scale = max_scale > 0 ? 63.f / max_scale : 0.f;
    float inv_min = max_min > 0 ? 63.f / max_min : 0.f;
    uint8_t *ysc = y + 4;
    pack_scales_k4(scales, mins, inv_scale, inv_min, ysc);
    uint16_t d16 = ggo_f32_to_f16(max_scale / 63.f), dmin16 = ggo_f32_to_f16(max_min / 63.f);
    put16(y, d16);
    put16(y + 2, dmin16);
    for (int j = 0; j < 8; ++j) {
        uint8_t sc, m;
        get_scale_min_k4(j, ysc, &sc, &m);
        const float d = ggo_f16_to_f32(d16) * (float)sc;
        if (!d) continue;
        const float dm = ggo_f16_to_f32(dmin16) * (float)m;
        for (int ii = 0; ii < 32; ++ii) {
            int l = nearest_int((x[32 * j + ii] + dm) / d);
            L[32 * j + ii] = (uint8_t)imax(0, imin(31, l));
        }
    }
    uint8_t *qh = y + 16, *ql = y + 48;
    memset(qh, 0, 32);
    uint8_t m1 = 1, m2 = 2;
    for (int n = 0; n < 256; n += 64) {
        for (int j = 0; j < 32; ++j) {
            int l1 = L[n + j];
            if (l1 > 15) { l1 -= 16; qh[j] |= m1; }
            int l2 = L[n + j + 32];
            if (l2 > 15) { l2 -= 16; qh[j] |= m2; }
            ql[j] = (uint8_t)(l1 | (l2 << 4));
        }
        m1 = (uint8_t)(m1 << 2);
        m2 = (uint8_t)(m2 << 2);
        ql += 32;
    }
}
static void q5_k_dequantize(const uint8_t *y, float *out) {
    const float d = ggo_f16_to_f32(get16(y)), min = ggo_f16_to_f32(get16(y + 2));
    const uint8_t *ql = y + 48, *qh = y + 16;
    int is = 0;
    uint8_t sc, m, u1 = 1, u2 = 2;
    for (int j = 0; j < 256; j += 64) {
        get_scale_min_k4(is + 0, y + 4, &sc, &m);
        const float d1 = d * (float)sc, m1 = min * (float)m;
        get_scale_min_k4(is + 1, y + 4, &sc, &m);
        const float d2 = d * (float)sc, m2 = min * (float)m;
        for (int l = 0; l < 32; ++l) *out++ = d1 * (float)((ql[l] & 0xF) + (qh[l] & u1 ? 16 : 0)) - m1;
        for (int l = 0; l < 32; ++l) *out++ = d2 * (float)((ql[l] >> 4) + (qh[l] & u2 ? 16 : 0)) - m2;
        ql += 32;
        is += 2;
        u1 = (uint8_t)(u1 << 2);
        u2 = (uint8_t)(u2 << 2);
    }
}

/* ---- Q6K: layout q6_k.rs:6-16 {ql[128] @0, qh[64] @128, scales[16] (i8) @192, delta @208} ---- */
static void q6_k_quantize(const float *x, uint8_t *y) {
    int8_t L[256];
    float scales[16];
    float max_scale = 0, max_abs_scale = 0;
    for (int ib = 0; ib < 16; ++ib) {
        const float scale = make_qx_quants(16, 32, x + 16 * ib, L + 16 * ib);
        scales[ib] = scale;
        const float abs_scale = fabsf(scale);
        if (abs_scale > max_abs_scale) { max_abs_scale = abs_scale; max_scale = scale; }
    }
    if (max_abs_scale < GROUP_MAX_EPS) {
        memset(y, 0, 210);
        return;
    }
    float iscale = -128.f / max_scale;
    uint16_t d16 = ggo_f32_to_f16(1 / iscale);
    put16(y + 208, d16);
    int8_t *ysc = (int8_t *)(y + 192);
    for (int ib = 0; ib < 16; ++ib) ysc[ib] = (int8_t)imin(127, nearest_int(iscale * scales[ib]));
    for (int j = 0; j < 16; ++j) {
        float d = ggo_f16_to_f32(d16) * (float)ysc[j];
        if (!d) continue;
        for (int ii = 0; ii < 16; ++ii) {
            int l = nearest_int(x[16 * j + ii] / d);
            l = imax(-32, imin(31, l));
            L[16 * j + ii] = (int8_t)(l + 32);
        }
    }
    uint8_t *ql = y, *qh = y + 128;
    for (int j = 0; j < 256; j += 128) {
        for (int l = 0; l < 32; ++l) {
            const uint8_t q1 = L[j + l + 0] & 0xF, q2 = L[j + l + 32] & 0xF;
            const uint8_t q3 = L[j + l + 64] & 0xF, q4 = L[j + l + 96] & 0xF;
            ql[l + 0] = (uint8_t)(q1 | (q3 << 4));
            ql[l + 32] = (uint8_t)(q2 | (q4 << 4));
            qh[l] = (uint8_t)((L[j + l] >> 4) | ((L[j + l + 32] >> 4) << 2) | ((L[j + l + 64] >> 4) << 4) | ((L[j + l + 96] >> 4) << 6));
        }
        ql += 64;
        qh += 32;
    }
}
static void q6_k_dequantize(const uint8_t *y, float *out) {
    const float d = ggo_f16_to_f32(get16(y + 208));
    const uint8_t *ql = y, *qh = y + 128;
    const int8_t *sc = (const int8_t *)(y + 192);
    for (int n = 0; n < 256; n += 128) {
        for (int l = 0; l < 32; ++l) {
            int is = l / 16;
            const int8_t q1 = (int8_t)((ql[l + 0] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
            const int8_t q2 = (int8_t)((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
            const int8_t q3 = (int8_t)((ql[l + 0] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
            const int8_t q4 = (int8_t)((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
            out[l + 0] = d * (float)sc[is + 0] * (float)q1;
            out[l + 32] = d * (float)sc[is + 2] * (float)q2;
            out[l + 64] = d * (float)sc[is + 4] * (float)q3;
            out[l + 96] = d * (float)sc[is + 6] * (float)q4;
        }
        out += 128;
        ql += 64;
        qh += 32;
        sc += 8;
    }
}

/* ---- f16 / bf16 as 1-element blocks: structs/half.rs:8-38 ---- */
static void f16_quantize(const float *x, uint8_t *y) { put16(y, ggo_f32_to_f16(x[0])); }
static void f16_dequantize(const uint8_t *y, float *x) { x[0] = ggo_f16_to_f32(get16(y)); }
static void bf16_quantize(const float *x, uint8_t *y) { put16(y, ggo_f32_to_bf16(x[0])); }
static void bf16_dequantize(const uint8_t *y, float *x) { x[0] = ggo_bf16_to_f32(get16(y)); }

/* ------------------------------------------------------------------------------------------ */
/* type table + slice drivers (lib.rs:62-90 f16/bf16 mediation, lib.rs:116-148 drivers)         */
/* ------------------------------------------------------------------------------------------ */
typedef void (*quant_fn)(const float *, uint8_t *);
typedef void (*dequant_fn)(const uint8_t *, float *);
typedef struct { uint32_t type, elems, bytes; quant_fn q; dequant_fn dq; } type_info;

static const type_info TYPES[] = {
    {GGO_F16, 1, 2, f16_quantize, f16_dequantize},
    {GGO_BF16, 1, 2, bf16_quantize, bf16_dequantize},
    {GGO_Q4_0, 32, 18, q4_0_quantize, q4_0_dequantize},
    {GGO_Q4_1, 32, 20, q4_1_quantize, q4_1_dequantize},
    {GGO_Q5_0, 32, 22, q5_0_quantize, q5_0_dequantize},
    {GGO_Q5_1, 32, 24, q5_1_quantize, q5_1_dequantize},
    {GGO_Q8_0, 32, 34, q8_0_quantize, q8_0_dequantize},
    {GGO_Q8_1, 32, 36, q8_1_quantize, q8_1_dequantize},
    {GGO_Q2K, 256, 84, q2_k_quantize, q2_k_dequantize},
    {GGO_Q3K, 256, 110, q3_k_quantize, q3_k_dequantize},
    {GGO_Q4K, 256, 144, q4_k_quantize, q4_k_dequantize},
    {GGO_Q5K, 256, 176, q5_k_quantize, q5_k_dequantize},
    {GGO_Q6K, 256, 210, q6_k_quantize, q6_k_dequantize},
    {GGO_Q8K, 256, 290, q8_k_quantize, q8_k_dequantize},
};

static const type_info *find_type(uint32_t t) {
    for (size_t i = 0; i < sizeof TYPES / sizeof TYPES[0]; i++)
        if (TYPES[i].type == t) return &TYPES[i];
    return NULL;
}
static int fdt_size(uint32_t fdt) { return fdt == GGO_F32 ? 4 : (fdt == GGO_F16 || fdt == GGO_BF16) ? 2 : 0; }

int ggo_block_info(uint32_t type, uint32_t *elems, uint32_t *bytes) {
    const type_info *ti = find_type(type);
    if (!ti) return GGO_UNSUPPORTED;
    if (elems) *elems = ti->elems;
    if (bytes) *bytes = ti->bytes;
    return GGO_OK;
}

typedef struct {
    const type_info *ti;
    uint32_t fdt;
    int quant;
    uint8_t *blocks;      /* packed side */
    uint8_t *floats;      /* float side  */
    size_t b0, b1;        /* block range */
} job;

/* f16 <-> f32 over a whole block with the F16C instructions when the CPU has them — what the `half`
 * crate does at run time (std feature detection).  VCVTPH2PS / VCVTPS2PH(RNE) give the same bits as
 * the software routines above for every input, NaNs included (quieted, payload kept); checked by
 * tests/test_oracle.py::test_f16c_mediation_equals_software. */
#ifdef GGO_X86
__attribute__((target("f16c,avx"))) static void widen_f16_hw(const uint8_t *src, float *dst, uint32_t n) {
    uint32_t i = 0;
    for (; i + 8 <= n; i += 8) _mm256_storeu_ps(dst + i, _mm256_cvtph_ps(_mm_loadu_si128((const __m128i *)(src + 2 * i))));
    for (; i < n; i++) dst[i] = ggo_f16_to_f32(get16(src + 2 * i));
}
__attribute__((target("f16c,avx"))) static void narrow_f16_hw(const float *src, uint8_t *dst, uint32_t n) {
    uint32_t i = 0;
    for (; i + 8 <= n; i += 8)
        _mm_storeu_si128((__m128i *)(dst + 2 * i), _mm256_cvtps_ph(_mm256_loadu_ps(src + i), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
    for (; i < n; i++) put16(dst + 2 * i, ggo_f32_to_f16(src[i]));
}
static int have_f16c(void) {
    static int v = -1;
    if (v < 0) v = __builtin_cpu_supports("f16c") && __builtin_cpu_supports("avx");
    return v;
}
#else
static int have_f16c(void) { return 0; }
static void widen_f16_hw(const uint8_t *s, float *d, uint32_t n) { (void)s; (void)d; (void)n; }
static void narrow_f16_hw(const float *s, uint8_t *d, uint32_t n) { (void)s; (void)d; (void)n; }
#endif

static void run_range(const job *j) {
    const type_info *ti = j->ti;
    const uint32_t n = ti->elems;
    const int fs = fdt_size(j->fdt);
    const int hw = have_f16c();
    float buf[256];
    for (size_t b = j->b0; b < j->b1; b++) {
        uint8_t *blk = j->blocks + b * ti->bytes;
        uint8_t *fl = j->floats + b * n * (size_t)fs;
        if (j->quant) {
            /* lib.rs:66-69, 82-84: widen every element to f32 first */
            if (j->fdt == GGO_F32) memcpy(buf, fl, n * 4u);
            else if (j->fdt == GGO_F16 && hw && n >= 8) widen_f16_hw(fl, buf, n);
            else if (j->fdt == GGO_F16) for (uint32_t i = 0; i < n; i++) buf[i] = ggo_f16_to_f32(get16(fl + 2 * i));
            else for (uint32_t i = 0; i < n; i++) buf[i] = ggo_bf16_to_f32(get16(fl + 2 * i));
            ti->q(buf, blk);
        } else {
            /* lib.rs:70-73, 87-89: f32 result, then RNE narrow */
            ti->dq(blk, buf);
            if (j->fdt == GGO_F32) memcpy(fl, buf, n * 4u);
            else if (j->fdt == GGO_F16 && hw && n >= 8) narrow_f16_hw(buf, fl, n);
            else if (j->fdt == GGO_F16) for (uint32_t i = 0; i < n; i++) put16(fl + 2 * i, ggo_f32_to_f16(buf[i]));
            else for (uint32_t i = 0; i < n; i++) put16(fl + 2 * i, ggo_f32_to_bf16(buf[i]));
        }
    }
}
static void *thread_main(void *p) { run_range((const job *)p); return NULL; }

static void run_parallel(job base, size_t nblocks, int threads) {
    if (threads <= 1 || nblocks < (size_t)threads * 4) {
        base.b0 = 0; base.b1 = nblocks;
        run_range(&base);
        return;
    }
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    job *jobs = (job *)malloc(sizeof(job) * (size_t)threads);
    for (int t = 0; t < threads; t++) {
        jobs[t] = base;
        jobs[t].b0 = nblocks * (size_t)t / (size_t)threads;
        jobs[t].b1 = nblocks * (size_t)(t + 1) / (size_t)threads;
        pthread_create(&tid[t], NULL, thread_main, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    free(tid);
    free(jobs);
}

int ggo_quantize_slice(uint32_t type, uint32_t fdt, void *dst, size_t dst_blocks, const void *src,
                       size_t src_elems, int threads) {
    const type_info *ti = find_type(type);
    if (!ti || !fdt_size(fdt)) return GGO_UNSUPPORTED;
    if (src_elems % ti->elems != 0) return GGO_INDIVISIBLE;        /* lib.rs:122-124 */
    if (dst_blocks != src_elems / ti->elems) return GGO_LENGTH_MISMATCH; /* lib.rs:125-127 */
    job j = {ti, fdt, 1, (uint8_t *)dst, (uint8_t *)(uintptr_t)src, 0, 0};
    run_parallel(j, dst_blocks, threads);
    return GGO_OK;
}

int ggo_dequantize_slice(uint32_t type, uint32_t fdt, void *dst, size_t dst_elems, const void *src,
                         size_t src_blocks, int threads) {
    const type_info *ti = find_type(type);
    if (!ti || !fdt_size(fdt)) return GGO_UNSUPPORTED;
    if (dst_elems % ti->elems != 0) return GGO_INDIVISIBLE;        /* lib.rs:136-138 */
    if (src_blocks != dst_elems / ti->elems) return GGO_LENGTH_MISMATCH; /* lib.rs:139-141 */
    job j = {ti, fdt, 0, (uint8_t *)(uintptr_t)src, (uint8_t *)dst, 0, 0};
    run_parallel(j, src_blocks, threads);
    return GGO_OK;
}
