// quant_k_kernel.cuh — f32 / f16 / bf16 -> Q2K Q3K Q4K Q5K Q6K (256-element super-blocks), sm_100a: the quantizers and the
// kernel template.  Kept in a header so tools/kq_sweep.cu can instantiate other configurations of exactly the code that
// ships (quant_k.cu holds the shipped configurations and the launchers).
// Compiled with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false: every expression below is
// evaluated exactly as written, one IEEE rounding per operator, left to right — the same way the
// CPU oracle (oracle/ggq_oracle.c, built with -ffp-contract=off) evaluates it.
//
// The reference only declares these layouts (/root/reference/ggml-quants/src/structs/q{2..6}_k.rs;
// quantize is `todo!()`), so the arithmetic is upstream ggml's `quantize_row_qN_K_ref`, whose scale
// search is *sequentially dependent across candidates* (an accepted candidate changes `min` for the
// next one).  It therefore cannot be split across lanes without changing results.  Mapping used:
//   * ONE LANE PER SUB-BLOCK (32 elements for Q4K/Q5K, 16 for Q2K/Q3K/Q6K): a warp owns 4 (resp. 2)
//     consecutive super-blocks; each lane keeps its sub-block in registers and runs the search in
//     the oracle's exact order — no redundant work, no cross-lane float reduction;
//   * the 8/16 per-sub-block results of a super-block are combined by xor-shuffles (max is
//     order-independent; "first with greatest |.|" uses a lowest-lane-wins arg-max);
//   * codes go through a per-warp shared-memory scratch and are assembled into the packed layout one 32-bit
//     (Q2K / Q4K / Q5K) or 16-bit (Q3K / Q6K) unit per lane and iteration, then copied to global memory;
//   * the launcher decides who hands out the warp passes (quant_k.cu): one pass per warp and the hardware's CTA
//     scheduler (Q2K / Q4K / Q5K / Q6K), or a persistent grid whose passes are tickets of a launch-wide counter (Q3K);
//     a warp that takes more than one pass has the rows of its next pass arrive by cp.async while this pass searches.
// This path is COMPUTE-bound (~400 dependent flops per element for Q4K), not HBM-bound; see DESIGN.md.
#pragma once
#include <type_traits>

#include "ggq_common.cuh"
#include "ggq_kernels.h"

namespace ggq {

constexpr unsigned KFULL = 0xFFFFFFFFu;
#define GROUP_MAX_EPS 1e-15f

__device__ __forceinline__ int nearest_int(float v) {
    const float t = v + 12582912.f;
    return (__float_as_int(t) & 0x007fffff) - 0x00400000;
}
// The same for an argument that can be inf * 0: super-blocks of subnormal weights overflow `63 / max_scale`, and a
// sub-block whose scale or min is exactly 0 then feeds a NaN to upstream's nearest_int().  On the reference platform that
// NaN is the default one (0xFFC00000: mantissa 0x400000), which the magic-number arithmetic turns into 0; the GPU's
// arithmetic NaN (0x7FFFFFFF) would come out as 4194303.  Infinities and huge values take the ordinary route (same bits
// on both machines).
__device__ __forceinline__ int nearest_int_of_product(float v) { return v != v ? 0 : nearest_int(v); }
// exact (float)l for a small integer l in [-64, 2^22)
__device__ __forceinline__ float i2f_small(int l) { return u2f_biased((uint32_t)(l + 64), 64.0f); }

template <int NW> __device__ __forceinline__ void set_code(uint32_t (&w)[NW], int i, int l) { w[i >> 2] |= (uint32_t)(l & 0xFF) << (8 * (i & 3)); }

// Rounding without leaving the FP pipes.  On upstream's domain (|v| <= 4194303, its assert)
//   clamp(nearest_int(v), lo, hi) == rint(clamp(v, lo, hi))            (rint is monotone, lo/hi integers)
// and rint(c) for |c| < 2^22 is (c + 1.5*2^23) - 1.5*2^23 in round-to-nearest-even, which is the very
// addition upstream's bit trick performs.  `rb` keeps the sum: its low mantissa byte is the code.
constexpr float RMAGIC = 12582912.f;
__device__ __forceinline__ float round_clamped(float v, float lo, float hi, float &rb) {
    rb = fminf(fmaxf(v, lo), hi) + RMAGIC;
    return rb - RMAGIC;
}
// byte k of `word` <- low byte of `bits`
__device__ __forceinline__ uint32_t put_byte(uint32_t word, uint32_t bits, int k) {
    return __byte_perm(word, bits, k == 0 ? 0x3214u : k == 1 ? 0x3240u : k == 2 ? 0x3410u : 0x4210u);
}
// exact float(byte k of word) for small unsigned codes
__device__ __forceinline__ float byte_as_float(uint32_t word, int k) {
    return __uint_as_float(__byte_perm(word, 0x4B400000u, 0x7650u + k)) - RMAGIC;
}

// ---- n / d for the 16 or 32 numerators of a sub-block that share one divisor ------------------------------------
// Upstream requantizes with `nearest_int((x + dm) / d)`: an IEEE division per element, ~18 instructions each the way
// the compiler expands it (reciprocal, two refinements, residual, range check).  With one divisor per sub-block the
// reciprocal is computed once, correctly rounded (`rcp.rn`), and Markstein's sequence gives the correctly rounded
// quotient in three operations:  q0 = RN(n r);  e = n - d q0 (exact, one FMA);  q = RN(q0 + e r)  —  the same value
// `div.rn` returns, PROVIDED nothing under- or overflows on the way: d is taken in 2^-40 .. 2^40 and |n| <= 2^60; if
// the exact residual is too small to be represented then |q0| < 2^-63 and the caller's nearest_int() is 0 for any
// last-bit error.  Everything else takes `div.rn`.
struct SharedDivisor {
    float d, r;
    bool ok;
};
__device__ __forceinline__ SharedDivisor shared_divisor(float d) {
    SharedDivisor sd;
    const float a = fabsf(d);
    sd.d = d;
    sd.ok = a >= 9.094947017729282e-13f && a <= 1.099511627776e12f;  // 2^-40 .. 2^40 (false for NaN)
    sd.r = __frcp_rn(d);
    return sd;
}
__device__ __forceinline__ float div_shared(float n, const SharedDivisor &sd) {
    if (sd.ok && fabsf(n) <= 1.152921504606846976e18f) {  // 2^60
        const float q0 = __fmul_rn(n, sd.r);
        const float e = __fmaf_rn(-sd.d, q0, n);
        return __fmaf_rn(e, sd.r, q0);
    }
    return __fdiv_rn(n, sd.d);
}

// The same sequence where the quotient itself is used (search scales, not just nearest_int of it): additionally
// 2^-60 <= |n|, so the residual e (a multiple of 2^(exponent(n) - 47)) and the quotient (>= 2^-100) are ordinary numbers
// and q is RN(n / d) without exception (host test: tests/cpp/test_shared_divisor.c, strict part).  n = 0, tiny, huge or
// NaN numerators and out-of-range divisors take div.rn.  SD 0 disables it (A/B in tools/kq_sweep.cu).
template <int SD> __device__ __forceinline__ float div_shared_strict(float n, const SharedDivisor &sd) {
    const float an = fabsf(n);
    if (SD && sd.ok && an >= 8.673617379884035e-19f && an <= 1.152921504606846976e18f) {  // 2^-60 .. 2^60
        const float q0 = __fmul_rn(n, sd.r);
        const float e = __fmaf_rn(-sd.d, q0, n);
        return __fmaf_rn(e, sd.r, q0);
    }
    return __fdiv_rn(n, sd.d);
}

// ---- upstream make_qkx2_quants ------------------------------------------------------------------
// Upstream keeps the codes L of the best candidate.  Every caller then REQUANTIZES the sub-block
// with the 6-/4-bit rounded scale and min and uses the search's L only when that rounded scale is 0,
// so the search here returns what produced its best codes — (iscale, min) of the accepted candidate —
// and qkx2_codes() re-evaluates upstream's expression  nearest_int(iscale * (x[i] - min))  from those
// in the rare case they are needed: same operands, same operation, same bits.  The candidate's codes
// live in registers as floats for the error pass and are never packed.
template <int N>
__device__ __forceinline__ void qkx2_codes(const float (&x)[N], const float isc, const float mn, const int nmax, uint32_t (&L)[N / 4]) {
    const float fmax_l = (float)nmax;
#pragma unroll
    for (int k = 0; k < N / 4; k++) L[k] = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        float rb;
        round_clamped(isc * (x[i] - mn), 0.f, fmax_l, rb);
        L[i >> 2] = put_byte(L[i >> 2], __float_as_uint(rb), i & 3);
    }
}

// Blackwell packed FP32 (FADD2 / FMUL2, `add/mul.rn.f32x2`): one issue slot performs the IEEE operation on
// two independent floats — measured 73.6 T FP32 op/s against 36.1 T for scalar FADD / FMUL on B200
// (tools/f32x2_probe.cu).  The search is issue-bound, so every per-element operation that is independent
// between neighbouring elements is done on the pairs (x[2k], x[2k+1]); the running sums stay scalar and in
// upstream's order (.x then .y), so every float is rounded exactly as before.
__device__ __forceinline__ float2 bcast2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 clamp2(float2 v, float lo, float hi) { return make_float2(fminf(fmaxf(v.x, lo), hi), fminf(fmaxf(v.y, lo), hi)); }
__device__ __forceinline__ float2 abs2(float2 v) { return make_float2(fabsf(v.x), fabsf(v.y)); }
// clamp(v, 0, hi) in ONE instruction (VIMNMX.RELU, `min.relu.s32` on the float's bits) instead of two FMNMX: non-negative
// floats order like their bit patterns, so the signed minimum with hi's bits is min(v, hi) for v >= 0, and every negative
// float (sign bit set: a negative integer, -0 included) falls to the relu's 0 = +0.0f.  Same value as
// fminf(fmaxf(v, 0), hi) for every v that is not a NaN (+inf -> hi, -inf -> 0); a NaN, which only non-finite input could
// produce here (outside upstream's domain: its nearest_int asserts), gives hi or 0 by its sign instead of 0.
__device__ __forceinline__ float clamp0_relu(float v, float hi) {
    int r;
    asm("min.relu.s32 %0, %1, %2;" : "=r"(r) : "r"(__float_as_int(v)), "r"(__float_as_int(hi)));
    return __int_as_float(r);
}
// (ns*l + nm) per lane with its two roundings.  Scalar on purpose: ptxas 12.9 contracts a mul.rn.f32x2 feeding an
// add.rn.f32x2 into one single-rounding FFMA2 even with --fmad=false (the explicit .rn protects only scalar
// code), which changes results.  No other packed multiply in this file feeds a packed add; the Makefile
// checks that the object contains no FFMA2.
__device__ __forceinline__ float2 affine2_two_roundings(float2 l, float ns, float nm) {
    return make_float2(__fadd_rn(__fmul_rn(l.x, ns), nm), __fadd_rn(__fmul_rn(l.y, ns), nm));
}
// The same two roundings in two packed instructions instead of four scalar ones: p = RN(l * ns) (FMUL2), then
// RN(p * one + nm) (FFMA2) where `one` is 1.0f handed in as a kernel argument — p * 1 is exact, so the FMA's single
// rounding is the rounding of p + nm, and because ptxas cannot know the value of `one` it can neither turn the FMA
// back into an add (and contract it with the multiply) nor fold the multiply into it.  AF 1 selects this form.
template <int AF> __device__ __forceinline__ float2 affine2(float2 l, float ns, float nm, float one) {
    if constexpr (AF) return __ffma2_rn(__fmul2_rn(l, bcast2(ns)), bcast2(one), bcast2(nm));
    else return affine2_two_roundings(l, ns, nm);
}
// rint(clamp(v)) for both lanes.  cvt.rni.f32.f32 (FRND, XU pipe) gives the same round-to-nearest-even integer
// as round_clamped()'s magic-number add but leaves the FP32 pipe, which bounds the search, two operations per
// element lighter (-3..-4.5 % kernel time).  rint(-0.x) is -0 where the add gives +0: a zero code contributes
// +-0 to every sum and product either way, so no compared value changes.
// RM selects where the rounding runs: 0 = FRND for every element (XU pipe, 8 cycles per warp instruction), 1 = the
// magic-number adds for every pair (two packed FP32-pipe instructions per pair), 2 = alternating pairs, which splits the
// rounding work between the two pipes.
// CL 1 (only where lo == 0): the clamp is one VIMNMX.RELU per element (clamp0_relu) instead of two FMNMX.
// CL 2: only the upper bound is applied, CL 3: none (make_qx_quants16 proves which candidates can reach which bound).
template <int RM, int CL = 0> __device__ __forceinline__ float2 round_clamped2(float2 v, float lo, float hi, int k) {
    const float2 c = CL == 1   ? make_float2(clamp0_relu(v.x, hi), clamp0_relu(v.y, hi))
                     : CL == 2 ? make_float2(fminf(v.x, hi), fminf(v.y, hi))
                     : CL == 3 ? v
                               : clamp2(v, lo, hi);
    // (not with CL 3: there `c` is the packed product itself, and ptxas contracts a packed multiply feeding a packed add
    // into one single-rounding FFMA2 — the unclamped candidates always round with FRND)
    if (CL != 3 && (RM == 1 || (RM == 2 && (k & 1)))) return __fadd2_rn(__fadd2_rn(c, bcast2(RMAGIC)), bcast2(-RMAGIC));
    return make_float2(rintf(c.x), rintf(c.y));
}

// The weights are a function of the element alone — |x| for Q2K, av_x + |x| for Q4K / Q5K — so they need not be kept
// in registers across the search: WMODE 0 takes |x| as an operand modifier, WMODE 1 recomputes av + |x| with one packed
// add wherever a pair is used (same operation, same bits), WMODE 2 reads the array `w` as upstream's signature does.
// Q2K ships WMODE 0 (its 16 freed registers buy a seventh resident CTA per SM), Q4K / Q5K WMODE 2 (see k45_lane).
// LF 1: the candidate's codes (floats) wait for the error pass in the lane's own shared-memory row `lrow` (N floats,
// 16-byte aligned, rows 144 bytes apart so the lanes' 128-bit accesses never share a bank) instead of N registers: two
// 128-bit shared accesses per 4 elements and candidate, none of them on the FP32 pipe that bounds the search — and 32
// registers fewer for the 32-element sub-blocks, which is one more resident CTA per SM.
template <int N, bool USE_MAD, int WMODE, int LF, int AF, int RM, int SD, int CL, int SP>
__device__ __forceinline__ float make_qkx2_quants(const float (&x)[N], const float (&w)[N], const float av, const int nmax, float &the_min,
                                                  const float rmin, const float rdelta, const int nstep, float &isc_best, float &mn_best,
                                                  float *__restrict__ lrow, const float one) {
    auto w1 = [&](int i) { return WMODE == 0 ? fabsf(x[i]) : WMODE == 1 ? av + fabsf(x[i]) : w[i]; };
    float mn = x[0], mx = x[0];
    float sum_w = w1(0);
    float sum_x = sum_w * x[0];
#pragma unroll
    for (int i = 1; i < N; ++i) {
        if (x[i] < mn) mn = x[i];
        if (x[i] > mx) mx = x[i];
        const float wi = w1(i);
        sum_w += wi;
        sum_x += wi * x[i];
    }
    if (mn > 0) mn = 0;
    if (mx == mn) {  // L[i] = 0: iscale 0 reproduces that
        the_min = -mn;
        isc_best = 0.f;
        mn_best = mn;
        return 0.f;
    }
    float2 x2[N / 2];
#pragma unroll
    for (int k = 0; k < N / 2; ++k) x2[k] = make_float2(x[2 * k], x[2 * k + 1]);
    auto w2 = [&](int k) {
        if constexpr (WMODE == 0) return abs2(x2[k]);
        else if constexpr (WMODE == 1) return __fadd2_rn(bcast2(av), abs2(x2[k]));
        else return make_float2(w[2 * k], w[2 * k + 1]);
    };
    const float fmax_l = (float)nmax;
    // every candidate's iscale divides by (max - min), which only changes when a candidate is accepted: one correctly
    // rounded reciprocal and three operations per quotient instead of a division each (div_shared_strict)
    // A sub-block whose range is below nmax / FLT_MAX (subnormal weights) overflows iscale to +inf.  Upstream's
    // nearest_int() then sees +-inf or NaN (inf * 0) for every element, and its magic-number add turns all three into a
    // value that the following clamp maps to 0 (release builds; the debug assert is the only other behaviour): the
    // candidate's codes are all zero.  Here the clamp comes first and would map +inf to nmax, so an infinite iscale
    // is replaced by 0 where it multiplies (0 * finite = 0 -> code 0); `scale = 1 / iscale` keeps the true value.
    auto usable = [](float isc) { return isc == __int_as_float(0x7f800000) ? 0.f : isc; };
    SharedDivisor range = shared_divisor(mx - mn);
    float iscale = div_shared_strict<SD>(fmax_l, range);
    float scale = 1 / iscale;
    iscale = usable(iscale);
    float best_mad = 0;
    isc_best = iscale;
    mn_best = mn;
    // error of (scale, min) with codes l:  diff = scale*l + min - x  is evaluated as  x + ((-scale)*l + (-min)) = -diff
    // (round-to-nearest is sign-symmetric, so this is the exact negation; only diff*diff or |diff| is used)
    {
        const float2 nmn = bcast2(-mn), isc2 = bcast2(iscale);
#pragma unroll
        for (int k = 0; k < N / 2; ++k) {
            const float2 l = round_clamped2<RM, CL>(__fmul2_rn(__fadd2_rn(x2[k], nmn), isc2), 0.f, fmax_l, k);
            float2 d = __fadd2_rn(x2[k], affine2<AF>(l, -scale, -mn, one));
            d = USE_MAD ? abs2(d) : __fmul2_rn(d, d);
            const float2 e = __fmul2_rn(w2(k), d);
            best_mad += e.x;
            best_mad += e.y;
        }
    }
    // SP 1: the NEXT candidate's iscale is divided out while this candidate is still being evaluated, from the range as
    // it stands (a candidate is rarely accepted); an accepted candidate changes the range and the quotient is taken
    // again — the same operands either way, so the same bits, but the division's latency (reciprocal, four dependent
    // FMAs, range check: ~50 cycles in which the warp can issue nothing else) leaves the path between two candidates.
    // Measured and left off: bit-identical, but the quotient carried across the candidate costs registers the 128-register
    // kernels do not have (spill 16 -> 24 bytes): Q4K 839 -> 858 us, Q5K 682 -> 700, Q2K 670 -> 690.
    float iscale_next = SP ? div_shared_strict<SD>(rmin + rdelta * 0.f + fmax_l, range) : 0.f;
    for (int is = 0; is <= nstep; ++is) {
        if constexpr (SP) {
            iscale = usable(iscale_next);
            iscale_next = div_shared_strict<SD>(rmin + rdelta * (float)(is + 1) + fmax_l, range);
        } else {
            iscale = usable(div_shared_strict<SD>(rmin + rdelta * (float)is + fmax_l, range));
        }
        float sum_l = 0, sum_l2 = 0, sum_xl = 0;
        float2 lf[LF ? 2 : N / 2];  // LF 1: only the pair being assembled into a 128-bit store
        {
            const float2 nmn = bcast2(-mn), isc2 = bcast2(iscale);
#pragma unroll
            for (int k = 0; k < N / 2; ++k) {
                const float2 l = round_clamped2<RM, CL>(__fmul2_rn(__fadd2_rn(x2[k], nmn), isc2), 0.f, fmax_l, k);
                if constexpr (LF) {
                    lf[k & 1] = l;
                    if (k & 1) *reinterpret_cast<float4 *>(lrow + 2 * (k - 1)) = make_float4(lf[0].x, lf[0].y, lf[1].x, lf[1].y);
                } else {
                    lf[k] = l;
                }
                const float2 wl = __fmul2_rn(w2(k), l);
                const float2 wl2 = __fmul2_rn(wl, l), wlx = __fmul2_rn(wl, x2[k]);
                sum_l += wl.x;
                sum_l2 += wl2.x;
                sum_xl += wlx.x;
                sum_l += wl.y;
                sum_l2 += wl2.y;
                sum_xl += wlx.y;
            }
        }
        const float D = sum_w * sum_l2 - sum_l * sum_l;
        if (D > 0) {
            const SharedDivisor sdD = shared_divisor(D);  // two quotients by D
            float this_scale = div_shared_strict<SD>(sum_w * sum_xl - sum_x * sum_l, sdD);
            float this_min = div_shared_strict<SD>(sum_l2 * sum_x - sum_l * sum_xl, sdD);
            if (this_min > 0) {
                this_min = 0;
                this_scale = sum_xl / sum_l2;
            }
            float mad = 0;
#pragma unroll
            for (int k = 0; k < N / 2; ++k) {
                float2 lk;
                if constexpr (LF) {
                    if ((k & 1) == 0) {
                        const float4 v = *reinterpret_cast<const float4 *>(lrow + 2 * k);
                        lf[0] = make_float2(v.x, v.y);
                        lf[1] = make_float2(v.z, v.w);
                    }
                    lk = lf[k & 1];
                } else {
                    lk = lf[k];
                }
                float2 d = __fadd2_rn(x2[k], affine2<AF>(lk, -this_scale, -this_min, one));
                d = USE_MAD ? abs2(d) : __fmul2_rn(d, d);
                const float2 e = __fmul2_rn(w2(k), d);
                mad += e.x;
                mad += e.y;
            }
            if (mad < best_mad) {
                isc_best = iscale;  // the codes just evaluated came from (iscale, mn before the update)
                mn_best = mn;
                best_mad = mad;
                scale = this_scale;
                mn = this_min;
                range = shared_divisor(mx - mn);
                if constexpr (SP) iscale_next = div_shared_strict<SD>(rmin + rdelta * (float)(is + 1) + fmax_l, range);
            }
        }
    }
    the_min = -mn;
    return scale;
}

// ---- upstream make_qx_quants(n=16, nmax, rmse_type=1, qw=NULL); codes stored as l + nmax -----------
// Same deferral: the search returns the iscale of its best candidate; qx_codes16() is upstream's
// `L[i] = nmax + clamp(nearest_int(iscale * x[i]))`, evaluated only when the caller needs the codes.
__device__ __forceinline__ void qx_codes16(const float (&x)[16], const float isc, const int nmax, uint32_t (&L)[4]) {
    const float lo = (float)(-nmax), hi = (float)(nmax - 1), fn = (float)nmax;
#pragma unroll
    for (int k = 0; k < 4; k++) L[k] = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float rb;
        round_clamped(isc * x[i], lo, hi, rb);
        L[i >> 2] = put_byte(L[i >> 2], __float_as_uint(rb + fn), i & 3);  // low byte of (l + nmax)
    }
}
// returns the scale; `isc_best` = iscale of the kept codes; `all_zero`: upstream's early exit (L[i] = 0, raw)
// WS 1: the per-element constants w = x*x and w*x live in the lane's shared-memory row (32 floats: w at [0, 16), w*x at
// [16, 32)) instead of 32 registers.
// CS 1: the clamp to [-nmax, nmax - 1] is applied only where a candidate can reach it.  Every |x| <= |mx|, and candidate
// `is` multiplies by iscale = RN(-(nmax + RN(0.1f is)) / mx), so |iscale x| <= (nmax + 0.1 is) (1 + 2^-23) (1 + 2^-24)^2:
//   is <= -6:  |v| <= 31.40001 < 31.5   -> rint(v) in [-31, 31]: neither bound can be reached (4 candidates)
//   is <=  4:  |v| <= 32.40001 < 32.5   -> rint(v) in [-32, 32]: only the upper bound nmax - 1 = 31 (9 candidates and
//              the first evaluation, whose iscale is -nmax / mx: |v| <= 32.00001)
//   is >=  5:  both (RN(-32.5 / mx) mx may land beyond the tie at -32.5)                       (5 candidates)
// (nmax = 32; the bounds use nmax only through these constants, hence the static_assert on the caller's side.)
// The values are the same floats whether a bound that cannot bind is applied or not: 32 + 16 fewer instructions per 16
// elements in the first two groups, on a kernel that is issue-bound.
template <int WS, int RM, int SD, int CS = 0>
__device__ __forceinline__ float make_qx_quants16(const float (&x)[16], const int nmax, float &isc_best, bool &all_zero, float *__restrict__ lrow) {
    float mx = 0, amax = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float ax = fabsf(x[i]);
        if (ax > amax) { amax = ax; mx = x[i]; }
    }
    isc_best = 0.f;
    all_zero = amax < GROUP_MAX_EPS;
    if (all_zero) return 0.f;
    const float lo = (float)(-nmax), hi = (float)(nmax - 1), fn = (float)nmax;
    float2 x2[8], w2[WS ? 2 : 8], wx2[WS ? 2 : 8];  // packed pairs (see make_qkx2_quants); w*x*l is evaluated (w*x)*l, w*l*l as (w*l)*l
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        x2[k] = make_float2(x[2 * k], x[2 * k + 1]);
        const float2 w = __fmul2_rn(x2[k], x2[k]), wx = __fmul2_rn(w, x2[k]);
        if constexpr (WS) {
            w2[k & 1] = w;
            wx2[k & 1] = wx;
            if (k & 1) {
                *reinterpret_cast<float4 *>(lrow + 2 * (k - 1)) = make_float4(w2[0].x, w2[0].y, w2[1].x, w2[1].y);
                *reinterpret_cast<float4 *>(lrow + 16 + 2 * (k - 1)) = make_float4(wx2[0].x, wx2[0].y, wx2[1].x, wx2[1].y);
            }
        } else {
            w2[k] = w;
            wx2[k] = wx;
        }
    }
    // sumlx, suml2 for `isc`, accumulated in element order
    auto sums = [&](float isc, float &sumlx, float &suml2, auto clamp_mode) {
        constexpr int CLM = decltype(clamp_mode)::value;   // 0 both bounds, 2 upper only, 3 none (round_clamped2)
        const float2 isc2 = bcast2(isc);
        sumlx = suml2 = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float2 l = round_clamped2<RM, CLM>(__fmul2_rn(isc2, x2[k]), lo, hi, k);
            float2 wk, wxk;
            if constexpr (WS) {
                if ((k & 1) == 0) {
                    const float4 a4 = *reinterpret_cast<const float4 *>(lrow + 2 * k), b4 = *reinterpret_cast<const float4 *>(lrow + 16 + 2 * k);
                    w2[0] = make_float2(a4.x, a4.y); w2[1] = make_float2(a4.z, a4.w);
                    wx2[0] = make_float2(b4.x, b4.y); wx2[1] = make_float2(b4.z, b4.w);
                }
                wk = w2[k & 1];
                wxk = wx2[k & 1];
            } else {
                wk = w2[k];
                wxk = wx2[k];
            }
            const float2 a = __fmul2_rn(wxk, l), b = __fmul2_rn(__fmul2_rn(wk, l), l);
            sumlx += a.x;
            suml2 += b.x;
            sumlx += a.y;
            suml2 += b.y;
        }
    };
    const SharedDivisor smx = shared_divisor(mx);  // all 19 candidates divide by mx
    float iscale = div_shared_strict<SD>(lo, smx);
    using Both = std::integral_constant<int, 0>;
    using Upper = std::integral_constant<int, CS ? 2 : 0>;
    using None = std::integral_constant<int, CS ? 3 : 0>;
    float sumlx, suml2;
    sums(iscale, sumlx, suml2, Upper{});
    isc_best = iscale;
    float scale = suml2 ? sumlx / suml2 : 0.0f;
    float best = scale * sumlx;
    auto candidate = [&](int is, auto clamp_mode) {
        iscale = div_shared_strict<SD>(-(fn + 0.1f * (float)is), smx);
        sums(iscale, sumlx, suml2, clamp_mode);
        if (suml2 > 0 && sumlx * sumlx > best * suml2) {
            isc_best = iscale;
            scale = sumlx / suml2;
            best = scale * sumlx;
        }
    };
    if constexpr (CS) {
        for (int is = -9; is <= -6; ++is) candidate(is, None{});
        for (int is = -5; is <= 4; ++is)
            if (is != 0) candidate(is, Upper{});
        for (int is = 5; is <= 9; ++is) candidate(is, Both{});
    } else {
        for (int is = -9; is <= 9; ++is)
            if (is != 0) candidate(is, Both{});
    }
    return scale;
}

// ---- upstream make_q3_quants(n=16, nmax=4, do_rmse=true); codes stored as l + nmax ----------------
__device__ __forceinline__ float make_q3_quants16(const float (&x)[16], const int nmax, uint32_t (&Lout)[4]) {
    float mx = 0, amax = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float ax = fabsf(x[i]);
        if (ax > amax) { amax = ax; mx = x[i]; }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) Lout[k] = 0;
    if (amax < GROUP_MAX_EPS) return 0.f;
    const float iscale = (float)(-nmax) / mx;
    float sumlx = 0, suml2 = 0;
    int L[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        int l = nearest_int(iscale * x[i]);
        l = max(-nmax, min(nmax - 1, l));
        L[i] = l;
        const float w = x[i] * x[i];
        sumlx += w * x[i] * i2f_small(l);
        suml2 += w * i2f_small(l) * i2f_small(l);
    }
    for (int itry = 0; itry < 5; ++itry) {
        int n_changed = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float w = x[i] * x[i];
            float slx = sumlx - w * x[i] * i2f_small(L[i]);
            if (slx > 0) {
                float sl2 = suml2 - w * i2f_small(L[i]) * i2f_small(L[i]);
                int new_l = nearest_int(x[i] * sl2 / slx);
                new_l = max(-nmax, min(nmax - 1, new_l));
                if (new_l != L[i]) {
                    slx += w * x[i] * i2f_small(new_l);
                    sl2 += w * i2f_small(new_l) * i2f_small(new_l);
                    if (sl2 > 0 && slx * slx * suml2 > sumlx * sumlx * sl2) {
                        L[i] = new_l;
                        sumlx = slx;
                        suml2 = sl2;
                        ++n_changed;
                    }
                }
            }
        }
        if (!n_changed) break;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) set_code(Lout, i, L[i] + nmax);
    return sumlx / suml2;
}

// ---- group reductions over the G lanes of one super-block ------------------------------------------
template <int G> __device__ __forceinline__ float group_max_from_zero(float v) {  // max_x = 0; if (v > max_x) max_x = v
    float m = v > 0.f ? v : 0.f;
#pragma unroll
    for (int s = 1; s < G; s <<= 1) { const float o = __shfl_xor_sync(KFULL, m, s); m = o > m ? o : m; }
    return m;
}
// first value (lowest lane) with strictly greatest |v|, starting from 0
template <int G> __device__ __forceinline__ float group_max_by_abs(float v, int lane) {
    float acc = fabsf(v) > 0.f ? v : 0.f;
#pragma unroll
    for (int s = 1; s < G; s <<= 1) {
        const float o = __shfl_xor_sync(KFULL, acc, s);
        const bool lower = (lane & s) == 0;
        const float first = lower ? acc : o, second = lower ? o : acc;
        acc = fabsf(second) > fabsf(first) ? second : first;
    }
    return acc;
}

// ---- per-type quantizers ---------------------------------------------------------------------------
// Scratch written by the lanes of one super-block, read by the bytewise assembler.
struct KScratch {
    uint8_t L[256];   // final codes, one byte per element
    uint8_t a[16];    // per-sub-block code A (scale)
    uint8_t b[16];    // per-sub-block code B (min)
    uint16_t d16, dmin16;
    uint32_t zero;    // whole block is zero bytes
};

template <uint32_t T> struct KQuant;

template <int SUB> __device__ __forceinline__ void put_codes(KScratch &s, int j, const uint32_t (&L)[SUB / 4]) {
#pragma unroll
    for (int k = 0; k < SUB / 4; k++) *reinterpret_cast<uint32_t *>(&s.L[j * SUB + 4 * k]) = L[k];
}

// Q4K and Q5K share everything but nmax / search range / final layout.
template <int NMAX, int LF, int WM, int AF, int RM, int SD, int CL, int SP> __device__ __forceinline__ void k45_lane(const float (&x)[32], int j, KScratch &s, float rmin, int nstep, float *lrow, float one) {
    float sum_x2 = 0;
#pragma unroll
    for (int l = 0; l < 32; ++l) sum_x2 += x[l] * x[l];
    const float av_x = sqrtf(sum_x2 / 32);
    // weights kept in registers (WMODE 2): with 32-element sub-blocks a fifth CTA per SM would need <= 96 registers,
    // which spills (measured: 940 vs 841 us for Q4K), and without it recomputing the weights is only extra work
    float w[WM == 2 ? 32 : 1];
    if constexpr (WM == 2) {
#pragma unroll
        for (int l = 0; l < 32; ++l) w[l] = av_x + fabsf(x[l]);
    }
    uint32_t L[8];
    float the_min, isc_best, mn_best;
    float scale;
    if constexpr (WM == 2) scale = make_qkx2_quants<32, false, 2, LF, AF, RM, SD, CL, SP>(x, w, av_x, NMAX, the_min, rmin, 0.1f, nstep, isc_best, mn_best, lrow, one);
    else scale = make_qkx2_quants<32, false, 1, LF, AF, RM, SD, CL, SP>(x, x, av_x, NMAX, the_min, rmin, 0.1f, nstep, isc_best, mn_best, lrow, one);
    const float max_scale = group_max_from_zero<8>(scale), max_min = group_max_from_zero<8>(the_min);
    const float inv_scale = max_scale > 0 ? 63.f / max_scale : 0.f;
    const float inv_min = max_min > 0 ? 63.f / max_min : 0.f;
    uint32_t ls = (uint32_t)nearest_int_of_product(inv_scale * scale) & 0xFFu;
    uint32_t lm = (uint32_t)nearest_int_of_product(inv_min * the_min) & 0xFFu;
    ls = min(ls, 63u);
    lm = min(lm, 63u);
    const uint16_t d16 = f2h(max_scale / 63.f), dmin16 = f2h(max_min / 63.f);
    const float d = h2f(d16) * (float)ls;
    if (d != 0.f) {
        const float dm = h2f(dmin16) * (float)lm;
#pragma unroll
        for (int k = 0; k < 8; k++) L[k] = 0;
        const SharedDivisor sd = shared_divisor(d);
#pragma unroll
        for (int ii = 0; ii < 32; ++ii) {
            int l = nearest_int(div_shared(x[ii] + dm, sd));
            l = max(0, min(NMAX, l));
            set_code(L, ii, l);
        }
    } else {
        qkx2_codes<32>(x, isc_best, mn_best, NMAX, L);  // upstream `if (!d) continue;`: the search's codes stay
    }
    put_codes<32>(s, j, L);
    s.a[j] = (uint8_t)ls;
    s.b[j] = (uint8_t)lm;
    if (j == 0) { s.d16 = d16; s.dmin16 = dmin16; s.zero = 0; }
}
// header bytes 0..15 of Q4K / Q5K: delta, min, 12 bytes of 6-bit scales/mins — byte k of the 12: k < 4: a[k] | (a[k+4] >> 4) << 6;
// k < 8: b[k-4] | (b[k] >> 4) << 6; else (a[k-4] & 15) | (b[k-4] & 15) << 4 —

// the same 16 header bytes as four little-endian words (a[] and b[] hold 6-bit codes)
__device__ __forceinline__ uint32_t k45_header_word(const KScratch &s, int w) {
    if (w == 0) return (uint32_t)s.d16 | ((uint32_t)s.dmin16 << 16);
    const uint32_t a0 = *reinterpret_cast<const uint32_t *>(&s.a[0]), a1 = *reinterpret_cast<const uint32_t *>(&s.a[4]);
    const uint32_t b0 = *reinterpret_cast<const uint32_t *>(&s.b[0]), b1 = *reinterpret_cast<const uint32_t *>(&s.b[4]);
    if (w == 1) return a0 | ((a1 & 0x30303030u) << 2);
    if (w == 2) return b0 | ((b1 & 0x30303030u) << 2);
    return (a1 & 0x0F0F0F0Fu) | ((b1 & 0x0F0F0F0Fu) << 4);
}
__device__ __forceinline__ uint32_t code_word(const KScratch &s, int w) { return reinterpret_cast<const uint32_t *>(s.L)[w]; }

// Types whose block is a whole number of 32-bit words also provide word(s, w), the w-th little-endian word of the
// packed block: the assembler then runs one iteration per four output bytes (the bytewise assembler this replaced
// cost ~25 instructions and 2-4 dependent one-byte shared loads per output byte; that phase was 5 % of the executed
// instructions but 10 % of the stall samples of the Q4K kernel).  Q3K / Q6K blocks (110 / 210 bytes) end in a 16-bit delta,
// returned by tail(s).
template <> struct KQuant<T_Q4K> {
    static constexpr int SUB = 32;
    static constexpr int WORDS = 36, TAIL = 0, REGS = 128;
    static __device__ __forceinline__ uint32_t word(const KScratch &s, int w) {
        if (w < 4) return k45_header_word(s, w);
        const int t = w - 4, p = t >> 3, jj = t & 7;
        return code_word(s, 16 * p + jj) | (code_word(s, 16 * p + 8 + jj) << 4);
    }
    template <class CFG> static __device__ __forceinline__ void lane(const float (&x)[32], int j, int, KScratch &s, float *lrow, float one) { k45_lane<15, CFG::LF, CFG::WM, CFG::AF, CFG::RM, CFG::SD, CFG::CL, CFG::SP>(x, j, s, -1.f, 20, lrow, one); }
};
template <> struct KQuant<T_Q5K> {
    static constexpr int SUB = 32;
    static constexpr int WORDS = 44, TAIL = 0, REGS = 128;
    static __device__ __forceinline__ uint32_t word(const KScratch &s, int w) {
        if (w < 4) return k45_header_word(s, w);
        if (w < 12) {  // qh: bit 2p / 2p+1 of byte l = fifth bit of elements 64p+l / 64p+32+l
            const int jj = w - 4;
            uint32_t h = 0;
#pragma unroll
            for (int p = 0; p < 4; p++)
                h |= (((code_word(s, 16 * p + jj) >> 4) & 0x01010101u) << (2 * p)) | (((code_word(s, 16 * p + 8 + jj) >> 4) & 0x01010101u) << (2 * p + 1));
            return h;
        }
        const int t = w - 12, p = t >> 3, jj = t & 7;
        return (code_word(s, 16 * p + jj) & 0x0F0F0F0Fu) | ((code_word(s, 16 * p + 8 + jj) & 0x0F0F0F0Fu) << 4);
    }
    template <class CFG> static __device__ __forceinline__ void lane(const float (&x)[32], int j, int, KScratch &s, float *lrow, float one) { k45_lane<31, CFG::LF, CFG::WM, CFG::AF, CFG::RM, CFG::SD, CFG::CL, CFG::SP>(x, j, s, -0.5f, 15, lrow, one); }
};

template <> struct KQuant<T_Q6K> {
    static constexpr int SUB = 16;
    // 210 bytes = 52 words (ql 32, qh 16, scales 4) + the 16-bit delta.  93 registers = 5 four-warp CTAs per SM;
    // capping the kernel at 6 CTAs per SM is slower
    static constexpr int WORDS = 52, TAIL = 1, REGS = 96;
    static __device__ __forceinline__ uint32_t word(const KScratch &s, int w) {
        if (s.zero) return 0;
        if (w < 32) {  // ql
            const int i = 32 * (w >> 4) + (w & 15);
            return (code_word(s, i) & 0x0F0F0F0Fu) | ((code_word(s, i + 16) & 0x0F0F0F0Fu) << 4);
        }
        if (w < 48) {  // qh
            const int jj = w - 32, i = 32 * (jj >> 3) + (jj & 7);
            return ((code_word(s, i) >> 4) & 0x03030303u) | (((code_word(s, i + 8) >> 4) & 0x03030303u) << 2) |
                   (((code_word(s, i + 16) >> 4) & 0x03030303u) << 4) | (((code_word(s, i + 24) >> 4) & 0x03030303u) << 6);
        }
        return *reinterpret_cast<const uint32_t *>(&s.a[4 * (w - 48)]);
    }
    static __device__ __forceinline__ uint32_t tail(const KScratch &s) { return s.zero ? 0u : (uint32_t)s.d16; }
    template <class CFG> static __device__ __forceinline__ void lane(const float (&x)[16], int j, int lane_id, KScratch &s, float *lrow, float) {
        uint32_t L[4] = {0, 0, 0, 0};
        float isc_best;
        bool sub_zero;
        const float scale = make_qx_quants16<CFG::LF, CFG::RM, CFG::SD, CFG::CS>(x, 32, isc_best, sub_zero, lrow);   // nmax = 32: see CS
        const float max_scale = group_max_by_abs<16>(scale, lane_id);
        const bool zero = fabsf(max_scale) < GROUP_MAX_EPS;
        int sc = 0;
        uint16_t d16 = 0;
        if (!zero) {  // a zero super-block is written as zero bytes whatever L holds
            const float iscale = -128.f / max_scale;
            d16 = f2h(1 / iscale);
            sc = min(127, nearest_int(iscale * scale));
            const float d = h2f(d16) * (float)(int)(int8_t)sc;
            if (d != 0.f) {
                const SharedDivisor sd = shared_divisor(d);
#pragma unroll
                for (int ii = 0; ii < 16; ++ii) {
                    int l = nearest_int(div_shared(x[ii], sd));
                    l = max(-32, min(31, l));
                    set_code(L, ii, l + 32);
                }
            } else if (!sub_zero) {
                qx_codes16(x, isc_best, 32, L);  // upstream `if (!d) continue;`: the search's codes stay
            }
        }
        put_codes<16>(s, j, L);
        s.a[j] = (uint8_t)(int8_t)sc;
        if (j == 0) { s.d16 = d16; s.dmin16 = 0; s.zero = zero ? 1u : 0u; }
    }
};

template <> struct KQuant<T_Q2K> {
    static constexpr int SUB = 16;
    static constexpr int WORDS = 21, TAIL = 0, REGS = 72;  // 72 registers now that the weights are an operand modifier
    static __device__ __forceinline__ uint32_t word(const KScratch &s, int w) {
        if (w < 4) return *reinterpret_cast<const uint32_t *>(&s.a[4 * w]);
        if (w < 20) {
            const int t = w - 4, n = t >> 3, jj = t & 7, base = 32 * n + jj;
            return code_word(s, base) | (code_word(s, base + 8) << 2) | (code_word(s, base + 16) << 4) | (code_word(s, base + 24) << 6);
        }
        return (uint32_t)s.d16 | ((uint32_t)s.dmin16 << 16);
    }
    template <class CFG> static __device__ __forceinline__ void lane(const float (&x)[16], int j, int, KScratch &s, float *lrow, float one) {
        // weights = |x|: an operand modifier, never materialised (WMODE 0) — 7 CTAs per SM instead of 6, -4 % time
        uint32_t L[4];
        float the_min, isc_best, mn_best;
        const float scale = make_qkx2_quants<16, true, 0, CFG::LF, CFG::AF, CFG::RM, CFG::SD, CFG::CL, CFG::SP>(x, x, 0.f, 3, the_min, -0.5f, 0.1f, 15, isc_best, mn_best, lrow, one);
        const float max_scale = group_max_from_zero<16>(scale), max_min = group_max_from_zero<16>(the_min);
        uint32_t b = 0;
        uint16_t d16 = 0, dmin16 = 0;
        if (max_scale > 0) {
            const float iscale = 15.f / max_scale;
            b = (uint32_t)nearest_int_of_product(iscale * scale) & 0xFFu;
            d16 = f2h(max_scale / 15.f);
        }
        if (max_min > 0) {
            const float iscale = 15.f / max_min;
            b |= ((uint32_t)nearest_int_of_product(iscale * the_min) << 4) & 0xFFu;
            dmin16 = f2h(max_min / 15.f);
        }
        const float d = h2f(d16) * (float)(b & 0xFu);
        if (d != 0.f) {
            const float dm = h2f(dmin16) * (float)(b >> 4);
#pragma unroll
            for (int k = 0; k < 4; k++) L[k] = 0;
            const SharedDivisor sd = shared_divisor(d);
#pragma unroll
            for (int ii = 0; ii < 16; ++ii) {
                int l = nearest_int(div_shared(x[ii] + dm, sd));
                l = max(0, min(3, l));
                set_code(L, ii, l);
            }
        } else {
            qkx2_codes<16>(x, isc_best, mn_best, 3, L);  // upstream `if (!d) continue;`: the search's codes stay
        }
        put_codes<16>(s, j, L);
        s.a[j] = (uint8_t)b;
        if (j == 0) { s.d16 = d16; s.dmin16 = dmin16; s.zero = 0; }
    }
};

template <> struct KQuant<T_Q3K> {
    static constexpr int SUB = 16;
    // 110 bytes = 27 words (hmask 8, qs 16, scales 3) + the 16-bit delta.  96 registers, no spills: a fifth
    // resident CTA per SM, -8 % time
    static constexpr int WORDS = 27, TAIL = 1, REGS = 96;
    static __device__ __forceinline__ uint32_t word(const KScratch &s, int w) {
        if (w < 8) {  // hmask: bit bq of byte o = (code[32 bq + o] > 3) = bit 2 of a 3-bit code
            uint32_t m = 0;
#pragma unroll
            for (int bq = 0; bq < 8; bq++) m |= ((code_word(s, 8 * bq + w) >> 2) & 0x01010101u) << bq;
            return m;
        }
        if (w < 24) {
            const int jj = w - 8, i = 32 * (jj >> 3) + (jj & 7);
            return (code_word(s, i) & 0x03030303u) | ((code_word(s, i + 8) & 0x03030303u) << 2) | ((code_word(s, i + 16) & 0x03030303u) << 4) |
                   ((code_word(s, i + 24) & 0x03030303u) << 6);
        }
        const uint32_t *A = reinterpret_cast<const uint32_t *>(s.a);  // 16 six-bit scale codes
        if (w < 26) return (A[w - 24] & 0x0F0F0F0Fu) | ((A[w - 22] & 0x0F0F0F0Fu) << 4);
        return ((A[0] >> 4) & 0x03030303u) | (((A[1] >> 4) & 0x03030303u) << 2) | (((A[2] >> 4) & 0x03030303u) << 4) | (((A[3] >> 4) & 0x03030303u) << 6);
    }
    static __device__ __forceinline__ uint32_t tail(const KScratch &s) { return s.d16; }
    template <class CFG> static __device__ __forceinline__ void lane(const float (&x)[16], int j, int lane_id, KScratch &s, float *, float) {
        uint32_t L[4];
        const float scale = make_q3_quants16(x, 4, L);
        const float max_scale = group_max_by_abs<16>(scale, lane_id);
        int c = 0;
        uint16_t d16 = 0;
        if (max_scale != 0.f) {
            const float iscale = -32.f / max_scale;
            int l = (int)(int8_t)nearest_int(iscale * scale);
            c = max(-32, min(31, l)) + 32;
            d16 = f2h(1 / iscale);
        }
        const float d = h2f(d16) * (float)(c - 32);
        if (d != 0.f) {
#pragma unroll
            for (int k = 0; k < 4; k++) L[k] = 0;
            const SharedDivisor sd = shared_divisor(d);
#pragma unroll
            for (int ii = 0; ii < 16; ++ii) {
                int l = nearest_int(div_shared(x[ii], sd));
                l = max(-4, min(3, l));
                set_code(L, ii, l + 4);
            }
        }
        put_codes<16>(s, j, L);
        s.a[j] = (uint8_t)c;
        if (j == 0) { s.d16 = d16; s.dmin16 = 0; s.zero = 0; }
    }
};

// ---------------------------------------------------------------------------------------------
// CFG: WARPS per CTA (warps are independent: own staging, own scratch, __syncwarp only — so a CTA may as well be one warp,
// which lets occupancy follow the register cap in steps of one warp rather than four), REGS (register cap per thread,
// __maxnreg__), LF (per-lane shared-memory row: candidate codes for the qkx2 searches, w / w*x for Q6K),
// WM (Q4K / Q5K weights: 2 registers, 1 recomputed), STAGES (input rows in flight).
template <int WARPS_, int REGS_, int LF_, int WM_, int STAGES_, int AF_ = 0, int RM_ = 0, int SD_ = 0, int CL_ = 0, int SP_ = 0, int CS_ = 0> struct KqCfg {
    static constexpr int WARPS = WARPS_, THREADS = WARPS_ * 32, REGS = REGS_, LF = LF_, WM = WM_, STAGES = STAGES_, AF = AF_, RM = RM_, SD = SD_, CL = CL_, SP = SP_, CS = CS_;
};
constexpr int KQ_LROW = 36;  // floats between the lanes' rows: 144 bytes, so eight lanes' 128-bit accesses cover all 32 banks

template <uint32_t T, class FT, class CFG>
__global__ void __launch_bounds__(CFG::THREADS) __maxnreg__(CFG::REGS)
quant_k_kernel(const typename FT::raw *__restrict__ src, uint8_t *__restrict__ dst, size_t nblocks, const float one, unsigned long long *__restrict__ work) {
    using KQ = KQuant<T>;
    constexpr int NSTG = CFG::STAGES, KQ_WARPS = CFG::WARPS, KQ_THREADS = CFG::THREADS;
    constexpr int SUB = KQ::SUB, NSUB = 256 / SUB, SBW = 32 / NSUB;  // super-blocks per warp pass
    constexpr int BYTES = BlockTraits<T>::BYTES;
    constexpr int OUT_BYTES = (SBW * BYTES + 15) & ~15;

    // Input staging: the SBW super-blocks of a warp pass are SBW*256 consecutive elements = 32 sub-block rows.
    // They are copied raw (16-byte cp.async, coalesced, no registers) into rows padded by 16 bytes, so that every
    // lane's LDS.128 of its own row is bank-conflict free, two stages deep: the rows of the NEXT pass are in flight
    // while this pass runs its search (the synchronous load-convert-scatter this replaces was 1.5 % of the executed
    // instructions but 7 % of the stall samples: every warp sat out a full HBM round trip per pass).
    constexpr int ROW_BYTES = SUB * FT::SIZE, RSTRIDE = ROW_BYTES + 16, CPR = ROW_BYTES / 16;
    __shared__ __align__(16) uint8_t xraw[KQ_WARPS][NSTG][32 * RSTRIDE];
    __shared__ KScratch scratch[KQ_WARPS][SBW];
    __shared__ __align__(16) uint8_t outb[KQ_WARPS][OUT_BYTES];
    __shared__ __align__(16) float lrows[CFG::LF ? KQ_THREADS * KQ_LROW : 4];
    float *lrow = lrows + (CFG::LF ? threadIdx.x * KQ_LROW : 0);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sbi = lane / NSUB, j = lane % NSUB;
    const size_t ngroups = (nblocks + SBW - 1) / SBW;
    const bool vec_in = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
    const size_t gstep = (size_t)gridDim.x * KQ_WARPS;

    auto issue = [&](size_t g, int stage) {  // request the rows of warp pass g (a no-op group past the end)
        if (g < ngroups) {
            const size_t sb0 = g * SBW;
            const int nsb = (int)min((size_t)SBW, nblocks - sb0);
            const uint8_t *in = reinterpret_cast<const uint8_t *>(src + sb0 * 256);
            uint8_t *st = xraw[warp][stage];
            const int live = nsb * (256 / SUB) * CPR;  // chunks that exist; the rest of the pass is zero rows
#pragma unroll
            for (int c = 0; c < CPR; c++) {
                const int ch = lane + 32 * c, row = ch / CPR, cc = ch % CPR;
                uint8_t *sdst = st + row * RSTRIDE + cc * 16;
                if (ch < live) {
                    if (vec_in) {
                        cp_async16(sdst, in + (size_t)ch * 16);
                    } else {  // source only element-aligned: synchronous element copies
                        using RAW = typename FT::raw;
                        const RAW *ge = reinterpret_cast<const RAW *>(in + (size_t)ch * 16);
#pragma unroll
                        for (int e = 0; e < 16 / FT::SIZE; e++) reinterpret_cast<RAW *>(sdst)[e] = ge[e];
                    }
                } else {
                    *reinterpret_cast<uint4 *>(sdst) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
        }
        cp_async_commit();
    };

    // The pass after `cur`.  `work` (a launch-wide counter, zero at launch) hands the passes out as tickets: the first
    // pass of every warp is its own index, tickets start behind those.  A pass takes tens of microseconds and a
    // data-dependent time (Q3K's refinement loops most of all); with a fixed stride the warps of an SM also stay in the
    // same phase of the search and queue for the same pipe.  Measured against the fixed stride on 58.7 M elements:
    // Q4K 836 -> 769 us, Q5K 678 -> 632, Q2K 664 -> 621, Q6K 481 -> 449, Q3K 449 -> 428.  `work == nullptr` (small
    // inputs: at most one pass per warp) keeps the fixed stride.
    auto next_of = [&](size_t cur) -> size_t {
        if (work != nullptr) {
            unsigned long long t = 0;
            if (lane == 0) t = atomicAdd(work, 1ull);
            return gstep + (size_t)__shfl_sync(KFULL, t, 0);
        }
        return cur + gstep;
    };
    size_t g = (size_t)blockIdx.x * KQ_WARPS + warp, gnext = 0;
    int stage = 0;
    if constexpr (NSTG == 2) issue(g, 0);
    for (; g < ngroups; g = gnext, stage ^= (NSTG - 1)) {
        const size_t sb0 = g * SBW;
        const int nsb = (int)min((size_t)SBW, nblocks - sb0);
        gnext = next_of(g);
        if constexpr (NSTG == 2) {
            issue(gnext, stage ^ 1);  // the other stage was consumed before the __syncwarp that ended the previous pass
            cp_async_wait<1>();           // this lane's chunks of pass g have landed ...
        } else {
            issue(g, 0);                  // single stage: the rows were consumed into registers before the previous pass searched
            cp_async_wait<0>();
        }
        __syncwarp();                 // ... and so have everybody else's
        float x[SUB];
        {
            const uint8_t *row = xraw[warp][stage] + lane * RSTRIDE;
            if constexpr (FT::SIZE == 4) {
#pragma unroll
                for (int i = 0; i < SUB / 4; i++) {
                    const float4 v = *reinterpret_cast<const float4 *>(row + 16 * i);
                    x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < SUB / 8; i++) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(row + 16 * i);
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if constexpr (std::is_same<FT, F16>::value) {
                            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[q]));
                            x[8 * i + 2 * q] = f.x; x[8 * i + 2 * q + 1] = f.y;
                        } else {
                            x[8 * i + 2 * q] = __uint_as_float(w[q] << 16);
                            x[8 * i + 2 * q + 1] = __uint_as_float(w[q] & 0xFFFF0000u);
                        }
                    }
                }
            }
        }
        KQ::template lane<CFG>(x, j, lane, scratch[warp][sbi], lrow, one);
        __syncwarp();
        {   // assemble the packed blocks: one 32-bit word (plus, for the 110- / 210-byte blocks, one trailing 16-bit
            // delta) per lane and iteration; odd Q3K / Q6K blocks start 2 bytes off a word boundary
            constexpr int U = KQ::WORDS + KQ::TAIL;
            static_assert(KQ::WORDS * 4 + KQ::TAIL * 2 == BYTES, "word() / tail() cover the whole block");
            for (int u = lane; u < nsb * U; u += 32) {
                const int sb = u / U, k = u % U;
                uint8_t *o = outb[warp] + sb * BYTES + 4 * k;
                if (KQ::TAIL && k == KQ::WORDS) {
                    if constexpr (KQ::TAIL != 0) *reinterpret_cast<uint16_t *>(o) = (uint16_t)KQ::tail(scratch[warp][sb]);
                } else {
                    const uint32_t v = KQ::word(scratch[warp][sb], k);
                    if (KQ::TAIL && (sb & 1)) {
                        *reinterpret_cast<uint16_t *>(o) = (uint16_t)v;
                        *reinterpret_cast<uint16_t *>(o + 2) = (uint16_t)(v >> 16);
                    } else {
                        *reinterpret_cast<uint32_t *>(o) = v;
                    }
                }
            }
        }
        __syncwarp();
        cta_copy_s2g(dst + sb0 * BYTES, outb[warp], (uint32_t)(nsb * BYTES), lane, 32);
        __syncwarp();
    }
    cp_async_wait<0>();
}

}  // namespace ggq
