// quant_legacy.cu — f32 / f16 / bf16 -> Q4_0 Q4_1 Q5_0 Q5_1 Q8_0 Q8_1 (32-element blocks) and Q8K
// (256-element block, reference layout), plus the f32/f16/bf16 element casts; sm_100a.
// Compiled with -fmad=false: the reference's Rust never fuses `x * recip + 8.5`, so neither may this.
//
// Replaces the per-block `Quantize::quantize` bodies behind `QuantExt::quantize_slice`
// (/root/reference/ggml-quants/src/lib.rs:121-133; per-type bodies cited at each encoder).
//
// Shape of the kernels (input-dominated stream; the first version with 4 lanes per block spent ~21
// instructions per element on shuffles / redundant divides and was issue-bound at 38 % of HBM peak):
//   * ONE THREAD PER 32-ELEMENT ROW (a legacy block, or 1/8 of a Q8K super-block): the folds of
//     structs.rs:91-107 are register-local, delta/recip cost two divides per 32 elements;
//   * rows are staged global -> shared with 16-byte cp.async (coalesced, no registers): 128-byte f32 rows padded by
//     16 bytes, 64-byte f16 / bf16 rows unpadded with their 16-byte chunks permuted (Stage::SWZ), Q8K super-blocks
//     contiguous — in every layout both the cp.async writes of a quarter warp and each thread's LDS.128 of its own row
//     are bank-conflict free;
//   * f16 input folds with packed HMNMX2; codes are produced with full-rate float ops (clamp, then a
//     round-toward-zero add of 2^23 leaves floor(v) in the mantissa) instead of F2I on the XU pipe;
//   * launch shape (launch_quant, quant_rows_oneshot): short-lived CTAs that request their tile(s) up front — one
//     128-row tile from f32 input, two 64-row tiles from 16-bit input (rows are half as long there, and the bytes in
//     flight per SM, not the instruction count, bound the encoders) — encode tile k while tile k+1 arrives, stage the
//     packed tiles in shared memory (tiles 1.. on top of input stages already consumed) and write them with cooperative
//     16-byte stores, or, for Q8K, as ONE 1-D bulk async store (`cp.async.bulk.global.shared::cta`, SASS UBLKCP).
#include <cstdlib>
#include <type_traits>

#include "ggq_common.cuh"
#include "ggq_kernels.h"

namespace ggq {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr float F32_MAX = 3.40282347e+38f;


// bf16 pair -> two f32: shift / mask (ALU pipe).  Blackwell's mixed-precision FMA (`fma.rn.f32.bf16`, SASS FHFMA.BF16:
// x * 1 + (-0) widens either half of the register exactly, on the FP32 pipe) was measured as a replacement: bit-identical
// output, no gain (Q5_0 -0.7 point, Q8K -1.2, Q5_1 +0.4; profiles/r02_quant_bf16_widen_ab.txt) — the ALU pipe is not
// what holds the bf16 encoders back.
__device__ __forceinline__ float2 widen_bf16_pair(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u)); }

// ---- a row of 32 elements, widened to f32 (lib.rs:66-69, 82-84); `raw` keeps the f16 pairs --------
template <class FT> struct Row;
template <> struct Row<F32> {
    float x[32];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int k = 0; k < 32; k++) x[k] = 0.0f;
    }
    __device__ __forceinline__ void load(const uint8_t *s, uint32_t = 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 v = *reinterpret_cast<const float4 *>(s + 16 * i);
            x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
        }
    }
};
template <> struct Row<BF16> {
    float x[32];
    uint32_t raw[16];  // bf16 pairs (x[2k], x[2k+1])
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int k = 0; k < 32; k++) x[k] = 0.0f;
#pragma unroll
        for (int k = 0; k < 16; k++) raw[k] = 0u;
    }
    __device__ __forceinline__ void load(const uint8_t *s, uint32_t xm = 0) {  // xm: Stage::xmask of this row
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint4 v = *reinterpret_cast<const uint4 *>(s + ((16 * i) ^ xm));
            raw[4 * i] = v.x; raw[4 * i + 1] = v.y; raw[4 * i + 2] = v.z; raw[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const float2 f = widen_bf16_pair(raw[k]);
            x[2 * k] = f.x; x[2 * k + 1] = f.y;
        }
    }
};
template <> struct Row<F16> {
    float x[32];
    uint32_t raw[16];  // half2 pairs (x[2k], x[2k+1])
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int k = 0; k < 32; k++) x[k] = 0.0f;
#pragma unroll
        for (int k = 0; k < 16; k++) raw[k] = 0u;
    }
    __device__ __forceinline__ void load(const uint8_t *s, uint32_t xm = 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint4 v = *reinterpret_cast<const uint4 *>(s + ((16 * i) ^ xm));
            raw[4 * i] = v.x; raw[4 * i + 1] = v.y; raw[4 * i + 2] = v.z; raw[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&raw[k]));
            x[2 * k] = f.x; x[2 * k + 1] = f.y;
        }
    }
};

// ---- staging layout --------------------------------------------------------------------------------------------------
// 32-element rows of a tile in shared memory.  Legacy blocks (one thread per row): 128-byte f32 rows padded by 16 bytes,
// 64-byte f16 / bf16 rows unpadded and chunk-permuted (SWZ below) — either way every thread's LDS.128 of its own row and
// every quarter warp's cp.async write is bank-conflict free.  Q8K (eight lanes per 256-element super-block, see
// Encoder<T_Q8K>): the eight rows of a super-block stay contiguous and each SUPER-BLOCK is padded by 64 bytes (its
// stride is 16 words mod 32), because there the eight lanes of a group read consecutive 8- / 16-byte pieces.
template <uint32_t T, class FT> struct Stage {
    static constexpr bool SB = (T == T_Q8K);
    // 64-byte rows (16-bit input) are stored UNPADDED with their four 16-byte chunks permuted: chunk c of row r sits at
    // chunk c ^ ((r >> 1) & 3).  A quarter warp's LDS.128 (eight consecutive rows, the same logical chunk) then touches
    // eight different 16-byte bank groups, and its eight cp.async writes (two rows x four chunks) 128 contiguous bytes —
    // with rows padded to 80 bytes the first and the last of those eight shared a bank group: a two-way conflict on
    // every asynchronous write, which, with the 25 % larger stage, held these kernels at 85-90 % of the copy peak.
    static constexpr bool SWZ = !SB && FT::SIZE == 2;
    static constexpr int ROW_BYTES = 32 * FT::SIZE;
    static constexpr int ROW_STRIDE = (SB || SWZ) ? ROW_BYTES : ROW_BYTES + 16;
    static constexpr int SB_PAD = SB ? 64 : 0;
    static __host__ __device__ constexpr int bytes(int rows) { return rows * ROW_STRIDE + rows / 8 * SB_PAD; }
    static __device__ __forceinline__ uint32_t row_off(uint32_t r) { return r * ROW_STRIDE + (r >> 3) * SB_PAD; }
    // XOR mask for a byte offset inside row r (0 when the layout is not permuted)
    static __device__ __forceinline__ uint32_t xmask(uint32_t r) { return SWZ ? ((r >> 1) & 3u) << 4 : 0u; }
};

// Q8K: lane j (0..7) of a group owns the four elements 32k + 4j .. 32k + 4j + 3 of every row k (0..7) of its super-block,
// i.e. x[4k + i] = element 32k + 4j + i.  A group then reads 64 (f16) or 128 (f32) consecutive bytes per k, and — what
// matters — writes its codes as consecutive words: with one row per lane the sixteen 2-byte stores of a lane went to
// addresses 32 bytes apart across the group, an 8-way bank conflict (ncu: 58 % of all shared-memory wavefronts were
// conflicts, mio_throttle / short_scoreboard the top stalls, 13.3 instructions per element).
template <class FT> __device__ __forceinline__ void load_interleaved(Row<FT> &r, const uint8_t *sb, int j);
template <> __device__ __forceinline__ void load_interleaved<F32>(Row<F32> &r, const uint8_t *sb, int j) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float4 v = *reinterpret_cast<const float4 *>(sb + 128 * k + 16 * j);
        r.x[4 * k] = v.x; r.x[4 * k + 1] = v.y; r.x[4 * k + 2] = v.z; r.x[4 * k + 3] = v.w;
    }
}
template <> __device__ __forceinline__ void load_interleaved<F16>(Row<F16> &r, const uint8_t *sb, int j) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint2 v = *reinterpret_cast<const uint2 *>(sb + 64 * k + 8 * j);
        r.raw[2 * k] = v.x; r.raw[2 * k + 1] = v.y;
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&r.raw[k]));
        r.x[2 * k] = f.x; r.x[2 * k + 1] = f.y;
    }
}
template <> __device__ __forceinline__ void load_interleaved<BF16>(Row<BF16> &r, const uint8_t *sb, int j) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint2 v = *reinterpret_cast<const uint2 *>(sb + 64 * k + 8 * j);
        r.raw[2 * k] = v.x; r.raw[2 * k + 1] = v.y;
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const float2 f = widen_bf16_pair(r.raw[k]);
        r.x[2 * k] = f.x; r.x[2 * k + 1] = f.y;
    }
}

// ---- folds of structs.rs:91-107 over the 32 elements of one row -----------------------------------
// All three are max/min folds that ignore NaN; max and min are exact, so evaluating them on packed f16
// pairs gives the same value as the reference's f32 fold.
// max_abs (structs.rs:91-94): fold acc.max(|x|) from 0
template <class FT> __device__ __forceinline__ float row_max_abs(const Row<FT> &r) {
    float a = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; i++) a = fmaxf(a, fabsf(r.x[i]));
    return a;
}
template <> __device__ __forceinline__ float row_max_abs<F16>(const Row<F16> &r) {
    __half2 a = __floats2half2_rn(0.0f, 0.0f);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const uint32_t m = r.raw[k] & 0x7FFF7FFFu;
        a = __hmax2(a, *reinterpret_cast<const __half2 *>(&m));
    }
    return fmaxf(__low2float(a), __high2float(a));
}
// largest and smallest element with a virtual 0 (the fold's start value): P >= 0 >= N
template <class FT> __device__ __forceinline__ void row_pos_neg(const Row<FT> &r, float &P, float &N) {
    float p = 0.0f, n = 0.0f;
#pragma unroll
    for (int i = 0; i < 32; i++) { p = fmaxf(p, r.x[i]); n = fminf(n, r.x[i]); }
    P = p; N = n;
}
template <> __device__ __forceinline__ void row_pos_neg<F16>(const Row<F16> &r, float &P, float &N) {
    __half2 p = __floats2half2_rn(0.0f, 0.0f), n = p;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const __half2 v = *reinterpret_cast<const __half2 *>(&r.raw[k]);
        p = __hmax2(p, v);
        n = __hmin2(n, v);
    }
    P = fmaxf(__low2float(p), __high2float(p));
    N = fminf(__low2float(n), __high2float(n));
}
// min_max (structs.rs:102-107): folds from (f32::MAX, f32::MIN); NaN dropped
template <class FT> __device__ __forceinline__ void row_min_max(const Row<FT> &r, float &mn, float &mx) {
    float lo = F32_MAX, hi = -F32_MAX;
#pragma unroll
    for (int i = 0; i < 32; i++) { lo = fminf(lo, r.x[i]); hi = fmaxf(hi, r.x[i]); }
    mn = lo; mx = hi;
}
template <> __device__ __forceinline__ void row_min_max<F16>(const Row<F16> &r, float &mn, float &mx) {
    const uint32_t pinf = 0x7C007C00u, ninf = 0xFC00FC00u;
    __half2 lo = *reinterpret_cast<const __half2 *>(&pinf), hi = *reinterpret_cast<const __half2 *>(&ninf);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const __half2 v = *reinterpret_cast<const __half2 *>(&r.raw[k]);
        lo = __hmin2(lo, v);
        hi = __hmax2(hi, v);
    }
    mn = fminf(F32_MAX, fminf(__low2float(lo), __high2float(lo)));
    mx = fmaxf(-F32_MAX, fmaxf(__low2float(hi), __high2float(hi)));
}

// bf16 input: the same packed folds with the bf16x2 min/max instructions (NaN operands are dropped)
__device__ __forceinline__ float bf_lo(__nv_bfloat162 v) { return __uint_as_float((*reinterpret_cast<uint32_t *>(&v)) << 16); }
__device__ __forceinline__ float bf_hi(__nv_bfloat162 v) { return __uint_as_float((*reinterpret_cast<uint32_t *>(&v)) & 0xFFFF0000u); }
__device__ __forceinline__ __nv_bfloat162 bf_pair(uint32_t w) { return *reinterpret_cast<const __nv_bfloat162 *>(&w); }
template <> __device__ __forceinline__ float row_max_abs<BF16>(const Row<BF16> &r) {
    __nv_bfloat162 a = bf_pair(0u);
#pragma unroll
    for (int k = 0; k < 16; k++) a = __hmax2(a, bf_pair(r.raw[k] & 0x7FFF7FFFu));
    return fmaxf(bf_lo(a), bf_hi(a));
}
template <> __device__ __forceinline__ void row_pos_neg<BF16>(const Row<BF16> &r, float &P, float &N) {
    __nv_bfloat162 p = bf_pair(0u), n = p;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const __nv_bfloat162 v = bf_pair(r.raw[k]);
        p = __hmax2(p, v);
        n = __hmin2(n, v);
    }
    P = fmaxf(bf_lo(p), bf_hi(p));
    N = fminf(bf_lo(n), bf_hi(n));
}
template <> __device__ __forceinline__ void row_min_max<BF16>(const Row<BF16> &r, float &mn, float &mx) {
    __nv_bfloat162 lo = bf_pair(0x7F807F80u), hi = bf_pair(0xFF80FF80u);  // +inf, -inf
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const __nv_bfloat162 v = bf_pair(r.raw[k]);
        lo = __hmin2(lo, v);
        hi = __hmax2(hi, v);
    }
    mn = fminf(F32_MAX, fminf(bf_lo(lo), bf_hi(lo)));
    mx = fmaxf(-F32_MAX, fmaxf(bf_lo(hi), bf_hi(hi)));
}

// max_by_abs (structs.rs:96-100) over G lanes x 32 elements: the FIRST x with strictly greatest |x|.
// With P = max(0, x...) and N = min(0, x...) the winner is P if P > -N, N if -N > P; only when both
// +a and -a occur (P == -N != 0) does the order matter, and then the first of them wins.
// `INTERLEAVED`: r.x[i] is element 32 (i / 4) + 4 j + (i % 4) of the group's 256 (Q8K), else element 32 j + i.
template <int G, class FT, bool INTERLEAVED = false> __device__ __forceinline__ float max_by_abs_from(const Row<FT> &r, int j, float P, float N) {
    auto index_of = [j](int i) { return INTERLEAVED ? 32 * (i >> 2) + 4 * j + (i & 3) : 32 * j + i; };
#pragma unroll
    for (int m = 1; m < G; m <<= 1) {
        P = fmaxf(P, __shfl_xor_sync(FULL, P, m));
        N = fminf(N, __shfl_xor_sync(FULL, N, m));
    }
    const float a = -N;
    float res = (P > a) ? P : ((a > P) ? N : 0.0f);
    // tie between +P and -P somewhere in the block: the first |x| == P in index order decides the sign
    const bool tie = (P == a) && (P != 0.0f);
    if (G == 1 ? tie : __any_sync(FULL, tie)) {  // G > 1: every lane of the warp runs the shuffles
        // one descending pass: the last assignment comes from the lowest matching index (indices ascend with i in both
        // layouts).  With 16-bit input — bf16 above all, 8 significant bits — a row holding both +max and -max is common
        // (several per cent of random rows, i.e. most warps), so this path is not a cold one.
        int first = 32 * G;
        float val = 0.0f;
#pragma unroll
        for (int i = 31; i >= 0; i--) {
            const bool hit = fabsf(r.x[i]) == P;
            if (G > 1) first = hit ? index_of(i) : first;
            val = hit ? r.x[i] : val;
        }
#pragma unroll
        for (int m = 1; m < G; m <<= 1) {
            const int of = __shfl_xor_sync(FULL, first, m);
            const float ov = __shfl_xor_sync(FULL, val, m);
            if (of < first) { first = of; val = ov; }
        }
        if (tie) res = val;
    }
    return res;
}

template <int G, class FT> __device__ __forceinline__ float block_max_by_abs(const Row<FT> &r, int j) {
    float P, N;
    row_pos_neg<FT>(r, P, N);
    return max_by_abs_from<G, FT>(r, j, P, N);
}
// row_pos_neg that also reports whether the row holds a NaN, at no extra cost for 16-bit rows: the max fold
// propagates NaN (`max.NaN.f16x2`), so P comes out NaN exactly when some element is (and is then unusable: the
// caller recomputes with row_pos_neg).  f32 rows have no such fold and always report "unknown" (true).
template <class FT> __device__ __forceinline__ bool row_pos_neg_nan(const Row<FT> &r, float &P, float &N) {
    row_pos_neg<FT>(r, P, N);
    return true;
}
template <> __device__ __forceinline__ bool row_pos_neg_nan<F16>(const Row<F16> &r, float &P, float &N) {
    __half2 p = __floats2half2_rn(0.0f, 0.0f), n = p;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const __half2 v = *reinterpret_cast<const __half2 *>(&r.raw[k]);
        p = __hmax2_nan(p, v);
        n = __hmin2(n, v);
    }
    const float pl = __low2float(p), ph = __high2float(p);
    P = fmaxf(pl, ph);
    N = fminf(__low2float(n), __high2float(n));
    return (pl != pl) || (ph != ph);
}
template <> __device__ __forceinline__ bool row_pos_neg_nan<BF16>(const Row<BF16> &r, float &P, float &N) {
    __nv_bfloat162 p = bf_pair(0u), n = p;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const __nv_bfloat162 v = bf_pair(r.raw[k]);
        p = __hmax2_nan(p, v);
        n = __hmin2(n, v);
    }
    const float pl = bf_lo(p), ph = bf_hi(p);
    P = fmaxf(pl, ph);
    N = fminf(bf_lo(n), bf_hi(n));
    return (pl != pl) || (ph != ph);
}

// min of the row as the reference's strict-compare fold reports it: when the minimum is a zero, the
// FIRST zero in index order keeps its sign (oracle/ggq_oracle.c min_max()).
template <class FT> __device__ __forceinline__ float fix_zero_min(const Row<FT> &r, float mn) {
    if (mn != 0.0f) return mn;
    float z = mn;
#pragma unroll
    for (int i = 31; i >= 0; i--) z = (r.x[i] == 0.0f) ? r.x[i] : z;
    return z;
}

// floor(v) for 0 <= v < 2^23 left in the low mantissa bits (RZ add never rounds up across an integer)
__device__ __forceinline__ uint32_t floor_bits(float v) { return __float_as_uint(__fadd_rz(v, 8388608.0f)); }
// ---- packed FP32 (Blackwell FADD2 / FMUL2, `*.f32x2`): the kernel is issue-bound, and a packed instruction
// does the IEEE operation on the pair (x[2k], x[2k+1]) in one issue slot.  One rule: a packed multiply must
// never feed a packed add — ptxas 12.9 contracts that pair into a single-rounding FFMA2 even with
// --fmad=false (see quant_k.cu; the Makefile rejects objects containing FFMA2).  The encoders therefore clamp
// between `x * recip` and the `+ 8.5` / `+ 0.5` (bounds shifted by the constant: identical results).
__device__ __forceinline__ float2 pair_of(const float *x, int k) { return make_float2(x[2 * k], x[2 * k + 1]); }
__device__ __forceinline__ float2 mul2(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
__device__ __forceinline__ float2 add2(float2 a, float c) { return __fadd2_rn(a, make_float2(c, c)); }
__device__ __forceinline__ float2 add_after_mul(float2 a, float c) { return make_float2(__fadd_rn(a.x, c), __fadd_rn(a.y, c)); }
__device__ __forceinline__ float2 clamp2(float2 v, float lo, float hi) { return make_float2(fminf(fmaxf(v.x, lo), hi), fminf(fmaxf(v.y, lo), hi)); }
// floor(v) of both lanes left in the low mantissa bits (see floor_bits)
__device__ __forceinline__ void floor_bits2(float2 v, uint32_t &b0, uint32_t &b1) {
    const float2 t = __fadd2_rz(v, make_float2(8388608.0f, 8388608.0f));
    b0 = __float_as_uint(t.x);
    b1 = __float_as_uint(t.y);
}
// Rust `v.round() as i8` as an int: half away from zero, NaN -> 0, +-inf saturate at int range
__device__ __forceinline__ int round_half_away(float v) {
    const float h = __uint_as_float((__float_as_uint(v) & 0x80000000u) | 0x3F000000u);  // copysign(0.5, v)
    return __float2int_rz(__fadd_rz(v, h));
}
// bytes [sat8(a), sat8(b), sat8(c), sat8(d)] (little-endian) from four ints
__device__ __forceinline__ uint32_t pack_sat_s8(int a, int b, int c, int d) {
    uint32_t t, w;
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(d), "r"(c), "r"(0));
    asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(b), "r"(a), "r"(t));
    return w;
}
// low bytes of four registers -> one word
__device__ __forceinline__ uint32_t gather_b0(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

__device__ __forceinline__ void sts16(uint8_t *p, uint32_t v) { *reinterpret_cast<uint16_t *>(p) = (uint16_t)v; }
// store NW words at a 2-byte aligned shared address
template <int NW> __device__ __forceinline__ void sts_words2(uint8_t *p, const uint32_t *w) {
#pragma unroll
    for (int i = 0; i < NW; i++) { sts16(p + 4 * i, w[i] & 0xFFFFu); sts16(p + 4 * i + 2, w[i] >> 16); }
}
template <int NW> __device__ __forceinline__ void sts_words4(uint8_t *p, const uint32_t *w) {
#pragma unroll
    for (int i = 0; i < NW; i++) *reinterpret_cast<uint32_t *>(p + 4 * i) = w[i];
}

// 16 payload bytes of Q4_x / low nibbles of Q5_x: byte i = code[i] | code[i+16] << 4, codes < 16 in
// the low byte of c[] (upper bits of c[] are don't-care: only byte 0 of the sum is gathered)
__device__ __forceinline__ void nibble_bytes(const uint32_t *c, uint32_t *w) {
    uint32_t b[16];
#pragma unroll
    for (int i = 0; i < 16; i++) b[i] = c[i] + (c[i + 16] << 4);
#pragma unroll
    for (int k = 0; k < 4; k++) w[k] = gather_b0(b[4 * k], b[4 * k + 1], b[4 * k + 2], b[4 * k + 3]);
}

// ---- encoders: one thread, one 32-element row -------------------------------------------------------
template <uint32_t T> struct Encoder;

// The 4- and 5-bit encoders are issue-bound from 16-bit input, and two of their ~8.5 instructions per element are
// the clamp around `v = x * recip` (or `(x - min) * recip`).  When delta is an ordinary number — 2^-100 <= |d| <=
// 2^100, so that neither d, 1/d nor any product under- or overflows — one side of that clamp can never act:
//   symmetric types   |x| <= |max|  =>  |v| <= N (1 + 3*2^-24), N = 8 / 16: v + N.5 stays above 0.49, only the top
//                     (x = -max: v + N.5 = 2N.5) needs the clamp;
//   affine types      0 <= x - min <= max - min  =>  0 <= v <= L (1 + 3*2^-24), L = 15 / 31: v + 0.5 < L + 1, only
//                     the bottom clamp is kept, and only because it is what maps a NaN element to code 0.
// Every other row (denormal / huge / infinite delta) takes the two-sided clamp below, as before.
__device__ __forceinline__ bool ordinary_scale(float d) {
    const float a = fabsf(d);
    return a >= 7.888609052210118e-31f && a <= 1.2676506002282294e30f;  // 2^-100 .. 2^100 (false for NaN)
}

// q4_0.rs:23-44
template <> struct Encoder<T_Q4_0> {
    static constexpr int G = 1;
    template <class FT> static __device__ __forceinline__ void run(const Row<FT> &r, int, uint8_t *blk) {
        const float mx = block_max_by_abs<1, FT>(r, 0);
        uint32_t w[4] = {0, 0, 0, 0};
        uint32_t d16 = 0;
        if (mx != 0.0f) {
            const float d = __fmul_rn(mx, -0.125f);  // == mx / -8 (power of two: same real value, same rounding)
            const float rc = __fdiv_rn(1.0f, d);
            d16 = f2h(d);
            uint32_t c[32];
            if (ordinary_scale(d)) {  // v >= -8.000002: the lower clamp cannot act (see ordinary_scale)
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const float2 p = mul2(pair_of(r.x, k), rc);
                    floor_bits2(add2(make_float2(fminf(p.x, 6.5f), fminf(p.y, 6.5f)), 8.5f), c[2 * k], c[2 * k + 1]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    // (x*recip + 8.5).min(15.) as u8 : NaN -> 15 (fminf drops it), negatives -> 0
                    // the clamp moves in front of the add (bounds shifted by 8.5, same results incl. NaN -> 15), so the
                    // add no longer follows the packed multiply directly and can be packed as well
                    const float2 p = mul2(pair_of(r.x, k), rc);
                    const float2 q = make_float2(fmaxf(fminf(p.x, 6.5f), -8.5f), fmaxf(fminf(p.y, 6.5f), -8.5f));
                    floor_bits2(add2(q, 8.5f), c[2 * k], c[2 * k + 1]);
                }
            }
            nibble_bytes(c, w);
        }
        sts16(blk, d16);
        sts_words2<4>(blk + 2, w);
    }
};

// q4_1.rs:23-47
template <> struct Encoder<T_Q4_1> {
    static constexpr int G = 1;
    template <class FT> static __device__ __forceinline__ void run(const Row<FT> &r, int, uint8_t *blk) {
        float mn, mx;
        row_min_max<FT>(r, mn, mx);
        mn = fix_zero_min<FT>(r, mn);
        uint32_t w[4] = {0, 0, 0, 0};
        uint32_t d16 = 0;
        if (mn != mx) {
            const float d = __fdiv_rn(__fsub_rn(mx, mn), 15.0f), rc = __fdiv_rn(1.0f, d);
            d16 = f2h(d);
            uint32_t c[32];
            if (ordinary_scale(d)) {  // v <= 15.000003: the upper clamp cannot act; the lower one maps NaN -> 0
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const float2 p = mul2(add2(pair_of(r.x, k), -mn), rc);
                    floor_bits2(add2(make_float2(fmaxf(p.x, -0.5f), fmaxf(p.y, -0.5f)), 0.5f), c[2 * k], c[2 * k + 1]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    // (((x - min)*recip + 0.5) as u8).min(15) : NaN -> 0
                    const float2 p = mul2(add2(pair_of(r.x, k), -mn), rc);
                    floor_bits2(add2(clamp2(p, -0.5f, 14.5f), 0.5f), c[2 * k], c[2 * k + 1]);  // clamp before the add, see Q4_0
                }
            }
            nibble_bytes(c, w);
        }
        *reinterpret_cast<uint32_t *>(blk) = d16 | ((uint32_t)f2h(mn) << 16);
        sts_words4<4>(blk + 4, w);
    }
};

// 5-bit codes (< 32, low byte of c[]) -> qh (bit i = bit 4 of code i) and the 16 nibble bytes, done on
// words of four codes: low nibbles by mask, the four high bits of a word are compressed to a nibble
// with one multiply ((h0 + h1<<8 + h2<<16 + h3<<24) * 0x01020408 puts h0..h3 at bits 24..27, no carries).
__device__ __forceinline__ void pack5(const uint32_t *c, uint32_t &qh, uint32_t *w) {
    uint32_t lo[8];
    qh = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t W = gather_b0(c[4 * k], c[4 * k + 1], c[4 * k + 2], c[4 * k + 3]);
        lo[k] = W & 0x0F0F0F0Fu;
        const uint32_t hb = (W >> 4) & 0x01010101u;
        qh |= ((hb * 0x01020408u) >> 24) << (4 * k);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) w[j] = lo[j] | (lo[j + 4] << 4);  // byte i = code[i] & 15 | (code[i+16] & 15) << 4
}

// q5_0.rs:26-58
template <> struct Encoder<T_Q5_0> {
    static constexpr int G = 1;
    template <class FT> static __device__ __forceinline__ void run(const Row<FT> &r, int, uint8_t *blk) {
        float P, N;
        const bool maybe_nan = row_pos_neg_nan<FT>(r, P, N);
        if (FT::SIZE == 2 && maybe_nan) row_pos_neg<FT>(r, P, N);  // a NaN in the row spoiled P: fold again without it
        const float mx = max_by_abs_from<1, FT>(r, 0, P, N);
        uint32_t w[4] = {0, 0, 0, 0};
        uint32_t d16 = 0, qh = 0;
        if (mx != 0.0f) {
            const float d = __fmul_rn(mx, -0.0625f);  // == mx / -16
            const float rc = __fdiv_rn(1.0f, d);
            d16 = f2h(d);
            uint32_t c[32];
            if (!maybe_nan && ordinary_scale(d)) {
                // v >= -16.000004: the lower clamp could only act on a NaN element (NaN -> code 0), and there is none
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const float2 p = mul2(pair_of(r.x, k), rc);
                    floor_bits2(add2(make_float2(fminf(p.x, 14.5f), fminf(p.y, 14.5f)), 16.5f), c[2 * k], c[2 * k + 1]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    // ((x*recip + 16.5) as u8).min(31) : NaN -> 0
                    const float2 p = mul2(pair_of(r.x, k), rc);
                    floor_bits2(add2(clamp2(p, -16.5f, 14.5f), 16.5f), c[2 * k], c[2 * k + 1]);  // clamp before the add, see Q4_0
                }
            }
            pack5(c, qh, w);
        }
        sts16(blk, d16);
        sts16(blk + 2, qh & 0xFFFFu);
        sts16(blk + 4, qh >> 16);
        sts_words2<4>(blk + 6, w);
    }
};

// q5_1.rs:26-62
template <> struct Encoder<T_Q5_1> {
    static constexpr int G = 1;
    template <class FT> static __device__ __forceinline__ void run(const Row<FT> &r, int, uint8_t *blk) {
        float mn, mx;
        row_min_max<FT>(r, mn, mx);
        mn = fix_zero_min<FT>(r, mn);
        uint32_t w[4] = {0, 0, 0, 0};
        uint32_t d16 = 0, qh = 0;
        if (mn != mx) {
            const float d = __fdiv_rn(__fsub_rn(mx, mn), 31.0f), rc = __fdiv_rn(1.0f, d);
            d16 = f2h(d);
            uint32_t c[32];
            if (ordinary_scale(d)) {  // v <= 31.000006: the upper clamp cannot act; the lower one maps NaN -> 0
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const float2 p = mul2(add2(pair_of(r.x, k), -mn), rc);
                    floor_bits2(add2(make_float2(fmaxf(p.x, -0.5f), fmaxf(p.y, -0.5f)), 0.5f), c[2 * k], c[2 * k + 1]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const float2 p = mul2(add2(pair_of(r.x, k), -mn), rc);
                    floor_bits2(add2(clamp2(p, -0.5f, 30.5f), 0.5f), c[2 * k], c[2 * k + 1]);  // clamp before the add, see Q4_0
                }
            }
            pack5(c, qh, w);
        }
        *reinterpret_cast<uint32_t *>(blk) = d16 | ((uint32_t)f2h(mn) << 16);
        *reinterpret_cast<uint32_t *>(blk + 4) = qh;
        sts_words4<4>(blk + 8, w);
    }
};

// q8_0.rs:23-41 and q8_1.rs:28-55 (Q8_1 adds sum = f16(Σq as f32 * delta), delta unrounded)
template <uint32_t T, bool WITH_SUM> struct Encoder8 {
    static constexpr int G = 1;
    template <class FT> static __device__ __forceinline__ void run(const Row<FT> &r, int, uint8_t *blk) {
        const float amax = row_max_abs<FT>(r);
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint32_t d16 = 0, s16 = 0;
        if (amax != 0.0f) {
            const float d = __fdiv_rn(amax, 127.0f), rc = __fdiv_rn(1.0f, d);
            d16 = f2h(d);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const float2 a = mul2(pair_of(r.x, 2 * k), rc), b = mul2(pair_of(r.x, 2 * k + 1), rc);
                w[k] = pack_sat_s8(round_half_away(a.x), round_half_away(a.y), round_half_away(b.x), round_half_away(b.y));
            }
            if constexpr (WITH_SUM) {
                int s = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) s = __dp4a((int)w[k], 0x01010101, s);
                // sum as f32 * delta: the only NaN this can produce is 0 * inf (amax = inf makes delta inf, recip 0 and every
                // code 0).  The reference computes it on x86, whose invalid-operation result is the negative quiet NaN
                // 0xFFC00000 -> f16 0xFE00 through half's payload-keeping narrow; the GPU's would be 0x7FFF.
                const float sf = __fmul_rn((float)s, d);
                s16 = (sf != sf) ? 0xFE00u : f2h(sf);
            }
        }
        if constexpr (WITH_SUM) {
            *reinterpret_cast<uint32_t *>(blk) = d16 | (s16 << 16);
            sts_words4<8>(blk + 4, w);
        } else {
            sts16(blk, d16);
            sts_words2<8>(blk + 2, w);
        }
    }
};
template <> struct Encoder<T_Q8_0> : Encoder8<T_Q8_0, false> {};
template <> struct Encoder<T_Q8_1> : Encoder8<T_Q8_1, true> {};

// q8_k.rs:27-54 — reference layout {delta: f16, quants: [i8;256], sums: [i16;16]}, 290 bytes.
// 8 consecutive lanes own one super-block; lane j holds elements 32k + 4j .. + 3 of every row k (load_interleaved).
template <> struct Encoder<T_Q8K> {
    static constexpr int G = 8;
    template <class FT> static __device__ __forceinline__ void run(const Row<FT> &r, int j, uint8_t *blk) {
        float P, N;
        row_pos_neg<FT>(r, P, N);
        const float mx = max_by_abs_from<8, FT, true>(r, j, P, N);
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint32_t d16 = 0;
        int v[4] = {0, 0, 0, 0};  // per-row sums of this lane's codes, two 16-bit sums per register
        if (mx != 0.0f) {  // uniform over the group (not over the warp: no shuffles in here)
            const float d = __fdiv_rn(mx, -127.0f), rc = __fdiv_rn(1.0f, d);
            d16 = f2h(d);
            int q[32];
            float2 prod[16];
#pragma unroll
            for (int k = 0; k < 16; k++) prod[k] = mul2(pair_of(r.x, k), rc);
#pragma unroll
            for (int i = 0; i < 32; i++) {
                // (x*recip).round().min(127.) as i8 : NaN.round() is NaN and NaN.min(127.) is 127.
                // trunc(min(t, 127)) == min(trunc(t), 127) and fminf drops a NaN t, so no branch.
                const float p = (i & 1) ? prod[i >> 1].y : prod[i >> 1].x;
                const float h = __uint_as_float((__float_as_uint(p) & 0x80000000u) | 0x3F000000u);  // copysign(0.5, p)
                q[i] = __float2int_rz(fminf(__fadd_rz(p, h), 127.0f));
            }
#pragma unroll
            for (int k = 0; k < 8; k++) w[k] = pack_sat_s8(q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3]);
#pragma unroll
            for (int m = 0; m < 4; m++) v[m] = __dp4a((int)w[2 * m], 0x01010101, 0) + (__dp4a((int)w[2 * m + 1], 0x01010101, 0) << 16);
        }
        // sums[g] covers elements 16g .. 16g + 15 = row k = g / 2, lanes 0..3 (g even) or 4..7 (g odd): four-lane reductions
        // of the eight per-row sums (|sum| <= 16 * 128: the halves of a register never carry into each other beyond what the
        // unpacking below undoes)
#pragma unroll
        for (int m = 0; m < 4; m++) {
            v[m] += __shfl_xor_sync(FULL, v[m], 1);
            v[m] += __shfl_xor_sync(FULL, v[m], 2);
        }
        const int m4 = j & 3;  // this lane stores the sums of rows 2 m4 and 2 m4 + 1 for its half of each row
        const int vm = m4 == 0 ? v[0] : m4 == 1 ? v[1] : m4 == 2 ? v[2] : v[3];
        const int s_lo = (int)(short)(vm & 0xFFFF);
        const int s_hi = (vm - s_lo) >> 16;
        // codes: word k of lane j is bytes 2 + 32k + 4j .. + 3 of the block: consecutive words across the group
#pragma unroll
        for (int k = 0; k < 8; k++) {
            sts16(blk + 2 + 32 * k + 4 * j, w[k] & 0xFFFFu);
            sts16(blk + 2 + 32 * k + 4 * j + 2, w[k] >> 16);
        }
        const int half = j >> 2;  // sums[2 * (2 m4) + half] and sums[2 * (2 m4 + 1) + half]
        sts16(blk + 258 + 2 * (4 * m4 + half), (uint32_t)s_lo & 0xFFFFu);
        sts16(blk + 258 + 2 * (4 * m4 + 2 + half), (uint32_t)s_hi & 0xFFFFu);
        if (j == 0) sts16(blk, d16);
    }
};

// ---------------------------------------------------------------------------------------------
// Short-lived CTAs: stage K tiles of ROWS rows up front, encode them one after the other, write the packed
// blocks — no ring, no running pointers.  The overlap of loads, math and stores comes from the hardware
// scheduling several resident CTAs per SM, the structure that took the cast kernel from 85 % to 98 % of the copy
// peak; per row it also sheds the bookkeeping of a persistent pipeline (the ring this replaced: 73-92 %).  K > 1
// exists for 16-bit input: a thread's row is then only 64 bytes, and one row per resident thread (registers allow
// ~1 000 threads per SM) is ~64 KB in flight per SM in the best case — too little to cover HBM latency at full
// bandwidth.  With K tiles requested before the first is encoded the bytes in flight per thread multiply and tile
// k+1 keeps arriving while tile k is encoded.  launch_quant picks (ROWS, K, register cap) per (type, float side).
template <int N> __device__ __forceinline__ void cp_async_wait_upto(int pending) {  // wait until <= pending groups are outstanding
    if constexpr (N == 0) {
        cp_async_wait<0>();
    } else {
        if (pending >= N) cp_async_wait<N>();
        else cp_async_wait_upto<N - 1>(pending);
    }
}
template <uint32_t T, class FT, int ROWS, int K, int MINB>
__global__ void __launch_bounds__(ROWS, MINB)
quant_rows_oneshot(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, size_t nblocks) {
    using TR = BlockTraits<T>;
    using E = Encoder<T>;
    constexpr int RPB = TR::ELEMS / 32, TILE_BLOCKS = ROWS / RPB;
    using ST = Stage<T, FT>;
    constexpr int ROW_BYTES = 32 * FT::SIZE, CPR = ROW_BYTES / 16, ROWS_PER_PASS = ROWS / CPR;
    constexpr int IN_STAGE = ST::bytes(ROWS), OUT_BYTES = TILE_BLOCKS * TR::BYTES;
    constexpr uint32_t PASS_STRIDE = ST::bytes(ROWS_PER_PASS);
    static_assert(ROWS_PER_PASS % 8 == 0, "a pass of chunks covers whole super-blocks (and leaves the row permutation unchanged)");
    static_assert(OUT_BYTES % 16 == 0, "tile must be a whole number of 16-byte chunks");
    // One buffer: [packed tile 0][input stage 0][input stage 1]...  Packed tile k lives at k * OUT_BYTES, i.e. tiles
    // 1.. overwrite input stages 0..k-1 — every thread has taken its rows of those stages before anyone passes the
    // barrier that precedes tile k's encode — so the K packed tiles are contiguous for the final copy and cost the
    // space of one.  (k + 1) * OUT_BYTES <= IN_OFF + k * IN_STAGE because a packed tile is smaller than its input.
    constexpr int IN_OFF = (OUT_BYTES + 127) & ~127;
    static_assert(OUT_BYTES <= IN_STAGE, "packed tiles must fit into consumed input stages");
    __shared__ __align__(128) uint8_t smem[IN_OFF + K * IN_STAGE];
    uint8_t *const out_st = smem, *const in_st = smem + IN_OFF;
    const int tid = threadIdx.x;
    const size_t nrows = nblocks * RPB, ntiles = (nrows + ROWS - 1) / ROWS;
    const bool vec_in = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
    const uint32_t s_chunk0 = ST::row_off((uint32_t)(tid / CPR)) + (((uint32_t)(tid % CPR) * 16) ^ ST::xmask((uint32_t)(tid / CPR)));
    const uint32_t my_row_off = ST::row_off(ST::SB ? (uint32_t)(tid & ~7) : (uint32_t)tid);
    const uint32_t my_xm = ST::xmask((uint32_t)tid);
    pdl_launch_dependents();
    pdl_wait();
    const bool vec_out = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
    for (size_t t0 = (size_t)blockIdx.x * K; t0 < ntiles; t0 += (size_t)gridDim.x * K) {  // one iteration unless the grid was capped
        if (vec_in && vec_out && (t0 + K) * ROWS <= nrows) {
            // ---- K full tiles, 16-byte aligned on both sides (every CTA but the last of an aligned tensor): straight-line
            // code.  The general path below spends ~130 bookkeeping instructions per warp and tile on ragged-tile
            // predicates and alignment dispatch, a third of what the encoder itself needs.
            const uint8_t *g = src + t0 * ((size_t)ROWS * ROW_BYTES) + (size_t)tid * 16;
#pragma unroll
            for (int k = 0; k < K; k++) {
#pragma unroll
                for (int c = 0; c < CPR; c++)
                    cp_async16(in_st + k * IN_STAGE + s_chunk0 + c * PASS_STRIDE, g + (size_t)k * ROWS * ROW_BYTES + (size_t)c * ROWS * 16);
                cp_async_commit();
            }
#pragma unroll
            for (int k = 0; k < K; k++) {
                cp_async_wait_upto<K - 1>(K - 1 - k);
                __syncthreads();
                Row<FT> r;
                if constexpr (ST::SB) load_interleaved<FT>(r, in_st + k * IN_STAGE + my_row_off, tid % RPB);
                else r.load(in_st + k * IN_STAGE + my_row_off, my_xm);
                E::template run<FT>(r, tid % RPB, out_st + k * OUT_BYTES + (uint32_t)(tid / RPB) * TR::BYTES);
            }
            // Q8K (2.3 KB per packed tile): the K tiles leave as ONE bulk async store issued by thread 0
            // (`cp.async.bulk.global.shared::cta`, SASS UBLKCP) — +3 points over cooperative 16-byte stores; for the
            // 1.1-2.3 KB tiles of the 32-element types the cooperative stores are 0.3-0.8 points ahead
            // (profiles/r02_quant_legacy16_sweep.txt)
#ifndef GGQ_ONESHOT_BULK_OUT
#define GGQ_ONESHOT_BULK_OUT (T == T_Q8K)
#endif
            if constexpr (GGQ_ONESHOT_BULK_OUT) {
                fence_proxy_async_smem();  // generic-proxy writes -> async-proxy reads
                __syncthreads();
                if (tid == 0) {
                    bulk_s2g(dst + t0 * (size_t)OUT_BYTES, out_st, (uint32_t)(K * OUT_BYTES));
                    bulk_commit();
                    bulk_wait_read<0>();  // shared memory must outlive the read
                }
            } else {
                __syncthreads();
                constexpr int NV = K * OUT_BYTES / 16;
                uint4 *gd = reinterpret_cast<uint4 *>(dst + t0 * (size_t)OUT_BYTES);
#pragma unroll
                for (int i = tid; i < NV; i += ROWS) gd[i] = reinterpret_cast<const uint4 *>(out_st)[i];
            }
            __syncthreads();  // only matters when the loop runs again
            continue;
        }
        // ---- request all K tiles ----
#pragma unroll
        for (int k = 0; k < K; k++) {
            const size_t t = t0 + k;
            if (t < ntiles) {
                const size_t rem = nrows - t * ROWS;
                const int rows = (int)min((size_t)ROWS, rem);
                const uint8_t *g = src + t * ((size_t)ROWS * ROW_BYTES);
                uint8_t *st = in_st + k * IN_STAGE;
                if (vec_in) {
#pragma unroll
                    for (int c = 0; c < CPR; c++)
                        if (rows == ROWS || tid / CPR + c * ROWS_PER_PASS < rows)
                            cp_async16(st + s_chunk0 + c * PASS_STRIDE, g + (size_t)tid * 16 + (size_t)c * ROWS * 16);
                } else {  // source not 16-byte aligned: element-granular staging
                    using RAW = typename FT::raw;
                    const RAW *ge = reinterpret_cast<const RAW *>(g);
                    for (int e = tid; e < rows * 32; e += ROWS)
                        *reinterpret_cast<RAW *>(st + ST::row_off(e / 32) + (((uint32_t)(e % 32) * FT::SIZE) ^ ST::xmask(e / 32))) = ge[e];
                }
            }
            cp_async_commit();  // one group per tile, empty or not
        }
        // ---- encode tile k while tiles k+1.. are still arriving ----
        size_t rows_total = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const size_t t = t0 + k;
            if (t < ntiles) {  // CTA-uniform
                const int rows = (int)min((size_t)ROWS, nrows - t * ROWS);
                cp_async_wait_upto<K - 1>(K - 1 - k);
                __syncthreads();
                if (RPB > 1 || tid < rows) {  // G > 1 encoders shuffle: every lane of the warp takes part
                    Row<FT> r;
                    if (tid >= rows) r.zero();
                    else if constexpr (ST::SB) load_interleaved<FT>(r, in_st + k * IN_STAGE + my_row_off, tid % RPB);
                    else r.load(in_st + k * IN_STAGE + my_row_off, my_xm);
                    E::template run<FT>(r, tid % RPB, out_st + k * OUT_BYTES + (uint32_t)(tid / RPB) * TR::BYTES);
                }
                rows_total += rows;
            }
        }
        __syncthreads();
        // the K tiles are consecutive in dst: one cooperative copy
        cta_copy_s2g(dst + t0 * (size_t)OUT_BYTES, out_st, (uint32_t)(rows_total / RPB) * TR::BYTES, tid, ROWS);
        __syncthreads();  // only matters when the loop runs again
    }
}

template <uint32_t T, class FT, int ROWS, int K, int MINB>
static cudaError_t launch_quant_oneshot(const void *src, void *dst, size_t nblocks, cudaStream_t stream) {
    constexpr int RPB = BlockTraits<T>::ELEMS / 32, TILE_BLOCKS = ROWS / RPB;
    size_t grid = (nblocks + (size_t)TILE_BLOCKS * K - 1) / ((size_t)TILE_BLOCKS * K);
    if (grid > 0x7FFFFFFFull) grid = 0x7FFFFFFFull;
    return launch_pdl(quant_rows_oneshot<T, FT, ROWS, K, MINB>, (unsigned)grid, ROWS, 0, stream, static_cast<const uint8_t *>(src), static_cast<uint8_t *>(dst), nblocks);
}

template <uint32_t T, class FT>
static cudaError_t launch_quant(const void *src, void *dst, size_t nblocks, cudaStream_t stream, DevInfo dev) {
    using TR = BlockTraits<T>;
    // measured (tools/codec_sweep.py, 58.7 M elements; profiles/r01_quant_oneshot_sweep.txt, r02_quant_legacy16_sweep.txt):
    //   f32 input: one 128-row tile per CTA runs at 100-102 % of the copy peak (a persistent ring: 89-92 %);
    //   16-bit input: two 64-row tiles per CTA, both requested up front (K = 3, 96- / 128-row tiles 1-3 points behind,
    //   32-row tiles 10; one tile per CTA or a ring 10-15).
    constexpr bool ONESHOT_F32 = (T != T_Q8K) && std::is_same<FT, F32>::value;
    constexpr bool ONESHOT_16 = (T != T_Q8K) && !std::is_same<FT, F32>::value;
    if constexpr (ONESHOT_F32) {
        return launch_quant_oneshot<T, FT, 128, 1, 1>(src, dst, nblocks, stream);
    } else if constexpr (ONESHOT_16) {
        // 64-byte rows unpadded and chunk-permuted (Stage::SWZ), packed tiles aliased onto consumed input stages: 9.3-10.5 KB
        // per CTA instead of 12.5, and every quarter-warp cp.async lands in 128 contiguous bytes (the 80-byte padded rows made
        // each of them a two-way bank conflict): 85-90 % -> 93-99.5 % of the measured copy peak.  The register cap is per type
        // (profiles/r02_quant_legacy16_sweep.txt): none for Q4_0 / Q4_1 / Q8_0 / Q8_1 (70-72 registers, 14 CTAs per SM), 64 /
        // 56 / 48 where more resident CTAs beat the compiler's freer schedule.
#ifdef GGQ_QL16_K
        return launch_quant_oneshot<T, FT, 64, GGQ_QL16_K, GGQ_QL16_MINB>(src, dst, nblocks, stream);
#else
        constexpr bool BF = std::is_same<FT, BF16>::value;
        constexpr int MINB = T == T_Q5_0 ? (BF ? 20 : 16) : T == T_Q5_1 ? 20 : (T == T_Q4_0 && BF) ? 18 : 1;
        return launch_quant_oneshot<T, FT, 64, 2, MINB>(src, dst, nblocks, stream);
#endif
    } else {  // Q8K
        // Eight interleaved lanes per super-block (Encoder<T_Q8K>).  Short-lived two-tile CTAs for every float side: from f32
        // 102-103 % of the copy peak (a persistent ring: 99 %); from 16-bit input 97-99 % with the registers capped (56, no
        // spills; left alone the compiler takes 118: 8 CTAs per SM and 74 %), the packed tiles on top of consumed input stages
        // (11.6 KB per CTA) and the bulk store — the ring that shipped before reached 83 / 77 % for f16 / bf16
        // (profiles/r02_quant_q8k_sweep.txt, r02_quant_legacy16_sweep.txt).  Round 1's one-row-per-lane mapping: 73 / 72 / 90 %.
        if constexpr (std::is_same<FT, F32>::value) return launch_quant_oneshot<T, FT, 64, 2, 1>(src, dst, nblocks, stream);
#ifdef GGQ_Q8K16_ONESHOT_K
        return launch_quant_oneshot<T, FT, 64, GGQ_Q8K16_ONESHOT_K, GGQ_Q8K16_MINB>(src, dst, nblocks, stream);
#else
        return launch_quant_oneshot<T, FT, 64, 2, 16>(src, dst, nblocks, stream);
#endif
    }
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

// Persistent CTAs over contiguous tiles of 8192 elements (each CTA reads/writes one contiguous 16-32 KB
// range at a time: better DRAM row locality than a grid-wide stride).  A chunk is V elements with the
// WIDER side exactly 16 bytes, so every warp access on either side is a contiguous run (no half-filled
// sectors); U independent loads per thread are in flight before the first store.
constexpr int CAST_THREADS = 256, CAST_TILE = 8192;
// 16-bit <-> 16-bit casts, one word = two elements.  Without a NaN in the pair the crate's "widen, then narrow RNE"
// is what the packed hardware conversions do; a pair holding a NaN takes the exact per-element route (payload
// kept, quiet bit set).  `ok` is false when either half is a NaN.
template <class ST, class DT> __device__ __forceinline__ uint32_t cast_pair16(uint32_t w, bool &ok);
template <> __device__ __forceinline__ uint32_t cast_pair16<F16, BF16>(uint32_t w, bool &ok) {
    const __half2 h = *reinterpret_cast<const __half2 *>(&w);
    ok = __hbeq2(h, h);
    const float2 f = __half22float2(h);
    const __nv_bfloat162 b = __floats2bfloat162_rn(f.x, f.y);
    return *reinterpret_cast<const uint32_t *>(&b);
}
template <> __device__ __forceinline__ uint32_t cast_pair16<BF16, F16>(uint32_t w, bool &ok) {
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162 *>(&w);
    ok = __hbeq2(b, b);
    const __half2 h = __floats2half2_rn(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
    return *reinterpret_cast<const uint32_t *>(&h);
}
template <class ST, class DT> struct IsPair16 { static constexpr bool value = false; };
template <> struct IsPair16<F16, BF16> { static constexpr bool value = true; };
template <> struct IsPair16<BF16, F16> { static constexpr bool value = true; };
template <int BYTES> struct VecOf;
template <> struct VecOf<16> { using type = uint4; };
template <> struct VecOf<8> { using type = uint2; };
template <class ST, class DT>
__global__ void __launch_bounds__(CAST_THREADS) cast_kernel(const typename ST::raw *__restrict__ src, typename DT::raw *__restrict__ dst, size_t n) {
    using SR = typename ST::raw;
    using DR = typename DT::raw;
    constexpr int WIDE = sizeof(SR) > sizeof(DR) ? sizeof(SR) : sizeof(DR);
    constexpr int V = 16 / WIDE;                       // elements per chunk: 8 (16-bit <-> 16-bit) or 4
    constexpr int U = CAST_TILE / (CAST_THREADS * V);  // chunks per thread per tile: 4 or 8
    using SV = typename VecOf<V * sizeof(SR)>::type;
    using DV = typename VecOf<V * sizeof(DR)>::type;
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
    pdl_launch_dependents();
    pdl_wait();
    const size_t full_tiles = n / CAST_TILE;
    if (vec) {
        for (size_t t = blockIdx.x; t < full_tiles; t += gridDim.x) {
            const SR *sp = src + t * CAST_TILE + threadIdx.x * V;
            DR *dp = dst + t * CAST_TILE + threadIdx.x * V;
            alignas(16) SR in[U][V];
#pragma unroll
            for (int u = 0; u < U; u++) *reinterpret_cast<SV *>(in[u]) = __ldg(reinterpret_cast<const SV *>(sp + u * CAST_THREADS * V));
#pragma unroll
            for (int u = 0; u < U; u++) {
                alignas(16) DR out[V];
                if constexpr (IsPair16<ST, DT>::value) {
                    bool all_ok = true;
#pragma unroll
                    for (int k = 0; k < V / 2; k++) {
                        bool ok;
                        reinterpret_cast<uint32_t *>(out)[k] = cast_pair16<ST, DT>(reinterpret_cast<const uint32_t *>(in[u])[k], ok);
                        all_ok &= ok;
                    }
                    if (!all_ok) {
#pragma unroll
                        for (int k = 0; k < V; k++) out[k] = narrow_exact<DT>(widen_exact<ST>(in[u][k]));
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < V; k++) out[k] = narrow_exact<DT>(widen_exact<ST>(in[u][k]));
                }
                *reinterpret_cast<DV *>(dp + u * CAST_THREADS * V) = *reinterpret_cast<DV *>(out);
            }
        }
    }
    // tail (and the whole range when a pointer is not 16-byte aligned): element-wise grid stride
    const size_t start = vec ? full_tiles * CAST_TILE : 0;
    for (size_t i = start + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = narrow_exact<DT>(widen_exact<ST>(src[i]));
}

template <class ST, class DT>
static cudaError_t launch_cast(const void *src, void *dst, size_t n, cudaStream_t stream, DevInfo dev) {
    auto kern = cast_kernel<ST, DT>;
    static std::atomic<int> occ_cache[MAX_DEVICES];
    int ctas_per_sm = 0;
    cudaError_t e = cached_occupancy(kern, CAST_THREADS, 0, dev.device, occ_cache, &ctas_per_sm);
    if (e != cudaSuccess) return e;
    // one tile per CTA (a persistent grid exposes the load latency of every tile; letting the hardware
    // overlap many short CTAs is what the 1:1 copy of rearrange.cu does at 6.5-6.9 TB/s); the grid is
    // capped only where it would overflow 2^31 - 1 blocks
    size_t want = (n + CAST_TILE - 1) / CAST_TILE;
    size_t grid = want ? want : 1;
    if (grid > 0x7FFFFFFFull) grid = (size_t)dev.sm_count * ctas_per_sm;
    return launch_pdl(kern, (unsigned)grid, CAST_THREADS, 0, stream, static_cast<const typename ST::raw *>(src), static_cast<typename DT::raw *>(dst), n);
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
