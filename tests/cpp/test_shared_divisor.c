/* The K-quant quantize kernels divide the 16 / 32 numerators of a sub-block by one shared divisor with
 *   r = RN(1/d);  q0 = RN(n r);  e = n - d q0 (one FMA, exact);  q = RN(q0 + e r)
 * (gguf_b200/csrc/quant_k.cu: div_shared) instead of an IEEE division per element.  This checks, on the host, that the
 * sequence returns the bits of `n / d` over the ranges the kernel admits (2^-40 <= |d| <= 2^40, |n| <= 2^60): random
 * pairs, and every numerator significand against divisors whose significands are the classic hard cases (all ones,
 * 1 + ulp, midpoints, repeating patterns).  Build: gcc -O2 -mfma -ffp-contract=off. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static uint64_t s = 88172645463325252ULL;
static uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static float f_from(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static int same_or_tiny(float n, float d, long *checked) {
    const float r = 1.0f / d, q0 = n * r, e = fmaf(-d, q0, n), q = fmaf(e, r, q0), t = n / d;
    ++*checked;
    if (!memcmp(&q, &t, 4)) return 1;
    return fabsf(t) < 0x1p-63f;  /* residual underflow: nearest_int() is 0 either way */
}

int main(void) {
    long checked = 0, bad = 0;
    for (long it = 0; it < 40000000L; it++) {
        const int ed = 87 + (int)(rnd() % 81);                         /* 2^-40 .. 2^40 */
        int en = (it & 1) ? ed - 6 + (int)(rnd() % 14) : 27 + (int)(rnd() % 160);
        if (en < 1) en = 1;
        if (en > 186) en = 186;                                       /* |n| < 2^60 */
        const float d = f_from(((uint32_t)rnd() & 0x807FFFFFu) | ((uint32_t)ed << 23));
        const float n = f_from(((uint32_t)rnd() & 0x807FFFFFu) | ((uint32_t)en << 23));
        if (!same_or_tiny(n, d, &checked) && bad++ < 10) printf("mismatch n=%a d=%a\n", n, d);
    }
    static const uint32_t hard[] = {0x7FFFFF, 0x7FFFFE, 0x000000, 0x000001, 0x400000, 0x3FFFFF, 0x400001, 0x2AAAAA, 0x555555, 0x6DB6DB};
    for (unsigned k = 0; k < sizeof hard / sizeof *hard; k++) {
        const float d = f_from((117u << 23) | hard[k]);
        for (uint32_t m = 0; m < (1u << 23); m += 3)
            for (int en = 118; en <= 121; en += 3)
                if (!same_or_tiny(f_from(((uint32_t)en << 23) | m), d, &checked) && bad++ < 10) printf("mismatch m=%x d=%a\n", m, d);
    }
    /* strict variant (div_shared_strict: the quotient itself is used): 2^-60 <= |n| <= 2^60, no tolerance at all; the
     * numerators the searches really form (14 .. 33 in steps of 0.1, both signs) against random divisors, and random pairs */
    long strict = 0;
    for (long it = 0; it < 30000000L; it++) {
        const int ed = 87 + (int)(rnd() % 81);
        const float d = f_from(((uint32_t)rnd() & 0x807FFFFFu) | ((uint32_t)ed << 23));
        float n;
        if (it % 3 == 0) n = (float)(140 + (int)(rnd() % 191)) * 0.1f * ((rnd() & 1) ? 1.f : -1.f);
        else n = f_from(((uint32_t)rnd() & 0x807FFFFFu) | ((uint32_t)(67 + (int)(rnd() % 121)) << 23));   /* 2^-60 .. 2^60 */
        const float r = 1.0f / d, q0 = n * r, e = fmaf(-d, q0, n), q = fmaf(e, r, q0), t = n / d;
        ++strict;
        if (memcmp(&q, &t, 4) && bad++ < 10) printf("strict mismatch n=%a d=%a\n", n, d);
    }
    checked += strict;
    printf("%ld quotients checked (%ld strict), %ld mismatches\n", checked, strict, bad);
    return bad != 0;
}
