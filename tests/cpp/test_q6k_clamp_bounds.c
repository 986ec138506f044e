/* make_qx_quants16 (gguf_b200/csrc/quant_k_kernel.cuh, CS 1) applies the clamp to [-32, 31] only to the candidates that
 * can reach a bound.  The claim, for every sub-block value |x| <= |mx| and candidate
 *     iscale = RN(-(32 + RN(0.1f * is)) / mx),   v = RN(iscale * x),   r = rint(v):
 *   is in [-9, -6]                 ->  -31 <= r <= 31   (no bound can bind)
 *   is in [-5, 4] and the first evaluation (iscale = RN(-32 / mx))   ->  r >= -32   (only the upper bound 31 can)
 * This evaluates exactly those float expressions on the host for the ratios that maximise |v| — x = +-mx and the
 * neighbouring floats — over random and adversarial mx (every exponent, significands all ones / zero / midpoints), plus
 * random ratios.  Build: gcc -O2 -ffp-contract=off. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static uint64_t s = 0x9E3779B97F4A7C15ULL;
static uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
static float f_from(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t u_from(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static long bad = 0, checked = 0;
static void check(float mx, float x) {
    if (!(fabsf(x) <= fabsf(mx))) return;
    const float fn = 32.f;
    for (int is = -10; is <= 4; ++is) {   /* is == -10 stands for the first evaluation: iscale = -32 / mx */
        volatile float num = is == -10 ? -fn : -(fn + 0.1f * (float)is);
        volatile float iscale = num / mx;
        volatile float v = iscale * x;
        const float r = rintf(v);
        ++checked;
        int ok;
        if (is >= -9 && is <= -6) ok = r >= -31.f && r <= 31.f;
        else ok = r >= -32.f;             /* is == 0 is skipped by the search; it satisfies the bound anyway */
        if (!ok && bad++ < 10) printf("bound violated: mx=%a x=%a is=%d v=%a r=%g\n", mx, x, is, v, r);
    }
}

int main(void) {
    static const uint32_t sig[] = {0x000000, 0x000001, 0x7FFFFF, 0x7FFFFE, 0x400000, 0x3FFFFF, 0x400001, 0x2AAAAA, 0x555555, 0x6DB6DB, 0x199999, 0x19999A, 0x4CCCCC, 0x4CCCCD};
    /* GROUP_MAX_EPS = 1e-15 (~2^-50) is the smallest amax the search sees; the largest is FLT_MAX */
    for (int e = 127 - 50; e <= 254; e++) {
        for (unsigned k = 0; k < sizeof sig / sizeof *sig + 2000; k++) {
            const uint32_t m = k < sizeof sig / sizeof *sig ? sig[k] : (uint32_t)rnd() & 0x7FFFFF;
            for (int sgn = 0; sgn < 2; sgn++) {
                const float mx = f_from(((uint32_t)sgn << 31) | ((uint32_t)e << 23) | m);
                if (fabsf(mx) < 1e-15f) continue;
                /* the extreme ratios: x = +-mx and up to 3 floats below in magnitude */
                for (int d = 0; d < 4; d++) {
                    const uint32_t a = (u_from(mx) & 0x7FFFFFFFu) - (uint32_t)d;
                    check(mx, f_from(a));
                    check(mx, f_from(a | 0x80000000u));
                }
                /* and random smaller values */
                for (int t = 0; t < 4; t++) {
                    const float ratio = (float)(rnd() % 16777216) / 16777216.0f;
                    check(mx, mx * ratio);
                    check(mx, -mx * ratio);
                }
            }
        }
    }
    printf("%ld candidate evaluations checked, %ld bound violations\n", checked, bad);
    return bad ? 1 : 0;
}
