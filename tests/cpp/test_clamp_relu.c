/* The K-quant searches clamp to [0, nmax] with ONE instruction, PTX `min.relu.s32` on the float's bit pattern
 * (gguf_b200/csrc/quant_k_kernel.cuh, clamp0_relu): d = max(min((int)bits(v), (int)bits(hi)), 0).  This checks on the host,
 * for EVERY float v that is not a NaN and hi = 3, 15, 31 (the nmax of Q2K / Q4K / Q5K), that the result is the value
 * fminf(fmaxf(v, 0), hi) returns (+0 and -0 compare equal: a zero code contributes +-0 to every sum either way).
 * `test_clamp_relu all` walks all 2^32 patterns per bound (12 834 570 246 non-NaN checks, 0 differ: two minutes on one
 * core); without the argument every 13th pattern plus 4096 patterns either side of 0, -0, hi, +-inf and the NaN
 * boundaries, which the CPU suite runs.  Build: gcc -O2. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static float f_from(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t u_from(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static long bad = 0, checked = 0;
static void check(uint32_t u, float hi) {
    const int32_t hb = (int32_t)u_from(hi);
    const float v = f_from(u);
    if (v != v) return;
    int32_t r = (int32_t)u < hb ? (int32_t)u : hb;   /* min.s32 */
    if (r < 0) r = 0;                                  /* .relu */
    const float got = f_from((uint32_t)r), want = fminf(fmaxf(v, 0.f), hi);
    ++checked;
    if (!(got == want) && bad++ < 10) printf("v=%a hi=%g: got %a want %a\n", v, hi, got, want);
}

int main(int argc, char **argv) {
    const float his[3] = {3.f, 15.f, 31.f};
    const uint32_t step = argc > 1 && !strcmp(argv[1], "all") ? 1u : 13u;
    for (int h = 0; h < 3; h++) {
        uint32_t u = 0;
        for (;;) {
            check(u, his[h]);
            if (u > 0xFFFFFFFFu - step) break;
            u += step;
        }
        const uint32_t centres[] = {0u, 0x80000000u, u_from(his[h]), u_from(-his[h]), 0x7F800000u, 0xFF800000u, 0x00800000u, 0x80800000u};
        for (unsigned c = 0; c < sizeof centres / sizeof *centres; c++)
            for (int d = -4096; d <= 4096; d++) check(centres[c] + (uint32_t)d, his[h]);
    }
    printf("%ld floats checked, %ld differ\n", checked, bad);
    return bad ? 1 : 0;
}
