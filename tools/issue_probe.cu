// issue_probe.cu — does a packed FP32 instruction (FMUL2 / FADD2, `*.f32x2`) leave its second pipe cycle free for an
// instruction of ANOTHER pipe?  The K-quant search mixes packed FP32 with FMNMX / VIMNMX (ALU), FRND (XU) and scalar FADD
// chains; whether its roofline is "instructions issued" or "FP32 lane-operations + everything else" depends on the answer.
// Each mode runs a long unrolled loop of independent chains (no latency exposure: 8 chains per kind, 8 warps per
// scheduler) and prints warp-instructions per cycle per scheduler, and cycles per loop body.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/issue_probe tools/issue_probe.cu
// Tuning harness, not product.
#include <cstdio>
#include <cuda_runtime.h>

#define P_FMUL2(k) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(a2))
#define P_FADD(k) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[k]) : "f"(a))
#define P_FMUL(k) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[k]) : "f"(a))
#define P_ALU(k) asm volatile("min.relu.s32 %0, %0, %1;" : "+r"(n[k]) : "r"(m))
#define P_FMNMX(k) asm volatile("min.f32 %0, %0, %1;" : "+f"(g[k]) : "f"(a))
#define P_FRND(k) asm volatile("cvt.rni.f32.f32 %0, %0;" : "+f"(g[k]))
#define P_FADD2(k) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(a2))
#define P_LOP(k) asm volatile("xor.b32 %0, %0, %1;" : "+r"(n[k]) : "r"(m))

// MODE: 0 FMUL2 only | 1 FMUL2 + VIMNMX 1:1 | 2 FMUL2 + FMNMX 1:1 | 3 FMUL2 + scalar FADD 1:1 | 4 scalar FMUL + VIMNMX 1:1
//       5 scalar FMUL only | 6 FMUL2 + FRND 4:1 | 7 VIMNMX only | 8 FMUL2 + LOP3 1:1 | 9 FMUL2 + 2 VIMNMX | 10 FRND only
//       11 scalar FADD + FRND 4:1 | 12 FMUL2 x8 then FADD x8 (grouped) | 13 (FMUL2, FMUL2, FADD, FADD) x4
//       14 FADD2 + scalar FADD 1:1 | 15 scalar FMUL + scalar FADD 1:1 | 16 the search's candidate-pass mix per pair:
//       FADD2 FMUL2 VIMNMX VIMNMX FRND FRND FMUL2 FMUL2 FMUL2 + 6 FADD on three chains
template <int MODE> __global__ void __launch_bounds__(256) probe(float *out, int iters, float a, int m) {
    unsigned long long p[8];
    float f[8], g[8];
    int n[8];
    unsigned long long a2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float v = 1.0f + 0.001f * (threadIdx.x + k);
        asm("mov.b64 %0, {%1, %1};" : "=l"(p[k]) : "f"(v));
        f[k] = v; g[k] = v * 3.0f; n[k] = threadIdx.x * 77 + k;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (MODE == 0) { P_FMUL2(k); }
            if (MODE == 1) { P_FMUL2(k); P_ALU(k); }
            if (MODE == 2) { P_FMUL2(k); P_FMNMX(k); }
            if (MODE == 3) { P_FMUL2(k); P_FADD(k); }
            if (MODE == 4) { P_FMUL(k); P_ALU(k); }
            if (MODE == 5) { P_FMUL(k); }
            if (MODE == 6) { P_FMUL2(k); if ((k & 3) == 0) P_FRND(k); }
            if (MODE == 7) { P_ALU(k); }
            if (MODE == 8) { P_FMUL2(k); P_LOP(k); }
            if (MODE == 9) { P_FMUL2(k); P_ALU(k); P_LOP((k + 4) & 7); }
            if (MODE == 10) { P_FRND(k); }
            if (MODE == 11) { P_FADD(k); if ((k & 3) == 0) P_FRND(k); }
            if (MODE == 13) { if ((k & 1) == 0) { P_FMUL2(k); P_FMUL2(k + 1); P_FADD(k); P_FADD(k + 1); } }
            if (MODE == 14) { P_FADD2(k); P_FADD(k); }
            if (MODE == 15) { P_FMUL(k); P_FADD((k + 4) & 7); }
            if (MODE == 16) {
                P_FADD2(k); P_FMUL2((k + 3) & 7); P_ALU(k); P_ALU((k + 1) & 7); P_FRND(k); P_FRND((k + 5) & 7);
                P_FMUL2((k + 1) & 7); P_FMUL2((k + 2) & 7); P_FMUL2((k + 4) & 7);
                P_FADD(k); P_FADD((k + 1) & 7); P_FADD((k + 2) & 7); P_FADD((k + 3) & 7); P_FADD((k + 4) & 7); P_FADD((k + 5) & 7);
            }
        }
        if (MODE == 12) {
#pragma unroll
            for (int k = 0; k < 8; k++) P_FMUL2(k);
#pragma unroll
            for (int k = 0; k < 8; k++) P_FADD(k);
        }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[k]));
        s += lo + hi + f[k] + g[k] + (float)n[k];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static int g_ctas = 4;   // CTAs of 8 warps per SM: 4 -> 8 warps per scheduler, 2 -> 4
template <int MODE> void run(const char *name, int instr_per_body, float *out, int sms, double ghz) {
    const int iters = 4000;
    for (int rep = 0; rep < 2; rep++) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        probe<MODE><<<sms * g_ctas, 256>>>(out, iters, 0.9999f, 0x7fffffff);   // g_ctas CTAs x 8 warps per SM
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 0) continue;
        // per scheduler: 8 warps x iters x instr_per_body instructions in ms * ghz * 1e6 cycles
        const double wps = 2.0 * g_ctas;   // warps per scheduler
        const double cycles = ms * 1e-3 * ghz * 1e9, instr = wps * iters * instr_per_body;
        printf("%-34s %8.3f ms  %5.2f warp-instr / cycle / scheduler   %6.2f cycles per body of %d\n", name, ms, instr / cycles, cycles / (wps * iters), instr_per_body);
    }
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    float *out;
    cudaMalloc(&out, (size_t)prop.multiProcessorCount * 4 * 256 * sizeof(float));
    printf("%s, %d SMs, %.3f GHz (attribute; cycles assume the SM runs at it)\n", prop.name, prop.multiProcessorCount, ghz);
    const int s = prop.multiProcessorCount;
  for (int pass = 0; pass < 2; pass++) {
    g_ctas = pass == 0 ? 4 : 2;
    printf("---- %d warps per scheduler ----\n", 2 * g_ctas);
    run<5>("scalar FMUL x8", 8, out, s, ghz);
    run<0>("FMUL2 x8", 8, out, s, ghz);
    run<7>("VIMNMX.RELU x8", 8, out, s, ghz);
    run<10>("FRND x8", 8, out, s, ghz);
    run<4>("scalar FMUL + VIMNMX (8+8)", 16, out, s, ghz);
    run<1>("FMUL2 + VIMNMX (8+8)", 16, out, s, ghz);
    run<8>("FMUL2 + LOP3 (8+8)", 16, out, s, ghz);
    run<2>("FMUL2 + FMNMX (8+8)", 16, out, s, ghz);
    run<9>("FMUL2 + VIMNMX + LOP3 (8+8+8)", 24, out, s, ghz);
    run<3>("FMUL2 + scalar FADD (8+8)", 16, out, s, ghz);
    run<6>("FMUL2 + FRND (8+2)", 10, out, s, ghz);
    run<11>("scalar FADD + FRND (8+2)", 10, out, s, ghz);
    run<15>("scalar FMUL + scalar FADD (8+8)", 16, out, s, ghz);
    run<14>("FADD2 + scalar FADD (8+8)", 16, out, s, ghz);
    run<12>("FMUL2 x8 then FADD x8", 16, out, s, ghz);
    run<13>("(FMUL2 FMUL2 FADD FADD) x4", 16, out, s, ghz);
    run<16>("candidate-pass mix x8 (15 each)", 120, out, s, ghz);
  }
    return 0;
}
