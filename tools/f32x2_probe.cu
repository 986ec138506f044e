// f32x2_probe.cu — what do the packed FP32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2, `*.f32x2`) buy?
// Times long dependent chains over 16 independent accumulators per thread: scalar FMUL+FADD, packed
// FMUL2 then FADD2 on the SAME value (ptxas 12.9 contracts that pair into one FFMA2 even with
// --fmad=false), and packed FMUL2 / FADD2 on separate chains (cannot be contracted).  Prints G FP32
// operations per second.  Tuning harness, not product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/f32x2_probe tools/f32x2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void probe(float *out, int iters, float a, float b) {
    float2 acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = make_float2(threadIdx.x + k, threadIdx.x - k);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (MODE == 0) {  // scalar: 4 instructions per pair
                acc[k].x = __fadd_rn(__fmul_rn(acc[k].x, a), b);
                acc[k].y = __fadd_rn(__fmul_rn(acc[k].y, a), b);
            } else if (MODE == 1) {  // packed, contractable: ptxas emits one FFMA2 per pair
                acc[k] = __fadd2_rn(__fmul2_rn(acc[k], a2), b2);
            } else {          // packed, not contractable: even accumulators multiply, odd ones add (2 instructions per 2 pairs... same op count)
                if (k & 1) acc[k] = __fadd2_rn(acc[k], b2); else acc[k] = __fmul2_rn(acc[k], a2);
                if (k & 1) acc[k] = __fadd2_rn(acc[k], a2); else acc[k] = __fmul2_rn(acc[k], b2);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += acc[k].x + acc[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float *out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    const int iters = 20000;
    for (int mode = 0; mode < 3; mode++)
        for (int rep = 0; rep < 2; rep++) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            if (mode == 0) probe<0><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f);
            else if (mode == 1) probe<1><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f);
            else probe<2><<<148 * 8, 256>>>(out, iters, 0.999f, 1.001f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double ops = 148.0 * 8 * 256 * iters * 32.0;  // FP32 operations (mul + add on 16 values)
            printf("%s: %.2f ms  %.1f G FP32 ops/s\n", mode == 0 ? "scalar FMUL+FADD          " : mode == 1 ? "packed, contracted (FFMA2)" : "packed FMUL2 / FADD2      ", ms, ops / ms / 1e6);
        }
    return 0;
}
