// dq_sweep.cu — tuning harness (not product): times configurations of the SHIPPED dequantize kernel
// template (gguf_b200/csrc/dequant_kernel.cuh) plus plain fill / copy kernels that calibrate what a
// write-heavy stream can reach on this GPU.  Buffers rotate over NSETS sets so the footprint
// exceeds the 126 MB L2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I gguf_b200/csrc -o tools/dq_sweep tools/dq_sweep.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dequant_kernel.cuh"
#include "ggq_kernels.h"

using namespace ggq;

#ifndef SWEEP_FT
#define SWEEP_FT F16   // -DSWEEP_FT=F32 sweeps the f32-output instantiations
#endif

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static int g_sms = 148;
static const int NSETS = 6;

__global__ void fill_kernel(uint4 *dst, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = make_uint4(1, 2, 3, 4);
}
__global__ void copy_kernel(uint4 *dst, const uint4 *src, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = __ldg(src + i);
}
// reads 1 uint4 per RATIO uint4 written (the dequant mix: Q4_0->f16 reads 0.28 B per B written)
template <int RATIO> __global__ void expand_kernel(uint4 *dst, const uint4 *src, size_t nsrc16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nsrc16; i += stride) {
        const uint4 v = __ldg(src + i);
        const size_t w = i / 32, l = i % 32;
#pragma unroll
        for (int r = 0; r < RATIO; r++) dst[(w * RATIO + r) * 32 + l] = v;
    }
}

// one-tile-per-CTA forms of the calibration kernels (grid = n16 / (256 * U)): what the hardware block scheduler
// reaches when every CTA is short-lived
template <int U> __global__ void fill1_kernel(uint4 *dst) {
    uint4 *p = dst + (size_t)blockIdx.x * 256 * U + threadIdx.x;
#pragma unroll
    for (int u = 0; u < U; u++) p[u * 256] = make_uint4(1, 2, 3, 4);
}
template <int U> __global__ void copy1_kernel(uint4 *dst, const uint4 *src) {
    const size_t o = (size_t)blockIdx.x * 256 * U + threadIdx.x;
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; u++) v[u] = __ldg(src + o + u * 256);
#pragma unroll
    for (int u = 0; u < U; u++) dst[o + u * 256] = v[u];
}
// 1 read : 4 write, one CTA = 256 source vectors -> 1024 destination vectors
__global__ void expand1_kernel(uint4 *dst, const uint4 *src) {
    const uint4 v = __ldg(src + (size_t)blockIdx.x * 256 + threadIdx.x);
    uint4 *p = dst + (size_t)blockIdx.x * 1024 + threadIdx.x;
#pragma unroll
    for (int r = 0; r < 4; r++) p[r * 256] = v;
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    void start() { cudaEventRecord(a); }
    float stop() { cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms; }
};

template <uint32_t T, class FT, int TILE, int STAGES, int THREADS, int MINB, int MODE, int SP>
void run_variant(const char *tname, size_t n_elems, uint8_t *const *in, void *const *out, int iters) {
    using TR = BlockTraits<T>;
    constexpr int TILE_BLOCKS = TILE / TR::ELEMS;
    constexpr int SMEM = dequant_smem_bytes<T, TILE, STAGES, MODE>();
    auto kern = dequant_kernel<T, FT, TILE, STAGES, THREADS, MINB, MODE, SP>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, SMEM));
    const size_t nblocks = n_elems / TR::ELEMS;
    const size_t ntiles = (nblocks + TILE_BLOCKS - 1) / TILE_BLOCKS;
    size_t grid = MODE == 0 ? (size_t)g_sms * occ : MODE == 3 ? (ntiles + STAGES - 1) / STAGES : ntiles;
    if (grid > ntiles) grid = ntiles;
    for (int i = 0; i < 3; i++) CK(launch_pdl(kern, (unsigned)grid, THREADS, SMEM, (cudaStream_t)0, (const uint8_t *)in[i % NSETS], (typename FT::raw *)out[i % NSETS], nblocks));
    CK(cudaDeviceSynchronize());
    Timer t;
    t.start();
    for (int i = 0; i < iters; i++) CK(launch_pdl(kern, (unsigned)grid, THREADS, SMEM, (cudaStream_t)0, (const uint8_t *)in[i % NSETS], (typename FT::raw *)out[i % NSETS], nblocks));
    const float ms = t.stop();
    CK(cudaGetLastError());
    const double bytes = (double)nblocks * TR::BYTES + (double)n_elems * FT::SIZE;
    printf("%-5s n=%-9zu tile=%-6d st=%d thr=%d minb=%d mode=%d sp=%d occ=%d grid=%-6zu : %8.2f us  %7.1f GB/s\n", tname, n_elems, TILE, STAGES,
           THREADS, MINB, MODE, SP, occ, grid, ms * 1e3 / iters, bytes * iters / (ms * 1e-3) / 1e9);
    fflush(stdout);
}

int main(int argc, char **argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t FFN = 4096ull * 14336, ATTN = 4096ull * 4096;
    const size_t in_bytes = FFN * 290 / 256 + 4096, out_bytes = FFN * SWEEP_FT::SIZE + 4096;
    uint8_t *in[NSETS];
    void *out[NSETS];
    for (int i = 0; i < NSETS; i++) {
        CK(cudaMalloc(&in[i], in_bytes));
        CK(cudaMalloc(&out[i], out_bytes));
        CK(cudaMemset(in[i], 0x11 * (i + 1), in_bytes));  // finite f16 scales (0x1111..), arbitrary codes
    }
    const int iters = argc > 1 ? atoi(argv[1]) : 60;
    std::vector<size_t> sizes;  // element counts (<= FFN) from the command line, default FFN and ATTN
    for (int a = 2; a < argc; a++) sizes.push_back((size_t)atoll(argv[a]));
    if (sizes.empty()) sizes = {FFN, ATTN};
    Timer t;

    // ---- calibration: what can pure streams reach ----
    for (int rep = 0; rep < 2; rep++) {
        const size_t n16 = FFN * 2 / 16;
        const int grid = g_sms * 8;
        t.start();
        for (int i = 0; i < iters; i++) fill_kernel<<<grid, 256>>>((uint4 *)out[i % NSETS], n16);
        float ms = t.stop();
        printf("fill   (write only, %zu MB)           : %8.2f us  %7.1f GB/s\n", n16 * 16 >> 20, ms * 1e3 / iters, (double)n16 * 16 * iters / (ms * 1e-3) / 1e9);
        t.start();
        for (int i = 0; i < iters; i++) copy_kernel<<<grid, 256>>>((uint4 *)out[i % NSETS], (const uint4 *)out[(i + 3) % NSETS], n16);
        ms = t.stop();
        printf("copy   (1 read : 1 write)             : %8.2f us  %7.1f GB/s\n", ms * 1e3 / iters, (double)n16 * 32 * iters / (ms * 1e-3) / 1e9);
        const size_t ns = FFN * 2 / 16 / 4;
        t.start();
        for (int i = 0; i < iters; i++) expand_kernel<4><<<grid, 256>>>((uint4 *)out[i % NSETS], (const uint4 *)in[i % NSETS], ns);
        ms = t.stop();
        printf("expand (1 read : 4 write)             : %8.2f us  %7.1f GB/s\n", ms * 1e3 / iters, (double)ns * 16 * 5 * iters / (ms * 1e-3) / 1e9);
        t.start();
        for (int i = 0; i < iters; i++) fill1_kernel<4><<<(unsigned)(n16 / 1024), 256>>>((uint4 *)out[i % NSETS]);
        ms = t.stop();
        printf("fill1  (one 16 KB tile per CTA)       : %8.2f us  %7.1f GB/s\n", ms * 1e3 / iters, (double)n16 * 16 * iters / (ms * 1e-3) / 1e9);
        t.start();
        for (int i = 0; i < iters; i++) copy1_kernel<4><<<(unsigned)(n16 / 1024), 256>>>((uint4 *)out[i % NSETS], (const uint4 *)out[(i + 3) % NSETS]);
        ms = t.stop();
        printf("copy1  (one 16 KB tile per CTA)       : %8.2f us  %7.1f GB/s\n", ms * 1e3 / iters, (double)n16 * 32 * iters / (ms * 1e-3) / 1e9);
        t.start();
        for (int i = 0; i < iters; i++) expand1_kernel<<<(unsigned)(ns / 256), 256>>>((uint4 *)out[i % NSETS], (const uint4 *)in[i % NSETS]);
        ms = t.stop();
        printf("expand1 (1:4, 4 KB in per CTA)        : %8.2f us  %7.1f GB/s\n", ms * 1e3 / iters, (double)ns * 16 * 5 * iters / (ms * 1e-3) / 1e9);
    }

#define V(T, NAME, N, TILE, ST, THR, MINB, MODE, SP) run_variant<T, SWEEP_FT, TILE, ST, THR, MINB, MODE, SP>(NAME, N, in, out, iters)
#if !defined(SWEEP_SET) || SWEEP_SET == 1
#define SWEEP(T, NAME, N)                      \
    V(T, NAME, N, 8192, 3, 256, 3, 0, 0);      \
    V(T, NAME, N, 8192, 2, 256, 3, 0, 0);      \
    V(T, NAME, N, 16384, 3, 256, 3, 0, 0);     \
    V(T, NAME, N, 16384, 2, 256, 3, 0, 0);     \
    V(T, NAME, N, 16384, 2, 512, 1, 0, 0);     \
    V(T, NAME, N, 16384, 3, 512, 1, 0, 0);     \
    V(T, NAME, N, 32768, 2, 512, 1, 0, 0);     \
    V(T, NAME, N, 4096, 1, 256, 3, 1, 0);      \
    V(T, NAME, N, 8192, 1, 256, 3, 1, 0);      \
    V(T, NAME, N, 16384, 1, 256, 3, 1, 0);     \
    V(T, NAME, N, 8192, 1, 512, 1, 1, 0);      \
    V(T, NAME, N, 16384, 1, 512, 1, 1, 0);     \
    V(T, NAME, N, 4096, 1, 128, 6, 1, 0);      \
    V(T, NAME, N, 4096, 1, 256, 3, 2, 0);      \
    V(T, NAME, N, 8192, 1, 256, 3, 2, 0);      \
    V(T, NAME, N, 16384, 1, 256, 3, 2, 0);     \
    V(T, NAME, N, 8192, 1, 512, 1, 2, 0);      \
    V(T, NAME, N, 16384, 1, 512, 1, 2, 0);     \
    V(T, NAME, N, 4096, 1, 128, 6, 2, 0);
#else
// one tile per CTA (MODE 1) needs many resident CTAs: force the register cap down with MINB
#if SWEEP_SET == 2 || SWEEP_SET == 4
#define SWEEP_A(T, NAME, N)                    \
    V(T, NAME, N, 16384, 1, 256, 4, 1, 0);     \
    V(T, NAME, N, 16384, 1, 256, 5, 1, 0);     \
    V(T, NAME, N, 16384, 1, 256, 6, 1, 0);     \
    V(T, NAME, N, 8192, 1, 128, 8, 1, 0);      \
    V(T, NAME, N, 8192, 1, 128, 10, 1, 0);     \
    V(T, NAME, N, 8192, 1, 128, 12, 1, 0);     \
    V(T, NAME, N, 4096, 1, 128, 10, 1, 0);     \
    V(T, NAME, N, 16384, 1, 128, 8, 1, 0);     \
    V(T, NAME, N, 16384, 1, 128, 10, 1, 0);    \
    V(T, NAME, N, 32768, 1, 256, 4, 1, 0);     \
    V(T, NAME, N, 32768, 1, 256, 6, 1, 0);
#else
#define SWEEP_A(T, NAME, N)
#endif
#if SWEEP_SET == 7
// ring (MODE 0) with small CTAs and forced occupancy, for small tensors
#define SWEEP_B(T, NAME, N)                    \
    V(T, NAME, N, 8192, 3, 256, 3, 0, 0);      \
    V(T, NAME, N, 8192, 3, 128, 8, 0, 0);      \
    V(T, NAME, N, 8192, 2, 128, 8, 0, 0);      \
    V(T, NAME, N, 4096, 3, 128, 8, 0, 0);      \
    V(T, NAME, N, 4096, 4, 128, 10, 0, 0);     \
    V(T, NAME, N, 8192, 3, 128, 6, 0, 0);      \
    V(T, NAME, N, 4096, 3, 64, 16, 0, 0);      \
    V(T, NAME, N, 8192, 4, 256, 4, 0, 0);      \
    V(T, NAME, N, 4096, 3, 256, 4, 0, 0);
#elif SWEEP_SET == 6
// MODE 3: 2-4 consecutive tiles per short-lived CTA, all bulk copies issued up front
#define SWEEP_B(T, NAME, N)                    \
    V(T, NAME, N, 8192, 3, 256, 3, 0, 0);      \
    V(T, NAME, N, 16384, 1, 128, 8, 1, 0);     \
    V(T, NAME, N, 8192, 1, 128, 8, 1, 0);      \
    V(T, NAME, N, 4096, 2, 128, 8, 3, 0);      \
    V(T, NAME, N, 4096, 4, 128, 8, 3, 0);      \
    V(T, NAME, N, 8192, 2, 128, 8, 3, 0);      \
    V(T, NAME, N, 8192, 2, 128, 10, 3, 0);     \
    V(T, NAME, N, 2048, 4, 128, 10, 3, 0);     \
    V(T, NAME, N, 2048, 4, 64, 16, 3, 0);      \
    V(T, NAME, N, 4096, 2, 64, 16, 3, 0);      \
    V(T, NAME, N, 8192, 2, 256, 4, 3, 0);      \
    V(T, NAME, N, 4096, 4, 256, 4, 3, 0);
#elif SWEEP_SET == 5
#define SWEEP_B(T, NAME, N)                    \
    V(T, NAME, N, 8192, 3, 256, 3, 0, 0);      \
    V(T, NAME, N, 8192, 2, 256, 3, 0, 0);      \
    V(T, NAME, N, 16384, 3, 256, 3, 0, 0);     \
    V(T, NAME, N, 16384, 3, 512, 1, 0, 0);     \
    V(T, NAME, N, 32768, 2, 512, 1, 0, 0);     \
    V(T, NAME, N, 8192, 1, 128, 8, 1, 0);      \
    V(T, NAME, N, 16384, 1, 128, 8, 1, 0);     \
    V(T, NAME, N, 4096, 1, 128, 10, 1, 0);     \
    V(T, NAME, N, 8192, 1, 64, 16, 1, 0);
#elif SWEEP_SET == 3 || SWEEP_SET == 4
#define SWEEP_B(T, NAME, N)                    \
    V(T, NAME, N, 32768, 1, 128, 8, 1, 0);     \
    V(T, NAME, N, 32768, 1, 128, 10, 1, 0);    \
    V(T, NAME, N, 16384, 1, 128, 12, 1, 0);    \
    V(T, NAME, N, 16384, 1, 128, 9, 1, 0);     \
    V(T, NAME, N, 16384, 1, 64, 16, 1, 0);     \
    V(T, NAME, N, 16384, 1, 64, 20, 1, 0);     \
    V(T, NAME, N, 8192, 1, 64, 16, 1, 0);      \
    V(T, NAME, N, 8192, 1, 64, 20, 1, 0);      \
    V(T, NAME, N, 8192, 1, 64, 24, 1, 0);      \
    V(T, NAME, N, 8192, 1, 128, 8, 1, 0);      \
    V(T, NAME, N, 16384, 1, 128, 8, 1, 0);     \
    V(T, NAME, N, 16384, 1, 128, 10, 1, 0);
#else
#define SWEEP_B(T, NAME, N)
#endif
#define SWEEP(T, NAME, N) SWEEP_A(T, NAME, N) SWEEP_B(T, NAME, N)
#endif
    for (int rep = 0; rep < 2; rep++)
        for (size_t n : sizes) {
            SWEEP(T_Q8_0, "Q8_0", n)
            SWEEP(T_Q4_0, "Q4_0", n)
            SWEEP(T_Q6K, "Q6K", n)
            SWEEP(T_Q2K, "Q2K", n)
            SWEEP(T_Q3K, "Q3K", n)
            SWEEP(T_Q4K, "Q4K", n)
            SWEEP(T_Q5K, "Q5K", n)
#if defined(SWEEP_SET) && SWEEP_SET >= 2
            SWEEP(T_Q4_1, "Q4_1", n)
            SWEEP(T_Q5_0, "Q5_0", n)
            SWEEP(T_Q5_1, "Q5_1", n)
            SWEEP(T_Q8_1, "Q8_1", n)
            SWEEP(T_Q8K, "Q8K", n)
#endif
        }
    return 0;
}
