// kq_sweep.cu — times configurations of the shipped K-quant quantize kernel template (gguf_b200/csrc/quant_k_kernel.cuh)
// on a 4096x14336 f16 (or f32) Gaussian tensor and checks that every configuration writes the same bytes as the first.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false --ftz=false --prec-div=true --prec-sqrt=true \
//        -I gguf_b200/csrc tools/kq_sweep.cu -o tools/kq_sweep
// The library's Makefile refuses objects with a contracted packed multiply-add (an FFMA2 whose multiplier is not the uniform
// register holding 1.0); this binary is not guarded, so a variant that invites the contraction shows up as "BYTES DIFFER"
// here (it did: Q6K unclamped candidates rounded with the magic-number add).  The same check by hand:
//   cuobjdump -sass tools/kq_sweep | grep FFMA2 | grep -v "UR[0-9]*\.F32"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <map>
#include <vector>

#include "quant_k_kernel.cuh"

using namespace ggq;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static int g_sms = 148;
static std::map<uint32_t, std::vector<uint8_t>> g_refs;   // bytes of the first configuration, per block type
static const char *g_only = nullptr;   // argv[2]: run only configurations whose name contains this (e.g. T_Q6K)
static bool g_first_only = false;      // argv[3] == "first": only the reference configuration of each type (for ncu)

static bool g_dyn = false;             // hand the warp passes out by the ticket counter
static int g_passes = 0;               // > 0: not persistent — a grid of ngroups / (WARPS * g_passes) CTAs, the hardware hands them out
template <uint32_t T, class FT, class CFG>
void run(const char *name_, const void *d_x, uint8_t *d_out, size_t nblocks, bool is_ref) {
    char name[96];
    if (g_passes) snprintf(name, sizeof name, "%s grid/%d", name_, g_passes);
    else snprintf(name, sizeof name, "%s%s", name_, g_dyn ? " tickets" : "");
    if (g_only && !strstr(name, g_only)) return;
    if (g_first_only && !is_ref) return;
    constexpr int SBW = 32 / (256 / KQuant<T>::SUB);
    constexpr int BYTES = BlockTraits<T>::BYTES;
    auto kern = quant_k_kernel<T, FT, CFG>;
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, CFG::THREADS, 0));
    const size_t ngroups = (nblocks + SBW - 1) / SBW;
    size_t grid = (size_t)g_sms * occ;
    if (grid > (ngroups + CFG::WARPS - 1) / CFG::WARPS) grid = (ngroups + CFG::WARPS - 1) / CFG::WARPS;
    if (g_passes) grid = (ngroups + (size_t)CFG::WARPS * g_passes - 1) / ((size_t)CFG::WARPS * g_passes);
    CK(cudaMemset(d_out, 0xEE, nblocks * BYTES));
    static unsigned long long *d_work = nullptr;
    if (!d_work) CK(cudaMalloc(&d_work, 8));
    CK(cudaMemsetAsync(d_work, 0, 8));
    kern<<<(unsigned)grid, CFG::THREADS>>>(static_cast<const typename FT::raw *>(d_x), d_out, nblocks, 1.0f, g_dyn ? d_work : nullptr);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int reps = 3;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) {
        if (g_dyn) CK(cudaMemsetAsync(d_work, 0, 8));   // inside the timed region: the shipped launcher pays it too
        kern<<<(unsigned)grid, CFG::THREADS>>>(static_cast<const typename FT::raw *>(d_x), d_out, nblocks, 1.0f, g_dyn ? d_work : nullptr);
    }
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<uint8_t> out(nblocks * BYTES);
    CK(cudaMemcpy(out.data(), d_out, out.size(), cudaMemcpyDeviceToHost));
    const char *same = "ref";
    std::vector<uint8_t> &g_ref = g_refs[T];
    if (is_ref) g_ref = out;
    else same = (out.size() == g_ref.size() && memcmp(out.data(), g_ref.data(), out.size()) == 0) ? "same bytes" : "BYTES DIFFER";
    printf("%-42s regs %3d  spill %3zu  smem %6zu  warps/SM %2d  %9.1f us   %s\n", name, fa.numRegs, (size_t)fa.localSizeBytes, (size_t)fa.sharedSizeBytes, occ * CFG::WARPS,
           ms * 1000.f / reps, same);
    fflush(stdout);
}

#define RUN(T, FT, W, REGS, LF, WM, ST, ref) run<T, FT, KqCfg<W, REGS, LF, WM, ST>>(#T " " #FT " <" #W "," #REGS "," #LF "," #WM "," #ST ">", d_x, d_out, nblocks, ref)
#define RUNA(T, FT, W, REGS, LF, WM, ST, AF) run<T, FT, KqCfg<W, REGS, LF, WM, ST, AF>>(#T " " #FT " <" #W "," #REGS "," #LF "," #WM "," #ST "," #AF ">", d_x, d_out, nblocks, false)
#define RUNS(T, FT, W, REGS, LF, WM, ST, AF, RM, SD) run<T, FT, KqCfg<W, REGS, LF, WM, ST, AF, RM, SD>>(#T " " #FT " <" #W "," #REGS "," #LF "," #WM "," #ST "," #AF "," #RM "," #SD ">", d_x, d_out, nblocks, false)
#define RUNC(T, FT, W, REGS, LF, WM, ST, AF, RM, SD, CL) run<T, FT, KqCfg<W, REGS, LF, WM, ST, AF, RM, SD, CL>>(#T " " #FT " <" #W "," #REGS "," #LF "," #WM "," #ST "," #AF "," #RM "," #SD "," #CL ">", d_x, d_out, nblocks, false)
#define RUNP(T, FT, W, REGS, LF, WM, ST, AF, RM, SD, CL, SP) run<T, FT, KqCfg<W, REGS, LF, WM, ST, AF, RM, SD, CL, SP>>(#T " " #FT " <" #W "," #REGS "," #LF "," #WM "," #ST "," #AF "," #RM "," #SD "," #CL "," #SP ">", d_x, d_out, nblocks, false)
#define RUNQ(T, FT, W, REGS, LF, WM, ST, AF, RM, SD, CL, SP, CS) run<T, FT, KqCfg<W, REGS, LF, WM, ST, AF, RM, SD, CL, SP, CS>>(#T " " #FT " <" #W "," #REGS "," #LF "," #WM "," #ST "," #AF "," #RM "," #SD "," #CL "," #SP "," #CS ">", d_x, d_out, nblocks, false)
#define RUNR(T, FT, W, REGS, LF, WM, ST, AF, RM) run<T, FT, KqCfg<W, REGS, LF, WM, ST, AF, RM>>(#T " " #FT " <" #W "," #REGS "," #LF "," #WM "," #ST "," #AF "," #RM ">", d_x, d_out, nblocks, false)

int main(int argc, char **argv) {
    const bool f32 = argc > 1 && !strcmp(argv[1], "f32");
    if (argc > 2 && strcmp(argv[2], "all")) g_only = argv[2];
    g_first_only = argc > 3 && !strcmp(argv[3], "first");
    const size_t n = (size_t)4096 * 14336, nblocks = n / 256;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    std::vector<float> x(n);
    std::mt19937 rng(2);
    std::normal_distribution<float> nd(0.f, 0.02f);
    for (auto &v : x) v = nd(rng);
    void *d_x;
    uint8_t *d_out;
    CK(cudaMalloc(&d_out, nblocks * 210));
    if (f32) {
        CK(cudaMalloc(&d_x, n * 4));
        CK(cudaMemcpy(d_x, x.data(), n * 4, cudaMemcpyHostToDevice));
    } else {
        std::vector<__half> h(n);
        for (size_t i = 0; i < n; i++) h[i] = __float2half_rn(x[i]);
        CK(cudaMalloc(&d_x, n * 2));
        CK(cudaMemcpy(d_x, h.data(), n * 2, cudaMemcpyHostToDevice));
    }
    printf("%s input, %zu elements, %d SMs; config <WARPS, REGS, LF, WM, STAGES>\n", f32 ? "f32" : "f16", n, g_sms);
    // rounds: 0 persistent grid, fixed stride (what shipped until round 2's last step) | 1 persistent, passes by ticket
    // (ships for Q3K) | 2.. not persistent: 1 / 2 / 3 / 6 passes per warp (1 ships for Q2K / Q4K / Q5K / Q6K)
    const int nrounds = getenv("KQ_ROUNDS") ? atoi(getenv("KQ_ROUNDS")) : 6;
    for (int dyn = 0; dyn < nrounds; dyn++) {
        g_dyn = dyn == 1;
        g_passes = dyn >= 2 ? (dyn == 5 ? 6 : dyn - 1) : 0;
        if (!f32) {
            RUN(T_Q4K, F16, 4, 128, 0, 2, 2, dyn == 0);
            RUNC(T_Q4K, F16, 4, 128, 0, 2, 2, 1, 0, 0, 1);   // shipped: packed affine, clamp as one VIMNMX.RELU
            RUNC(T_Q4K, F16, 4, 168, 0, 2, 2, 1, 0, 0, 1);   // 140 registers, three warps per scheduler, no spill
            RUNC(T_Q4K, F16, 2, 128, 0, 2, 2, 1, 0, 0, 1);   // two-warp CTAs
            RUNC(T_Q4K, F16, 1, 128, 0, 2, 2, 1, 0, 0, 1);   // one-warp CTAs (shipped)
            RUNC(T_Q4K, F16, 1, 128, 0, 2, 2, 1, 0, 1, 1);   // + shared-divisor quotients (the one-warp build has no spill to double)
            RUNC(T_Q4K, F16, 1, 136, 0, 2, 2, 1, 0, 0, 1);   // 136 registers: 15 one-warp CTAs per SM
            RUNC(T_Q4K, F16, 1, 120, 0, 2, 2, 1, 0, 0, 1);   // 120 registers: 17
            RUNC(T_Q4K, F16, 4, 128, 0, 2, 2, 1, 2, 0, 1);   // alternate pairs round on the FP32 pipe
            RUNC(T_Q4K, F16, 4, 128, 0, 2, 1, 1, 0, 0, 1);   // one input stage (a one-shot CTA has no next pass to prefetch)
            RUN(T_Q5K, F16, 4, 128, 0, 2, 2, dyn == 0);
            RUNC(T_Q5K, F16, 4, 128, 0, 2, 2, 1, 0, 0, 1);
            RUNC(T_Q5K, F16, 1, 128, 0, 2, 2, 1, 0, 0, 1);
            RUNC(T_Q5K, F16, 1, 128, 0, 2, 2, 1, 0, 1, 1);
            RUN(T_Q2K, F16, 4, 72, 0, 2, 2, dyn == 0);
            RUNC(T_Q2K, F16, 1, 64, 0, 2, 2, 1, 0, 0, 1);
            RUNC(T_Q2K, F16, 4, 64, 0, 2, 2, 1, 0, 0, 1);
            RUNC(T_Q2K, F16, 2, 64, 0, 2, 2, 1, 0, 0, 1);
            RUN(T_Q6K, F16, 4, 96, 0, 2, 2, dyn == 0);
            RUNR(T_Q6K, F16, 4, 96, 0, 2, 2, 0, 2);
            RUNQ(T_Q6K, F16, 4, 96, 0, 2, 2, 0, 2, 0, 0, 0, 1);   // clamp only where a candidate can reach the bound
            RUNQ(T_Q6K, F16, 4, 96, 0, 2, 2, 0, 0, 0, 0, 0, 1);
            RUNQ(T_Q6K, F16, 1, 96, 0, 2, 2, 0, 2, 0, 0, 0, 1);
            RUNC(T_Q2K, F16, 1, 64, 0, 2, 2, 1, 2, 0, 1);          // Q2K / Q5K with alternate pairs rounded on the FP32 pipe
            RUNC(T_Q5K, F16, 1, 128, 0, 2, 2, 1, 2, 0, 1);
            RUNR(T_Q6K, F16, 4, 96, 0, 2, 1, 0, 2);
            RUN(T_Q3K, F16, 4, 96, 0, 2, 2, dyn == 0);
            RUN(T_Q3K, F16, 2, 96, 0, 2, 2, false);
        } else {
            RUN(T_Q4K, F32, 4, 128, 0, 2, 2, dyn == 0);
            RUNC(T_Q4K, F32, 4, 128, 0, 2, 2, 1, 0, 0, 1);
            RUN(T_Q6K, F32, 4, 96, 0, 2, 2, dyn == 0);
            RUNR(T_Q6K, F32, 4, 96, 0, 2, 2, 0, 2);
            RUN(T_Q3K, F32, 4, 96, 0, 2, 2, dyn == 0);
        }
    }
    return 0;
}
