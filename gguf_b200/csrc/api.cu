// api.cu — the C ABI of libggq.so (include/ggq.h): validation with the reference's check order,
// device-pointer entry points, and the host-pointer slice API with its H2D -> kernel -> D2H pipeline.
//
// There is no CPU fallback anywhere in this file: if CUDA is unavailable every compute entry
// point fails with GGQ_ERR_CUDA.
#include "../../include/ggq.h"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ggq_kernels.h"

namespace {

using namespace ggq;

thread_local std::string t_err;
thread_local int t_device = -1;  // ggq_set_device override for the calling thread
std::atomic<uint64_t> g_launches{0};

int fail(int code, const std::string &msg) {
    t_err = msg;
    return code;
}
int fail_cuda(cudaError_t e, const char *what) {
    return fail(GGQ_ERR_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}

struct TypeInfo { uint32_t type, elems, bytes; };
const TypeInfo TYPES[] = {
    {GGQ_F16, 1, 2},    {GGQ_BF16, 1, 2},   {GGQ_Q4_0, 32, 18}, {GGQ_Q4_1, 32, 20}, {GGQ_Q5_0, 32, 22},
    {GGQ_Q5_1, 32, 24}, {GGQ_Q8_0, 32, 34}, {GGQ_Q8_1, 32, 36}, {GGQ_Q2K, 256, 84}, {GGQ_Q3K, 256, 110},
    {GGQ_Q4K, 256, 144}, {GGQ_Q5K, 256, 176}, {GGQ_Q6K, 256, 210}, {GGQ_Q8K, 256, 290},
};
const TypeInfo *find_type(uint32_t t) {
    for (const auto &ti : TYPES)
        if (ti.type == t) return &ti;
    return nullptr;
}
size_t fdt_size(uint32_t fdt) { return fdt == GGQ_F32 ? 4 : (fdt == GGQ_F16 || fdt == GGQ_BF16) ? 2 : 0; }

// ---- device bookkeeping -------------------------------------------------------------------------
std::mutex g_dev_mu;
int g_sm_count[MAX_DEVICES];

int resolve_device(DevInfo *out) {
    int dev = t_device;
    if (dev >= 0) {
        cudaError_t e = cudaSetDevice(dev);
        if (e != cudaSuccess) return fail_cuda(e, "cudaSetDevice");
    } else {
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return fail_cuda(e, "cudaGetDevice (no CUDA device? libggq has no CPU fallback)");
    }
    if (dev < 0 || dev >= MAX_DEVICES) return fail(GGQ_ERR_INVALID, "device ordinal out of range");
    int sm;
    {
        std::lock_guard<std::mutex> lk(g_dev_mu);
        sm = g_sm_count[dev];
        if (sm == 0) {
            cudaError_t e = cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
            if (e != cudaSuccess) return fail_cuda(e, "cudaDeviceGetAttribute");
            g_sm_count[dev] = sm;
        }
    }
    out->device = dev;
    out->sm_count = sm;
    return GGQ_OK;
}

// ---- validated dispatch (shared by the host and device entry points) -----------------------------
struct Plan {
    const TypeInfo *ti;
    uint32_t fdt;
    size_t nblocks;
};

// lib.rs:121-127
int plan_quantize(uint32_t type, uint32_t fdt, size_t dst_blocks, size_t src_elems, Plan *p) {
    const TypeInfo *ti = find_type(type);
    if (!ti || !fdt_size(fdt)) return fail(GGQ_ERR_UNSUPPORTED, "unsupported block type or float dtype");
    if (src_elems % ti->elems != 0) return fail(GGQ_ERR_INDIVISIBLE, "src.len() % N != 0");
    if (dst_blocks != src_elems / ti->elems) return fail(GGQ_ERR_LENGTH_MISMATCH, "dst.len() != src.len() / N");
    *p = {ti, fdt, dst_blocks};
    return GGQ_OK;
}
// lib.rs:135-141
int plan_dequantize(uint32_t type, uint32_t fdt, size_t dst_elems, size_t src_blocks, Plan *p) {
    const TypeInfo *ti = find_type(type);
    if (!ti || !fdt_size(fdt)) return fail(GGQ_ERR_UNSUPPORTED, "unsupported block type or float dtype");
    if (dst_elems % ti->elems != 0) return fail(GGQ_ERR_INDIVISIBLE, "dst.len() % N != 0");
    if (src_blocks != dst_elems / ti->elems) return fail(GGQ_ERR_LENGTH_MISMATCH, "src.len() != dst.len() / N");
    *p = {ti, fdt, src_blocks};
    return GGQ_OK;
}

cudaError_t enqueue(bool quant, const Plan &p, void *d_dst, const void *d_src, size_t nblocks, cudaStream_t st, DevInfo dev) {
    const uint32_t ty = p.ti->type;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (ty == GGQ_F16 || ty == GGQ_BF16)  // 1-element blocks: structs/half.rs
        return quant ? cast_elems(p.fdt, ty, d_src, d_dst, nblocks, st, dev) : cast_elems(ty, p.fdt, d_src, d_dst, nblocks, st, dev);
    if (!quant) return dequant_blocks(ty, p.fdt, d_src, d_dst, nblocks, st, dev);
    if (ty >= GGQ_Q2K && ty <= GGQ_Q6K) return quant_blocks_k(ty, p.fdt, d_src, d_dst, nblocks, st, dev);
    return quant_blocks_legacy(ty, p.fdt, d_src, d_dst, nblocks, st, dev);
}

// ---- host pipeline -------------------------------------------------------------------------------
constexpr int NSLOTS = 3;
constexpr size_t CHUNK_ELEMS = size_t(1) << 23;          // 8 Mi elements of the float side per chunk
constexpr size_t SLOT_BYTES = CHUNK_ELEMS * 4 + 4096;    // either side of a chunk fits (<= 4 B/elem)

struct Slot {
    void *d_in = nullptr, *d_out = nullptr, *h_in = nullptr, *h_out = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
};
struct Pipeline {
    int device = -1;
    Slot slots[NSLOTS];
};

std::mutex g_pool_mu;
std::vector<Pipeline *> g_pool;  // idle pipelines (any device)

void destroy_pipeline(Pipeline *p) {
    for (auto &s : p->slots) {
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.done) cudaEventDestroy(s.done);
    }
    delete p;
}

int acquire_pipeline(int device, Pipeline **out) {
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (size_t i = 0; i < g_pool.size(); i++)
            if (g_pool[i]->device == device) {
                *out = g_pool[i];
                g_pool.erase(g_pool.begin() + i);
                return GGQ_OK;
            }
    }
    Pipeline *p = new Pipeline();
    p->device = device;
    for (auto &s : p->slots) {
        cudaError_t e;
        if ((e = cudaMalloc(&s.d_in, SLOT_BYTES)) != cudaSuccess || (e = cudaMalloc(&s.d_out, SLOT_BYTES)) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming)) != cudaSuccess) {
            destroy_pipeline(p);
            return fail_cuda(e, "pipeline setup");
        }
    }
    *out = p;
    return GGQ_OK;
}
void release_pipeline(Pipeline *p) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pool.push_back(p);
}

bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

void parallel_memcpy(void *dst, const void *src, size_t n) {
    constexpr size_t MIN_PER_THREAD = size_t(2) << 20;
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = n / MIN_PER_THREAD;
    if (nt > 8) nt = 8;
    if (hw && nt > hw) nt = hw;
    if (nt <= 1) {
        memcpy(dst, src, n);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = (n / nt + 63) & ~size_t(63);
    for (size_t t = 1; t < nt; t++) {
        const size_t o = t * per;
        if (o >= n) break;
        const size_t len = (o + per > n || t == nt - 1) ? n - o : per;
        th.emplace_back([=] { memcpy(static_cast<char *>(dst) + o, static_cast<const char *>(src) + o, len); });
    }
    memcpy(dst, src, per < n ? per : n);
    for (auto &t : th) t.join();
}

// Runs one host-pointer slice call: chunks of CHUNK_ELEMS elements flow through NSLOTS
// (stream, device in/out, pinned bounce in/out) slots so H2D(c+1), kernel(c) and D2H(c-1) overlap.
int run_host(bool quant, const Plan &p, void *dst, const void *src) {
    if (p.nblocks == 0) return GGQ_OK;
    if (!dst || !src) return fail(GGQ_ERR_INVALID, "null pointer with non-zero length");
    DevInfo dev;
    int rc = resolve_device(&dev);
    if (rc != GGQ_OK) return rc;
    Pipeline *pl = nullptr;
    if ((rc = acquire_pipeline(dev.device, &pl)) != GGQ_OK) return rc;

    const size_t fsz = fdt_size(p.fdt);
    const size_t blocks_per_chunk = CHUNK_ELEMS / p.ti->elems;
    const size_t nchunks = (p.nblocks + blocks_per_chunk - 1) / blocks_per_chunk;
    const size_t in_per_block = quant ? p.ti->elems * fsz : p.ti->bytes;
    const size_t out_per_block = quant ? p.ti->bytes : p.ti->elems * fsz;
    const bool pin_in = is_pinned(src), pin_out = is_pinned(dst);

    // lazily allocate bounce buffers only when the caller's memory is pageable
    cudaError_t e = cudaSuccess;
    for (auto &s : pl->slots) {
        if (!pin_in && !s.h_in && (e = cudaHostAlloc(&s.h_in, SLOT_BYTES, cudaHostAllocDefault)) != cudaSuccess) break;
        if (!pin_out && !s.h_out && (e = cudaHostAlloc(&s.h_out, SLOT_BYTES, cudaHostAllocDefault)) != cudaSuccess) break;
    }
    if (e != cudaSuccess) {
        release_pipeline(pl);
        return fail_cuda(e, "cudaHostAlloc");
    }

    auto chunk_blocks = [&](size_t c) { return (c + 1 == nchunks) ? p.nblocks - c * blocks_per_chunk : blocks_per_chunk; };
    const char *srcb = static_cast<const char *>(src);
    char *dstb = static_cast<char *>(dst);

    for (size_t c = 0; c < nchunks + NSLOTS && e == cudaSuccess; c++) {
        Slot &s = pl->slots[c % NSLOTS];
        if (c >= NSLOTS) {  // retire chunk c - NSLOTS
            const size_t r = c - NSLOTS;
            if ((e = cudaEventSynchronize(s.done)) != cudaSuccess) break;
            if (!pin_out) parallel_memcpy(dstb + r * blocks_per_chunk * out_per_block, s.h_out, chunk_blocks(r) * out_per_block);
        }
        if (c < nchunks) {
            const size_t nb = chunk_blocks(c);
            const char *hsrc = srcb + c * blocks_per_chunk * in_per_block;
            if (!pin_in) {
                parallel_memcpy(s.h_in, hsrc, nb * in_per_block);
                hsrc = static_cast<const char *>(s.h_in);
            }
            if ((e = cudaMemcpyAsync(s.d_in, hsrc, nb * in_per_block, cudaMemcpyHostToDevice, s.stream)) != cudaSuccess) break;
            if ((e = enqueue(quant, p, s.d_out, s.d_in, nb, s.stream, dev)) != cudaSuccess) break;
            void *hdst = pin_out ? static_cast<void *>(dstb + c * blocks_per_chunk * out_per_block) : s.h_out;
            if ((e = cudaMemcpyAsync(hdst, s.d_out, nb * out_per_block, cudaMemcpyDeviceToHost, s.stream)) != cudaSuccess) break;
            if ((e = cudaEventRecord(s.done, s.stream)) != cudaSuccess) break;
        }
    }
    if (e != cudaSuccess) {
        for (auto &s : pl->slots) cudaStreamSynchronize(s.stream);
        cudaGetLastError();
        release_pipeline(pl);
        return fail_cuda(e, quant ? "quantize_slice" : "dequantize_slice");
    }
    release_pipeline(pl);
    return GGQ_OK;
}

int run_device(bool quant, const Plan &p, void *dst, const void *src, void *stream) {
    if (p.nblocks == 0) return GGQ_OK;
    if (!dst || !src) return fail(GGQ_ERR_INVALID, "null pointer with non-zero length");
    DevInfo dev;
    int rc = resolve_device(&dev);
    if (rc != GGQ_OK) return rc;
    cudaError_t e = enqueue(quant, p, dst, src, p.nblocks, static_cast<cudaStream_t>(stream), dev);
    if (e == cudaErrorNotSupported) return fail(GGQ_ERR_UNSUPPORTED, "codec not implemented for this type");
    if (e != cudaSuccess) return fail_cuda(e, quant ? "quantize_slice_device" : "dequantize_slice_device");
    return GGQ_OK;
}

}  // namespace

extern "C" {

int ggq_block_info(uint32_t type, uint32_t *elems, uint32_t *bytes) {
    const TypeInfo *ti = find_type(type);
    if (!ti) return fail(GGQ_ERR_UNSUPPORTED, "unsupported block type");
    if (elems) *elems = ti->elems;
    if (bytes) *bytes = ti->bytes;
    return GGQ_OK;
}

const char *ggq_last_error(void) { return t_err.c_str(); }

int ggq_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int ggq_set_device(int device) {
    if (device < 0 || device >= ggq_device_count()) return fail(GGQ_ERR_INVALID, "no such CUDA device");
    t_device = device;
    return GGQ_OK;
}

int ggq_quantize_slice(uint32_t type, uint32_t fdt, void *dst, size_t dst_blocks, const void *src, size_t src_elems) {
    Plan p;
    int rc = plan_quantize(type, fdt, dst_blocks, src_elems, &p);
    return rc != GGQ_OK ? rc : run_host(true, p, dst, src);
}
int ggq_dequantize_slice(uint32_t type, uint32_t fdt, void *dst, size_t dst_elems, const void *src, size_t src_blocks) {
    Plan p;
    int rc = plan_dequantize(type, fdt, dst_elems, src_blocks, &p);
    return rc != GGQ_OK ? rc : run_host(false, p, dst, src);
}
int ggq_quantize_slice_device(uint32_t type, uint32_t fdt, void *dst, size_t dst_blocks, const void *src, size_t src_elems, void *stream) {
    Plan p;
    int rc = plan_quantize(type, fdt, dst_blocks, src_elems, &p);
    return rc != GGQ_OK ? rc : run_device(true, p, dst, src, stream);
}
int ggq_dequantize_slice_device(uint32_t type, uint32_t fdt, void *dst, size_t dst_elems, const void *src, size_t src_blocks, void *stream) {
    Plan p;
    int rc = plan_dequantize(type, fdt, dst_elems, src_blocks, &p);
    return rc != GGQ_OK ? rc : run_device(false, p, dst, src, stream);
}

void *ggq_host_alloc(size_t bytes) {
    void *p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        fail_cuda(e, "cudaHostAlloc");
        return nullptr;
    }
    return p;
}
void ggq_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

uint64_t ggq_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char *ggq_version(void) { return "ggq-b200 0.1 (sm_100a)"; }

}  // extern "C"
