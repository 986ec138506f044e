// api.cu — the C ABI of libggq.so (include/ggq.h): validation with the reference's check order,
// device-pointer entry points, and the host-pointer slice API with its H2D -> kernel -> D2H pipeline.
//
// There is no CPU fallback anywhere in this file: if CUDA is unavailable every compute entry
// point fails with GGQ_ERR_CUDA.
#include "../../include/ggq.h"

#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ggq_internal.h"
#include "ggq_kernels.h"

namespace {

using namespace ggq;

thread_local std::string t_err;
thread_local ggq::PipeCounters t_pipe;  // see ggq_internal.h
uint64_t now_ns() { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
cudaError_t timed_event_sync(cudaEvent_t ev) {  // cudaEventSynchronize, counted as time blocked on the GPU
    const uint64_t t0 = now_ns();
    const cudaError_t e = cudaEventSynchronize(ev);
    t_pipe.gpu_wait_ns += now_ns() - t0;
    return e;
}
thread_local int t_device = -1;  // ggq_set_device override for the calling thread
std::atomic<uint64_t> g_launches{0};
thread_local int t_shard_devices = 1;  // ggq_set_shard_devices: GPUs the calling thread's host-pointer calls are split over

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

// A "tensor type" is F32 or one of the block types (F16 / BF16 count as 1-element blocks).
bool is_float_type(uint32_t t) { return t == GGQ_F32 || t == GGQ_F16 || t == GGQ_BF16; }
bool type_geometry(uint32_t t, size_t *elems, size_t *bytes) {
    if (t == GGQ_F32) { *elems = 1; *bytes = 4; return true; }
    const TypeInfo *ti = find_type(t);
    if (!ti) return false;
    *elems = ti->elems;
    *bytes = ti->bytes;
    return true;
}
size_t type_nbytes(uint32_t t, size_t n_elems) {
    size_t e = 1, b = 1;
    type_geometry(t, &e, &b);
    return n_elems / e * b;
}

// one kernel launch: `from` -> `to` over n_elems elements; never block -> block (see expand_chain)
cudaError_t enqueue_hop(uint32_t from, uint32_t to, void *d_dst, const void *d_src, size_t n_elems, cudaStream_t st, DevInfo dev) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (is_float_type(from) && is_float_type(to)) return cast_elems(from, to, d_src, d_dst, n_elems, st, dev);  // structs/half.rs
    if (is_float_type(from)) {  // quantize::<to, from, N>
        const size_t nb = n_elems / find_type(to)->elems;
        if (to >= GGQ_Q2K && to <= GGQ_Q6K) return quant_blocks_k(to, from, d_src, d_dst, nb, st, dev);
        return quant_blocks_legacy(to, from, d_src, d_dst, nb, st, dev);
    }
    if (is_float_type(to)) return dequant_blocks(from, to, d_src, d_dst, n_elems / find_type(from)->elems, st, dev);
    return cudaErrorInvalidValue;
}

// cast.rs:136 `_ => cast(row, &cast(row, data, from, Ty::F32), Ty::F32, to)`: a quantized source is
// mediated by F32.  Equal neighbours are dropped.
int expand_chain(const uint32_t *types, int n, std::vector<uint32_t> *out) {
    if (n < 1) return fail(GGQ_ERR_INVALID, "empty cast chain");
    out->clear();
    for (int i = 0; i < n; i++) {
        size_t e, b;
        if (!type_geometry(types[i], &e, &b)) return fail(GGQ_ERR_UNSUPPORTED, "unsupported tensor type in cast chain");
        if (!out->empty() && out->back() == types[i]) continue;
        if (!out->empty() && !is_float_type(out->back()) && !is_float_type(types[i])) out->push_back(GGQ_F32);
        out->push_back(types[i]);
    }
    return GGQ_OK;
}

// ---- host pipeline -------------------------------------------------------------------------------
constexpr int NSLOTS = 3;
constexpr size_t CHUNK_ELEMS = size_t(1) << 23;          // 8 Mi elements per chunk (multiple of every block size)
constexpr size_t FIRST_CHUNK_ELEMS = size_t(1) << 20;    // ramp: 1, 2, 4, 8, 8, ... Mi elements
constexpr size_t SLOT_BYTES = CHUNK_ELEMS * 4 + 16384;   // any representation of a chunk fits (<= 4 B/elem) + slack for 4 KiB-aligned direct reads

struct Slot {
    void *d_a = nullptr, *d_b = nullptr, *h_in = nullptr, *h_out = nullptr;  // device ping-pong, pinned bounce
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
        if (s.d_a) cudaFree(s.d_a);
        if (s.d_b) cudaFree(s.d_b);
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
        if ((e = cudaMalloc(&s.d_a, SLOT_BYTES)) != cudaSuccess || (e = cudaMalloc(&s.d_b, SLOT_BYTES)) != cudaSuccess ||
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

// cudaMemoryTypeHost = page-locked (DMA-able); Unregistered = ordinary pageable memory
cudaMemoryType pointer_kind(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return cudaMemoryTypeUnregistered;
    }
    return a.type;
}
bool is_pinned(const void *p) { return pointer_kind(p) == cudaMemoryTypeHost; }

// ---- persistent copy pool: bounce copies between pageable caller memory and the pinned staging
// buffers are split over a few long-lived worker threads (spawning threads per 16-32 MB chunk cost as
// much as the copy itself).  The pool is created on first use and intentionally never destroyed.
class CopyPool {
   public:
    static CopyPool &get() {
        static CopyPool *p = new CopyPool();
        return *p;
    }
    int size() const { return (int)workers_.size(); }
    // run fn(0..parts-1), part 0 on the calling thread, and return when all parts are done
    void run(int parts, const std::function<void(int)> &fn) {
        if (parts <= 1) { if (parts == 1) fn(0); return; }
        struct Job { std::mutex mu; std::condition_variable cv; int pending; } job;
        job.pending = parts - 1;
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (int i = 1; i < parts; i++)
                q_.push_back([&job, &fn, i] {
                    fn(i);
                    std::lock_guard<std::mutex> lk(job.mu);
                    if (--job.pending == 0) job.cv.notify_one();
                });
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> lk(job.mu);
        job.cv.wait(lk, [&] { return job.pending == 0; });
    }

   private:
    CopyPool() {
        unsigned hw = std::thread::hardware_concurrency();
        // the copies use non-temporal stores and keep scaling up to the core count on the 16-core B200 hosts
        // (pageable Q8_0 dequantize: 40 / 56 / 65 GB/s with 5 / 8 / 15 helpers)
        int n = hw ? (int)std::min(16u, std::max(2u, hw - 1)) : 4;
        if (const char *env = getenv("GGQ_COPY_THREADS")) {  // tuning knob: helper threads for pageable bounce copies
            const int v = atoi(env);
            if (v >= 1 && v <= 64) n = v;
        }
        for (int i = 0; i < n; i++)
            workers_.emplace_back([this] {
                for (;;) {
                    std::function<void()> f;
                    {
                        std::unique_lock<std::mutex> lk(mu_);
                        cv_.wait(lk, [&] { return !q_.empty(); });
                        f = std::move(q_.front());
                        q_.pop_front();
                    }
                    f();
                }
            });
        for (auto &t : workers_) t.detach();
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    std::vector<std::thread> workers_;
};

void parallel_memcpy(void *dst, const void *src, size_t n) {
    constexpr size_t MIN_PER_PART = size_t(1) << 20;
    CopyPool &pool = CopyPool::get();
    const int parts = (int)std::min<size_t>((size_t)pool.size() + 1, std::max<size_t>(1, n / MIN_PER_PART));
    if (parts <= 1) {
        ggq::stream_copy(dst, src, n);
        return;
    }
    const size_t per = ((n + parts - 1) / parts + 4095) & ~size_t(4095);
    pool.run(parts, [=](int k) {
        const size_t o = (size_t)k * per;
        if (o < n) ggq::stream_copy(static_cast<char *>(dst) + o, static_cast<const char *>(src) + o, std::min(per, n - o));
    });
}

// Runs one cast chain over `n_elems` elements: chunks of CHUNK_ELEMS elements flow through NSLOTS
// (stream, device ping-pong, pinned in/out) slots so H2D(c+1), kernels(c) and D2H(c-1) overlap;
// intermediates of a multi-hop chain never leave the device.  `io` supplies and consumes the bytes:
//   direct_src / direct_dst  non-null => pinned caller memory, DMA straight from / to it;
//   otherwise read(pinned, byte_off, nbytes) fills a pinned bounce buffer and
//   write(pinned, byte_off, nbytes) drains one (memcpy for pageable memory, pread/pwrite for files).
struct ChainJob {
    std::vector<uint32_t> chain;  // >= 2 types
    size_t n_elems;
    ggq::ChainIO io;
    size_t elem_base = 0;  // this job covers elements [elem_base, elem_base + n_elems) of the tensor `io` addresses
};

// Streams every job, back to back, through ONE pipeline: the D2H of job k's last chunks overlaps the
// H2D / kernels of job k+1's first chunks (no drain between tensors).
int run_jobs_io(const std::vector<ChainJob> &jobs) {
    struct Chunk { uint32_t job; size_t e0, e1; };
    std::vector<Chunk> chunks;
    // Chunk boundaries (elements).  The first chunks of the call are small (1, 2, 4 Mi elements) so
    // the first D2H starts early; steady state uses CHUNK_ELEMS.
    size_t ramp = FIRST_CHUNK_ELEMS;
    for (uint32_t j = 0; j < jobs.size(); j++)
        for (size_t e = 0; e < jobs[j].n_elems;) {
            const size_t n = std::min(jobs[j].n_elems - e, ramp);
            chunks.push_back({j, e, e + n});
            e += n;
            ramp = std::min(ramp * 2, CHUNK_ELEMS);
        }
    if (chunks.empty()) return GGQ_OK;
    DevInfo dev;
    int rc = resolve_device(&dev);
    if (rc != GGQ_OK) return rc;
    Pipeline *pl = nullptr;
    if ((rc = acquire_pipeline(dev.device, &pl)) != GGQ_OK) return rc;

    // lazily allocate bounce buffers only when some job needs them
    bool need_in = false, need_out = false;
    for (const auto &j : jobs) { need_in |= !j.io.direct_src; need_out |= !j.io.direct_dst; }
    cudaError_t e = cudaSuccess;
    for (auto &s : pl->slots) {
        if (need_in && !s.h_in && (e = cudaHostAlloc(&s.h_in, SLOT_BYTES, cudaHostAllocDefault)) != cudaSuccess) break;
        if (need_out && !s.h_out && (e = cudaHostAlloc(&s.h_out, SLOT_BYTES, cudaHostAllocDefault)) != cudaSuccess) break;
    }
    if (e != cudaSuccess) {
        release_pipeline(pl);
        return fail_cuda(e, "cudaHostAlloc");
    }

    const size_t nchunks = chunks.size();
    bool io_ok = true;
    for (size_t c = 0; c < nchunks + NSLOTS && e == cudaSuccess && io_ok; c++) {
        Slot &s = pl->slots[c % NSLOTS];
        if (c >= NSLOTS) {  // retire chunk c - NSLOTS
            const Chunk &r = chunks[c - NSLOTS];
            const ChainJob &rj = jobs[r.job];
            if ((e = timed_event_sync(s.done)) != cudaSuccess) break;
            const uint32_t t_out = rj.chain.back();
            if (!rj.io.direct_dst) io_ok = rj.io.write(s.h_out, type_nbytes(t_out, rj.elem_base + r.e0), type_nbytes(t_out, r.e1 - r.e0));
        }
        if (c < nchunks && io_ok) {
            const Chunk &k = chunks[c];
            const ChainJob &job = jobs[k.job];
            const uint32_t t_in = job.chain.front(), t_out = job.chain.back();
            const size_t ne = k.e1 - k.e0;
            const size_t in_off = type_nbytes(t_in, job.elem_base + k.e0), out_off = type_nbytes(t_out, job.elem_base + k.e0);
            const char *hsrc;
            if (job.io.direct_src) {
                hsrc = static_cast<const char *>(job.io.direct_src) + in_off;
            } else if (job.io.read_shift) {
                const long shift = job.io.read_shift(s.h_in, SLOT_BYTES, in_off, type_nbytes(t_in, ne));
                if (!(io_ok = shift >= 0)) break;
                hsrc = static_cast<const char *>(s.h_in) + shift;
            } else {
                if (!(io_ok = job.io.read(s.h_in, in_off, type_nbytes(t_in, ne)))) break;
                hsrc = static_cast<const char *>(s.h_in);
            }
            if ((e = cudaMemcpyAsync(s.d_a, hsrc, type_nbytes(t_in, ne), cudaMemcpyHostToDevice, s.stream)) != cudaSuccess) break;
            t_pipe.h2d_bytes += type_nbytes(t_in, ne);
            t_pipe.d2h_bytes += type_nbytes(t_out, ne);
            void *cur = s.d_a, *nxt = s.d_b;
            for (size_t h = 0; h + 1 < job.chain.size(); h++) {
                if ((e = enqueue_hop(job.chain[h], job.chain[h + 1], nxt, cur, ne, s.stream, dev)) != cudaSuccess) break;
                std::swap(cur, nxt);
            }
            if (e != cudaSuccess) break;
            void *hdst = job.io.direct_dst ? static_cast<void *>(static_cast<char *>(job.io.direct_dst) + out_off) : s.h_out;
            if ((e = cudaMemcpyAsync(hdst, cur, type_nbytes(t_out, ne), cudaMemcpyDeviceToHost, s.stream)) != cudaSuccess) break;
            if ((e = cudaEventRecord(s.done, s.stream)) != cudaSuccess) break;
        }
    }
    if (e != cudaSuccess || !io_ok) {
        for (auto &s : pl->slots) cudaStreamSynchronize(s.stream);
        cudaGetLastError();
        release_pipeline(pl);
        return e != cudaSuccess ? fail_cuda(e, "host cast pipeline") : fail(GGQ_ERR_INVALID, "I/O callback failed in the cast pipeline");
    }
    release_pipeline(pl);
    return GGQ_OK;
}

// One job through run_jobs_io.
int run_chain_io(const std::vector<uint32_t> &chain, size_t n_elems, const ggq::ChainIO &io) {
    if (n_elems == 0) return GGQ_OK;
    std::vector<ChainJob> jobs(1);
    jobs[0].chain = chain;
    jobs[0].n_elems = n_elems;
    jobs[0].io = io;
    return run_jobs_io(jobs);
}

// A pageable destination is typically what the reference's caller just mapped (cast.rs:158-161 `MmapMut::map_anon`):
// untouched anonymous memory, so the bounce copies out of the pinned staging buffers take one page fault per 4 KiB
// (measured: 604 MB of f16 output fault in at ~7 GB/s whatever the thread count, CPU reference and this library
// alike — the kernel's fault path, not the copy, is the bound).  Asking for transparent huge pages on the 2 MiB-
// aligned interior turns that into one fault per 2 MiB, zeroed by whichever copy thread touches it first.  A hint
// only: contents and mapping stay the caller's, errors are ignored, already-populated ranges are unaffected.
// (Populating an untouched destination ahead of the copies on background threads — MADV_POPULATE_WRITE in 32 MiB
// pieces while the first chunks are in flight — was measured and is not faster: 26.0-26.2 against 26.8-27.9 GB/s for
// the bench step; zeroing the pages costs the same memory bandwidth whoever does it.)
void hint_huge_pages(void *dst, size_t nbytes) {
#ifdef MADV_HUGEPAGE
    static const bool off = getenv("GGQ_NO_THP_HINT") != nullptr;
    constexpr uintptr_t HP = uintptr_t(2) << 20;
    if (off || nbytes < 4 * HP) return;
    const uintptr_t a = (reinterpret_cast<uintptr_t>(dst) + HP - 1) & ~(HP - 1);
    const uintptr_t b = (reinterpret_cast<uintptr_t>(dst) + nbytes) & ~(HP - 1);
    if (b > a) madvise(reinterpret_cast<void *>(a), b - a, MADV_HUGEPAGE);
#endif
}

// memcpy-backed ChainIO over caller memory (pinned => direct DMA)
int make_mem_io(void *dst, size_t dst_bytes, const void *src, ggq::ChainIO *io) {
    const cudaMemoryType ks = pointer_kind(src), kd = pointer_kind(dst);
    if (ks == cudaMemoryTypeDevice || kd == cudaMemoryTypeDevice)
        return fail(GGQ_ERR_INVALID, "device pointer passed to a host-pointer entry point (use the *_device variants)");
    if (kd == cudaMemoryTypeUnregistered) hint_huge_pages(dst, dst_bytes);
    io->direct_src = ks == cudaMemoryTypeHost ? src : nullptr;
    io->direct_dst = kd == cudaMemoryTypeHost ? dst : nullptr;
    const char *s0 = static_cast<const char *>(src);
    char *d0 = static_cast<char *>(dst);
    io->read = [s0](void *pinned, size_t off, size_t n) { parallel_memcpy(pinned, s0 + off, n); return true; };
    io->write = [d0](const void *pinned, size_t off, size_t n) { parallel_memcpy(d0 + off, pinned, n); return true; };
    return GGQ_OK;
}

// ---- the ONE multi-GPU partitioner (ggq_plan_shards) -----------------------------------------------
// Blocks are independent (lib.rs:129-131), so a call's jobs are split "by tensor and by block range" with a
// single rule: lay the jobs end to end in SHARD_UNIT-element units (a multiple of every block size and of 8
// blocks, so every cut keeps both sides 16-byte aligned), weigh a unit by the bytes it moves over PCIe
// (input + output representation), and give device d the units whose weight midpoint falls in
// [W d / n, W (d+1) / n).  Consecutive units of one job on one device form one piece.  Each device then streams
// its pieces through its own pipeline over its own PCIe link; there is no inter-GPU traffic.  A call moving
// less than MIN_BYTES_PER_DEVICE per device uses fewer devices.
constexpr size_t SHARD_UNIT = size_t(1) << 20;
constexpr uint64_t MIN_BYTES_PER_DEVICE = uint64_t(16) << 20;
struct Piece { uint32_t job; int device; size_t e0, e1; };
struct JobGeom { size_t n_elems; uint32_t t_in, t_out; };

std::vector<Piece> plan_pieces(const std::vector<JobGeom> &jobs, int ndev) {
    std::vector<Piece> out;
    uint64_t total = 0;
    for (const auto &j : jobs) total += type_nbytes(j.t_in, j.n_elems) + type_nbytes(j.t_out, j.n_elems);
    if (ndev < 1) ndev = 1;
    ndev = (int)std::min<uint64_t>((uint64_t)ndev, std::max<uint64_t>(1, total / MIN_BYTES_PER_DEVICE));
    uint64_t before = 0;  // weight of the units already placed
    for (uint32_t ji = 0; ji < jobs.size(); ji++) {
        const JobGeom &j = jobs[ji];
        for (size_t e = 0; e < j.n_elems; e += SHARD_UNIT) {
            const size_t ne = std::min(SHARD_UNIT, j.n_elems - e);
            const uint64_t w = type_nbytes(j.t_in, ne) + type_nbytes(j.t_out, ne);
            // device of the unit's midpoint: floor((before + w/2) * ndev / total), in 128-bit-safe arithmetic
            const unsigned __int128 mid2 = (unsigned __int128)(2 * before + w) * (unsigned)ndev;
            int d = total ? (int)(mid2 / (2 * (unsigned __int128)total)) : 0;
            if (d >= ndev) d = ndev - 1;
            before += w;
            if (!out.empty() && out.back().job == ji && out.back().device == d && out.back().e1 == e) out.back().e1 = e + ne;
            else out.push_back({ji, d, e, e + ne});
        }
    }
    return out;
}

// Runs the jobs on the calling thread's device, or — after ggq_set_shard_devices(n > 1) on this thread — split
// over devices 0..n-1 by plan_pieces().  The caller's CUDA device and ggq_set_device() state are restored.
int run_jobs_sharded(const std::vector<ChainJob> &jobs) {
    const int ndev = (t_device >= 0) ? 1 : t_shard_devices;
    if (ndev <= 1 || jobs.empty()) return run_jobs_io(jobs);
    std::vector<JobGeom> geom;
    for (const auto &j : jobs) geom.push_back({j.n_elems, j.chain.front(), j.chain.back()});
    const std::vector<Piece> pieces = plan_pieces(geom, ndev);
    int used = 0;
    for (const auto &p : pieces) used = std::max(used, p.device + 1);
    if (used <= 1) return run_jobs_io(jobs);
    std::vector<std::vector<ChainJob>> per_dev(used);
    for (const auto &p : pieces) {
        ChainJob c = jobs[p.job];
        c.elem_base = jobs[p.job].elem_base + p.e0;
        c.n_elems = p.e1 - p.e0;
        per_dev[p.device].push_back(std::move(c));
    }
    std::vector<int> rcs(used, GGQ_OK);
    std::vector<std::string> errs(used);
    auto dev_fn = [&](int d) {
        t_device = d;  // pin this thread to device d
        rcs[d] = run_jobs_io(per_dev[d]);
        if (rcs[d] != GGQ_OK) errs[d] = t_err;
    };
    const int saved = t_device;
    int saved_cuda = -1;
    if (cudaGetDevice(&saved_cuda) != cudaSuccess) { cudaGetLastError(); saved_cuda = -1; }
    std::vector<std::thread> th;
    for (int d = 1; d < used; d++) th.emplace_back(dev_fn, d);
    dev_fn(0);
    t_device = saved;
    if (saved_cuda >= 0) cudaSetDevice(saved_cuda);
    for (auto &t : th) t.join();
    for (int d = 0; d < used; d++)
        if (rcs[d] != GGQ_OK) return fail(rcs[d], errs[d]);
    return GGQ_OK;
}

int run_chain_host(const std::vector<uint32_t> &chain, void *dst, const void *src, size_t n_elems) {
    if (n_elems == 0) return GGQ_OK;
    if (!dst || !src) return fail(GGQ_ERR_INVALID, "null pointer with non-zero length");
    if (chain.size() == 1) {  // same type: plain copy (cast.rs never calls this; kept total)
        memmove(dst, src, type_nbytes(chain.front(), n_elems));
        return GGQ_OK;
    }
    std::vector<ChainJob> jobs(1);
    jobs[0].chain = chain;
    jobs[0].n_elems = n_elems;
    int rc = make_mem_io(dst, type_nbytes(chain.back(), n_elems), src, &jobs[0].io);  // rejects device pointers, notes which side is pinned
    if (rc != GGQ_OK) return rc;
    return run_jobs_sharded(jobs);
}

int run_host(bool quant, const Plan &p, void *dst, const void *src) {
    std::vector<uint32_t> chain = quant ? std::vector<uint32_t>{p.fdt, p.ti->type} : std::vector<uint32_t>{p.ti->type, p.fdt};
    if (chain[0] == chain[1]) chain.insert(chain.begin() + 1, GGQ_F32);  // e.g. quantize::<f16, f16, 1>: still mediated by f32 (lib.rs:66-73)
    return run_chain_host(chain, dst, src, p.nblocks * p.ti->elems);
}

int run_device(bool quant, const Plan &p, void *dst, const void *src, void *stream) {
    if (p.nblocks == 0) return GGQ_OK;
    if (!dst || !src) return fail(GGQ_ERR_INVALID, "null pointer with non-zero length");
    // cast.rs:163-177 `reslice` panics on a misaligned slice; here: the float side needs its element
    // alignment, the packed side 2 bytes (every block starts with or contains f16 fields)
    const void *fl = quant ? src : dst, *pk = quant ? dst : src;
    if ((reinterpret_cast<uintptr_t>(fl) & (fdt_size(p.fdt) - 1)) || (reinterpret_cast<uintptr_t>(pk) & 1))
        return fail(GGQ_ERR_INVALID, "data is not aligned");
    DevInfo dev;
    int rc = resolve_device(&dev);
    if (rc != GGQ_OK) return rc;
    const uint32_t from = quant ? p.fdt : p.ti->type, to = quant ? p.ti->type : p.fdt;
    cudaError_t e = enqueue_hop(from, to, dst, src, p.nblocks * p.ti->elems, static_cast<cudaStream_t>(stream), dev);
    if (e != cudaSuccess) return fail_cuda(e, quant ? "quantize_slice_device" : "dequantize_slice_device");
    return GGQ_OK;
}

// cast.rs:93-138 `cast(row, data, from, to)` for every pair of supported tensor types
int plan_cast(const uint32_t *types, int n_types, size_t n_elems, std::vector<uint32_t> *chain) {
    int rc = expand_chain(types, n_types, chain);
    if (rc != GGQ_OK) return rc;
    for (uint32_t t : *chain) {
        size_t e, b;
        type_geometry(t, &e, &b);
        if (n_elems % e != 0) return fail(GGQ_ERR_INDIVISIBLE, "element count is not a multiple of a block size in the chain");
    }
    return GGQ_OK;
}

// ---- rearrange (mem-rearrange's Rearranging) ---------------------------------------------------------
int plan_rearrange(const ggq_layout *dl, const ggq_layout *sl, size_t unit, StridedLayout *d, StridedLayout *s) {
    if (!dl || !sl) return fail(GGQ_ERR_INVALID, "null layout");
    if (dl->ndim > GGQ_MAX_NDIM || sl->ndim > GGQ_MAX_NDIM) return fail(GGQ_ERR_INVALID, "layout has more than GGQ_MAX_NDIM dims");
    if (unit == 0) return fail(GGQ_ERR_INVALID, "unit is zero");
    if (dl->ndim != sl->ndim) return fail(GGQ_ERR_LENGTH_MISMATCH, "ShapeMismatch: dst and src layouts differ in ndim");
    d->ndim = s->ndim = (int)dl->ndim;
    d->offset = dl->offset;
    s->offset = sl->offset;
    for (uint32_t i = 0; i < dl->ndim; i++) {
        if (dl->shape[i] != sl->shape[i]) return fail(GGQ_ERR_LENGTH_MISMATCH, "ShapeMismatch: dst and src layouts differ in shape");
        if (dl->shape[i] > 1 && dl->strides[i] == 0) return fail(GGQ_ERR_INVALID, "dst layout has a zero stride on a dim of extent > 1");
        d->shape[i] = s->shape[i] = dl->shape[i];
        d->strides[i] = dl->strides[i];
        s->strides[i] = sl->strides[i];
    }
    return GGQ_OK;
}

// [lo, hi) byte range a layout addresses relative to its base pointer; false when it addresses nothing
bool layout_span(const ggq_layout &l, size_t unit, int64_t *lo, int64_t *hi, uint64_t *addressed) {
    int64_t a = l.offset, b = l.offset;
    uint64_t n = unit;
    for (uint32_t i = 0; i < l.ndim; i++) {
        if (l.shape[i] == 0) return false;
        const int64_t ext = (int64_t)(l.shape[i] - 1) * l.strides[i];
        (ext < 0 ? a : b) += ext;
        n *= l.shape[i];
    }
    *lo = a;
    *hi = b + (int64_t)unit;
    *addressed = n;
    return true;
}

int rearrange_on_stream(void *dst, const ggq_layout *dl, const void *src, const ggq_layout *sl, size_t unit, cudaStream_t st) {
    StridedLayout d, s;
    int rc = plan_rearrange(dl, sl, unit, &d, &s);
    if (rc != GGQ_OK) return rc;
    uint64_t launches = 0;
    cudaError_t e = rearrange_strided(dst, d, src, s, unit, st, &launches);
    g_launches.fetch_add(launches, std::memory_order_relaxed);
    if (e != cudaSuccess) return fail_cuda(e, "rearrange");
    return GGQ_OK;
}

std::once_flag g_pool_cfg[MAX_DEVICES];
void configure_mem_pool(int device) {  // keep freed stream-ordered allocations cached instead of returning them to the OS
    std::call_once(g_pool_cfg[device], [device] {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    });
}

}  // namespace

namespace ggq {

PipeCounters take_pipe_counters() {
    const PipeCounters c = t_pipe;
    t_pipe = PipeCounters();
    return c;
}

size_t device_free_bytes() {
    DevInfo dev;
    if (resolve_device(&dev) != GGQ_OK) return 0;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return free_b;
}

struct Resident::Impl {
    DevInfo dev;
    Pipeline *pl = nullptr;
    cudaStream_t stream = nullptr;
    std::vector<void *> live;
};

Resident::Resident() : impl_(new Impl()) {
    if ((status_ = resolve_device(&impl_->dev)) != GGQ_OK) return;
    if ((status_ = acquire_pipeline(impl_->dev.device, &impl_->pl)) != GGQ_OK) return;
    impl_->stream = impl_->pl->slots[0].stream;
    configure_mem_pool(impl_->dev.device);
    cudaError_t e = cudaSuccess;
    for (auto &s : impl_->pl->slots) {
        if (!s.h_in && (e = cudaHostAlloc(&s.h_in, SLOT_BYTES, cudaHostAllocDefault)) != cudaSuccess) break;
        if (!s.h_out && (e = cudaHostAlloc(&s.h_out, SLOT_BYTES, cudaHostAllocDefault)) != cudaSuccess) break;
    }
    if (e != cudaSuccess) status_ = fail_cuda(e, "cudaHostAlloc");
}

Resident::~Resident() {
    if (impl_->pl) {
        for (void *p : impl_->live) cudaFreeAsync(p, impl_->stream);
        cudaStreamSynchronize(impl_->stream);
        cudaGetLastError();
        release_pipeline(impl_->pl);
    }
    delete impl_;
}

int Resident::alloc(size_t nbytes, void **d) {
    cudaError_t e = cudaMallocAsync(d, nbytes ? nbytes : 1, impl_->stream);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMallocAsync");
    impl_->live.push_back(*d);
    return GGQ_OK;
}

void Resident::free(void *d) {
    auto it = std::find(impl_->live.begin(), impl_->live.end(), d);
    if (it == impl_->live.end()) return;
    impl_->live.erase(it);
    cudaFreeAsync(d, impl_->stream);
}

// pread (or memcpy) chunk c+1 into a pinned buffer while chunk c's H2D runs
int Resident::upload(void *d_dst, size_t nbytes, const ReadFn &read) {
    constexpr size_t CH = size_t(16) << 20;
    size_t c = 0;
    for (size_t off = 0; off < nbytes; off += CH, c++) {
        Slot &s = impl_->pl->slots[c % NSLOTS];
        const size_t n = std::min(CH, nbytes - off);
        cudaError_t e = timed_event_sync(s.done);  // the H2D that last used this bounce buffer
        if (e != cudaSuccess) return fail_cuda(e, "upload");
        t_pipe.h2d_bytes += n;
        if (!read(s.h_in, off, n)) return fail(GGQ_ERR_INVALID, "I/O callback failed while uploading a tensor");
        if ((e = cudaMemcpyAsync(static_cast<char *>(d_dst) + off, s.h_in, n, cudaMemcpyHostToDevice, impl_->stream)) != cudaSuccess ||
            (e = cudaEventRecord(s.done, impl_->stream)) != cudaSuccess)
            return fail_cuda(e, "upload");
    }
    return GGQ_OK;
}

// D2H of chunk c+1 and c+2 run while chunk c is handed to `write`
int Resident::download(const void *d_src, size_t nbytes, const WriteFn &write) {
    constexpr size_t CH = size_t(16) << 20;
    const size_t nch = (nbytes + CH - 1) / CH;
    for (size_t c = 0; c < nch + (NSLOTS - 1); c++) {
        if (c < nch) {
            Slot &s = impl_->pl->slots[c % NSLOTS];
            const size_t off = c * CH, n = std::min(CH, nbytes - off);
            cudaError_t e;
            if ((e = cudaMemcpyAsync(s.h_out, static_cast<const char *>(d_src) + off, n, cudaMemcpyDeviceToHost, impl_->stream)) != cudaSuccess ||
                (e = cudaEventRecord(s.done, impl_->stream)) != cudaSuccess)
                return fail_cuda(e, "download");
        }
        if (c >= NSLOTS - 1) {
            const size_t r = c - (NSLOTS - 1);
            Slot &s = impl_->pl->slots[r % NSLOTS];
            const size_t off = r * CH, n = std::min(CH, nbytes - off);
            cudaError_t e = timed_event_sync(s.done);
            if (e != cudaSuccess) return fail_cuda(e, "download");
            t_pipe.d2h_bytes += n;
            if (!write(s.h_out, off, n)) return fail(GGQ_ERR_INVALID, "I/O callback failed while downloading a tensor");
        }
    }
    return GGQ_OK;
}

int Resident::cast(const uint32_t *types, int n_types, size_t n_elems, void *d_dst, const void *d_src) {
    std::vector<uint32_t> chain;
    int rc = plan_cast(types, n_types, n_elems, &chain);
    if (rc != GGQ_OK) return rc;
    if (chain.size() == 1) {
        cudaError_t e = cudaMemcpyAsync(d_dst, d_src, type_nbytes(chain[0], n_elems), cudaMemcpyDeviceToDevice, impl_->stream);
        return e == cudaSuccess ? GGQ_OK : fail_cuda(e, "cast (copy)");
    }
    const void *cur = d_src;
    void *tmp_prev = nullptr;
    for (size_t h = 0; h + 1 < chain.size(); h++) {
        void *out = d_dst;
        if (h + 2 < chain.size() && (rc = alloc(type_nbytes(chain[h + 1], n_elems), &out)) != GGQ_OK) return rc;
        cudaError_t e = enqueue_hop(chain[h], chain[h + 1], out, cur, n_elems, impl_->stream, impl_->dev);
        if (e != cudaSuccess) return fail_cuda(e, "cast");
        if (tmp_prev) free(tmp_prev);
        tmp_prev = out == d_dst ? nullptr : out;
        cur = out;
    }
    return GGQ_OK;
}

int Resident::rearrange(void *d_dst, const ggq_layout &dl, const void *d_src, const ggq_layout &sl, size_t unit) {
    return rearrange_on_stream(d_dst, &dl, d_src, &sl, unit, impl_->stream);
}

// internal (not part of the C ABI): used by convert.cpp to stream files through the pipeline
int cast_chain_io(const uint32_t *types, int n_types, size_t n_elems, const ChainIO &io) {
    std::vector<uint32_t> chain;
    int rc = plan_cast(types, n_types, n_elems, &chain);
    if (rc != GGQ_OK) return rc;
    if (chain.size() < 2) return fail(GGQ_ERR_INVALID, "cast_chain_io needs at least one hop");
    return run_chain_io(chain, n_elems, io);
}
}  // namespace ggq

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

int ggq_set_shard_devices(int n_devices) {
    const int avail = ggq_device_count();
    if (n_devices <= 0) n_devices = avail;
    if (n_devices > avail) return fail(GGQ_ERR_INVALID, "more shard devices than CUDA devices");
    if (n_devices < 1) n_devices = 1;
    t_shard_devices = n_devices;
    return n_devices;
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

int ggq_slices(const struct ggq_slice_job *jobs, size_t n_jobs) {
    if (n_jobs && !jobs) return fail(GGQ_ERR_INVALID, "null job table");
    std::vector<ChainJob> cj;
    for (size_t i = 0; i < n_jobs; i++) {
        const ggq_slice_job &j = jobs[i];
        Plan p;
        int rc = j.quantize ? plan_quantize(j.type, j.fdt, j.dst_len, j.src_len, &p) : plan_dequantize(j.type, j.fdt, j.dst_len, j.src_len, &p);
        if (rc != GGQ_OK) return rc;  // nothing has been computed yet: all jobs are validated first
        if (p.nblocks == 0) continue;
        if (!j.dst || !j.src) return fail(GGQ_ERR_INVALID, "null pointer with non-zero length");
        ChainJob c;
        c.chain = j.quantize ? std::vector<uint32_t>{p.fdt, p.ti->type} : std::vector<uint32_t>{p.ti->type, p.fdt};
        if (c.chain[0] == c.chain[1]) c.chain.insert(c.chain.begin() + 1, GGQ_F32);
        c.n_elems = p.nblocks * p.ti->elems;
        if ((rc = make_mem_io(j.dst, type_nbytes(c.chain.back(), c.n_elems), j.src, &c.io)) != GGQ_OK) return rc;
        cj.push_back(std::move(c));
    }
    return run_jobs_sharded(cj);
}

int ggq_slices_device(const struct ggq_slice_job *jobs, size_t n_jobs, void *stream) {
    if (n_jobs && !jobs) return fail(GGQ_ERR_INVALID, "null job table");
    std::vector<Plan> plans(n_jobs);
    for (size_t i = 0; i < n_jobs; i++) {  // validate everything before anything is enqueued
        const ggq_slice_job &j = jobs[i];
        int rc = j.quantize ? plan_quantize(j.type, j.fdt, j.dst_len, j.src_len, &plans[i]) : plan_dequantize(j.type, j.fdt, j.dst_len, j.src_len, &plans[i]);
        if (rc != GGQ_OK) return rc;
        if (plans[i].nblocks == 0) continue;
        if (!j.dst || !j.src) return fail(GGQ_ERR_INVALID, "null pointer with non-zero length");
        const void *fl = j.quantize ? j.src : j.dst, *pk = j.quantize ? (const void *)j.dst : j.src;
        if ((reinterpret_cast<uintptr_t>(fl) & (fdt_size(j.fdt) - 1)) || (reinterpret_cast<uintptr_t>(pk) & 1))
            return fail(GGQ_ERR_INVALID, "data is not aligned");
    }
    DevInfo dev;
    int rc = resolve_device(&dev);
    if (rc != GGQ_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Stream order = job order, so a later job may read what an earlier one wrote.  Runs of consecutive dequantize jobs
    // with one float side and a real block type share one grid (dequant.cu: dequant_blocks_batch); everything
    // else (quantize jobs, f16 / bf16 "blocks") is one launch per job.
    size_t i = 0;
    while (i < n_jobs) {
        const ggq_slice_job &j = jobs[i];
        const bool batchable = !j.quantize && !is_float_type(j.type);
        if (!batchable) {
            if (plans[i].nblocks) {
                const Plan &p = plans[i];
                const uint32_t from = j.quantize ? p.fdt : p.ti->type, to = j.quantize ? p.ti->type : p.fdt;
                // from == to (quantize::<f16, f16, 1>) is still mediated by f32 inside the cast kernel (lib.rs:66-73)
                cudaError_t e = enqueue_hop(from, to, j.dst, j.src, p.nblocks * p.ti->elems, st, dev);
                if (e != cudaSuccess) return fail_cuda(e, "ggq_slices_device");
            }
            i++;
            continue;
        }
        std::vector<DequantJob> run;
        const uint32_t fdt = j.fdt;
        // everything in one grid: keeping the big tensors on their dedicated kernels and batching only the small ones was
        // measured (6.32 against 6.75 TB/s for the bench step)
        cudaError_t e = cudaSuccess;
        for (; i < n_jobs && !jobs[i].quantize && !is_float_type(jobs[i].type) && jobs[i].fdt == fdt; i++)
            if (plans[i].nblocks) run.push_back({jobs[i].type, jobs[i].src, jobs[i].dst, plans[i].nblocks});
        if (run.empty()) continue;
        if (run.size() == 1) {
            g_launches.fetch_add(1, std::memory_order_relaxed);
            e = dequant_blocks(run[0].type, fdt, run[0].src, run[0].dst, run[0].nblocks, st, dev);
        } else {
            uint64_t launches = 0;
            e = dequant_blocks_batch(fdt, run.data(), run.size(), st, dev, &launches);
            g_launches.fetch_add(launches, std::memory_order_relaxed);
        }
        if (e != cudaSuccess) return fail_cuda(e, "ggq_slices_device");
    }
    return GGQ_OK;
}

size_t ggq_plan_shards(const struct ggq_slice_job *jobs, size_t n_jobs, int n_devices, struct ggq_shard_piece *out, size_t cap) {
    if (n_jobs && !jobs) return 0;
    std::vector<JobGeom> geom;
    std::vector<size_t> index;  // geom entry -> caller's job index (empty jobs carry no piece)
    for (size_t i = 0; i < n_jobs; i++) {
        const ggq_slice_job &j = jobs[i];
        Plan p;
        int rc = j.quantize ? plan_quantize(j.type, j.fdt, j.dst_len, j.src_len, &p) : plan_dequantize(j.type, j.fdt, j.dst_len, j.src_len, &p);
        if (rc != GGQ_OK) return 0;
        if (p.nblocks == 0) continue;
        geom.push_back({p.nblocks * p.ti->elems, j.quantize ? p.fdt : p.ti->type, j.quantize ? p.ti->type : p.fdt});
        index.push_back(i);
    }
    const std::vector<Piece> pieces = plan_pieces(geom, n_devices);
    for (size_t k = 0; k < pieces.size() && k < cap && out; k++)
        out[k] = {(uint32_t)index[pieces[k].job], pieces[k].device, pieces[k].e0, pieces[k].e1};
    return pieces.size();
}

int ggq_cast(const uint32_t *types, int n_types, void *dst, const void *src, size_t n_elems) {
    std::vector<uint32_t> chain;
    int rc = plan_cast(types, n_types, n_elems, &chain);
    return rc != GGQ_OK ? rc : run_chain_host(chain, dst, src, n_elems);
}

int ggq_rearrange_device(void *dst, const struct ggq_layout *dl, const void *src, const struct ggq_layout *sl, size_t unit, void *stream) {
    StridedLayout d, s;
    int rc = plan_rearrange(dl, sl, unit, &d, &s);  // validation first: no CUDA call is needed to reject a bad layout
    if (rc != GGQ_OK) return rc;
    if (!dst || !src) return fail(GGQ_ERR_INVALID, "null pointer");
    DevInfo dev;
    if ((rc = resolve_device(&dev)) != GGQ_OK) return rc;
    return rearrange_on_stream(dst, dl, src, sl, unit, static_cast<cudaStream_t>(stream));
}

int ggq_rearrange(void *dst, const struct ggq_layout *dl, const void *src, const struct ggq_layout *sl, size_t unit) {
    StridedLayout d, s;
    int rc = plan_rearrange(dl, sl, unit, &d, &s);
    if (rc != GGQ_OK) return rc;
    int64_t dlo, dhi, slo, shi;
    uint64_t dn, sn;
    if (!layout_span(*dl, unit, &dlo, &dhi, &dn) || !layout_span(*sl, unit, &slo, &shi, &sn)) return GGQ_OK;  // empty shape
    if (!dst || !src) return fail(GGQ_ERR_INVALID, "null pointer");
    if (pointer_kind(src) == cudaMemoryTypeDevice || pointer_kind(dst) == cudaMemoryTypeDevice)
        return fail(GGQ_ERR_INVALID, "device pointer passed to a host-pointer entry point (use ggq_rearrange_device)");
    const size_t dbytes = (size_t)(dhi - dlo), sbytes = (size_t)(shi - slo);
    const char *hs = static_cast<const char *>(src) + slo;
    char *hd = static_cast<char *>(dst) + dlo;
    ggq::Resident res;
    if (res.status() != GGQ_OK) return res.status();
    void *d_src = nullptr, *d_dst = nullptr;
    if ((rc = res.alloc(sbytes, &d_src)) != GGQ_OK || (rc = res.alloc(dbytes, &d_dst)) != GGQ_OK) return rc;
    if ((rc = res.upload(d_src, sbytes, [hs](void *pinned, size_t off, size_t n) { parallel_memcpy(pinned, hs + off, n); return true; })) != GGQ_OK) return rc;
    if (dn != dbytes)  // the layout leaves holes in its span: keep what the caller has there
        if ((rc = res.upload(d_dst, dbytes, [hd](void *pinned, size_t off, size_t n) { parallel_memcpy(pinned, hd + off, n); return true; })) != GGQ_OK) return rc;
    // device buffers start at the lowest addressed byte: shift the offsets
    ggq_layout dl2 = *dl, sl2 = *sl;
    dl2.offset -= dlo;
    sl2.offset -= slo;
    if ((rc = res.rearrange(d_dst, dl2, d_src, sl2, unit)) != GGQ_OK) return rc;
    return res.download(d_dst, dbytes, [hd](const void *pinned, size_t off, size_t n) { parallel_memcpy(hd + off, pinned, n); return true; });
}

size_t ggq_type_nbytes(uint32_t type, size_t n_elems) {
    size_t e, b;
    if (!type_geometry(type, &e, &b) || n_elems % e) return 0;
    return n_elems / e * b;
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

void ggq_shutdown(void) {
    std::vector<Pipeline *> idle;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        idle.swap(g_pool);
    }
    int cur = -1;
    cudaGetDevice(&cur);
    for (Pipeline *p : idle) {
        cudaSetDevice(p->device);
        destroy_pipeline(p);
    }
    quant_k_release_work();
    const int ndev = ggq_device_count();
    for (int d = 0; d < ndev && d < MAX_DEVICES; d++) {  // give cached stream-ordered allocations back
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    }
    if (cur >= 0) cudaSetDevice(cur);
    cudaGetLastError();
}

uint64_t ggq_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char *ggq_version(void) { return "ggq-b200 0.1 (sm_100a)"; }

}  // extern "C"
