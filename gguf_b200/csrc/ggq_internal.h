// ggq_internal.h — C++ plumbing shared inside libggq.so (not part of the C ABI in include/ggq.h).
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>

#include "../../include/ggq.h"

namespace ggq {

// Where a cast chain's bytes come from and go to.  Offsets are byte offsets inside the tensor.
struct ChainIO {
    const void *direct_src = nullptr;  // pinned host memory holding the whole input (DMA source), or null
    void *direct_dst = nullptr;        // pinned host memory for the whole output, or null
    std::function<bool(void *pinned, size_t byte_off, size_t nbytes)> read;         // fill a pinned bounce buffer
    std::function<bool(const void *pinned, size_t byte_off, size_t nbytes)> write;  // drain one
    // Optional replacement for `read` (O_DIRECT sources): may place the bytes anywhere in [pinned, pinned + cap) and returns
    // the offset they start at (< 0: failure), so a 4 KiB-aligned read can cover an unaligned tensor range without a copy.
    std::function<long(void *pinned, size_t cap, size_t byte_off, size_t nbytes)> read_shift;
};

// cast.rs:93-138 over a chain of types, streaming through the H2D -> kernels -> D2H pipeline of the
// calling thread's device.  Returns a ggq_status; ggq_last_error() has the message.
int cast_chain_io(const uint32_t *types, int n_types, size_t n_elems, const ChainIO &io);

// Device-resident evaluation for tensors whose bytes are not a plain stream (rearranged rows): one
// worker thread's context on its GPU — a stream, stream-ordered allocations, and the pinned bounce
// buffers of a pooled pipeline for chunked upload / download.  Everything is enqueued on one stream
// in call order; download() returns after the last byte has been handed to `write`.
// memcpy with non-temporal stores (host_copy.cpp): for the bounce copies to / from pinned staging buffers.
void stream_copy(void *dst, const void *src, size_t n);

// Per-thread accounting of the host pipelines (run_jobs_io, Resident::upload / download): nanoseconds the calling
// thread spent blocked on the GPU and the bytes it moved over PCIe since the last take.  For ggq_convert_stats.
struct PipeCounters { uint64_t gpu_wait_ns = 0, h2d_bytes = 0, d2h_bytes = 0; };
PipeCounters take_pipe_counters();

// Free device memory (bytes) on the calling thread's device, 0 when it cannot be queried.
size_t device_free_bytes();

using ReadFn = std::function<bool(void *pinned, size_t byte_off, size_t nbytes)>;
using WriteFn = std::function<bool(const void *pinned, size_t byte_off, size_t nbytes)>;
class Resident {
   public:
    Resident();   // binds to the calling thread's device (ggq_set_device); check status()
    ~Resident();  // synchronises the stream and frees what alloc() handed out
    Resident(const Resident &) = delete;
    Resident &operator=(const Resident &) = delete;
    int status() const { return status_; }
    int alloc(size_t nbytes, void **d);
    void free(void *d);
    int upload(void *d_dst, size_t nbytes, const ReadFn &read);
    int download(const void *d_src, size_t nbytes, const WriteFn &write);
    int cast(const uint32_t *types, int n_types, size_t n_elems, void *d_dst, const void *d_src);
    int rearrange(void *d_dst, const ggq_layout &dl, const void *d_src, const ggq_layout &sl, size_t unit);

   private:
    struct Impl;
    Impl *impl_ = nullptr;
    int status_ = 0;
};

}  // namespace ggq
