// ggq_internal.h — C++ plumbing shared inside libggq.so (not part of the C ABI in include/ggq.h).
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>

namespace ggq {

// Where a cast chain's bytes come from and go to.  Offsets are byte offsets inside the tensor.
struct ChainIO {
    const void *direct_src = nullptr;  // pinned host memory holding the whole input (DMA source), or null
    void *direct_dst = nullptr;        // pinned host memory for the whole output, or null
    std::function<bool(void *pinned, size_t byte_off, size_t nbytes)> read;         // fill a pinned bounce buffer
    std::function<bool(const void *pinned, size_t byte_off, size_t nbytes)> write;  // drain one
};

// cast.rs:93-138 over a chain of types, streaming through the H2D -> kernels -> D2H pipeline of the
// calling thread's device.  Returns a ggq_status; ggq_last_error() has the message.
int cast_chain_io(const uint32_t *types, int n_types, size_t n_elems, const ChainIO &io);

}  // namespace ggq
