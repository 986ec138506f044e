// host_copy.cpp — the bounce copies between pageable caller memory and the pinned staging buffers.
// They are pure streaming copies whose destination is not read again by this core (the DMA engine or
// the caller consumes it), so non-temporal stores skip the read-for-ownership of every destination
// line: a third less DRAM traffic than a cached memcpy of a chunk that is too small for glibc's own
// non-temporal threshold.  AVX2 when the CPU has it, memcpy otherwise.
#include <cstdint>
#include <cstring>
#include <immintrin.h>

#include "ggq_internal.h"

namespace ggq {

__attribute__((target("avx2"))) static void stream_copy_avx2(char *d, const char *s, size_t n) {
    // head: bring the destination to 32-byte alignment
    const size_t head = (32 - (reinterpret_cast<uintptr_t>(d) & 31)) & 31;
    if (head) {
        const size_t h = head < n ? head : n;
        memcpy(d, s, h);
        d += h; s += h; n -= h;
    }
    while (n >= 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + 64));
        const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + 96), e);
        d += 128; s += 128; n -= 128;
    }
    if (n) memcpy(d, s, n);
    _mm_sfence();
}

void stream_copy(void *dst, const void *src, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && n >= 4096) stream_copy_avx2(static_cast<char *>(dst), static_cast<const char *>(src), n);
    else memcpy(dst, src, n);
}

}  // namespace ggq
