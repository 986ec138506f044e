// rearrange.cu — strided byte-run copy on the device: what `Rearranging::launch(dst, src)` of the
// crate `mem-rearrange` 0.1.0 does on the host for the reference's block-granular tensor operators
// (/root/reference/xtask/src/utils/operator/merge.rs:311-313, 344-350 concat / split along an axis in
// block units; operator/permute_qk.rs:55-66 the rotary row interleave of attn_q / attn_k).  The crate
// is a crates.io dependency that is not vendored in the reference (Cargo.lock:314-317, 340-343); its
// published contract is restated here: for every index tuple of the common shape, copy `unit` bytes
// from  src + src.offset + Σ i_k·src.stride_k  to  dst + dst.offset + Σ i_k·dst.stride_k.
//
// Design: HBM-bound byte work, no arithmetic.  The host normalises the two layouts (drops extent-1
// dims, orders dims by destination stride, merges dims that are contiguous in BOTH layouts, folds
// the unit-stride dim into a contiguous "run") and launches one grid of 16-byte vector copies:
//   grid.x × threads enumerate (two innermost outer dims, vector),   grid.y / .z the next two dims
// so a thread pays two 32-bit divisions per 16 bytes and every warp writes consecutive vectors of a
// run.  Each thread keeps UNROLL independent loads in flight before its first store.  Vector width
// falls back to 8/4/2/1 bytes when a pointer, stride or run length is not 16-byte aligned (block
// rows are only 2-byte aligned in general: a Q4_0 row of 32·k elements is 18·k bytes).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "ggq_common.cuh"
#include "ggq_kernels.h"

namespace ggq {

struct RearrangeParams {
    uint32_t n0, n1;  // extents of the two innermost outer dims (decomposed per thread)
    uint32_t vecs;    // W-byte vectors per run
    int64_t ds[4], ss[4];  // byte strides of the four outer dims; dims 2 and 3 are blockIdx.y / .z
};

template <int W> struct RVec;
template <> struct RVec<16> { using type = uint4; };
template <> struct RVec<8> { using type = uint2; };
template <> struct RVec<4> { using type = uint32_t; };
template <> struct RVec<2> { using type = uint16_t; };
template <> struct RVec<1> { using type = uint8_t; };

constexpr int RE_THREADS = 256;

template <int W, int UNROLL>
__global__ void __launch_bounds__(RE_THREADS) rearrange_kernel(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, RearrangeParams p) {
    using V = typename RVec<W>::type;
    dst += (int64_t)blockIdx.y * p.ds[2] + (int64_t)blockIdx.z * p.ds[3];
    src += (int64_t)blockIdx.y * p.ss[2] + (int64_t)blockIdx.z * p.ss[3];
    const uint64_t total = (uint64_t)p.n0 * p.n1 * p.vecs;  // < 2^32 (host splits larger extents)
    const uint64_t base = (uint64_t)blockIdx.x * (RE_THREADS * UNROLL) + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    V v[UNROLL];
    int64_t doff[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
        const uint64_t idx = base + (uint64_t)u * RE_THREADS;
        if (idx < total) {
            const uint32_t r = (uint32_t)idx / p.vecs, k = (uint32_t)idx - r * p.vecs;
            const uint32_t i1 = r / p.n0, i0 = r - i1 * p.n0;
            v[u] = __ldg(reinterpret_cast<const V *>(src + (int64_t)i0 * p.ss[0] + (int64_t)i1 * p.ss[1] + (int64_t)k * W));
            doff[u] = (int64_t)i0 * p.ds[0] + (int64_t)i1 * p.ds[1] + (int64_t)k * W;
        }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++)
        if (base + (uint64_t)u * RE_THREADS < total) *reinterpret_cast<V *>(dst + doff[u]) = v[u];
}

namespace {

struct Dim { uint64_t n; int64_t ds, ss; };

template <int W>
cudaError_t launch_w(uint8_t *dst, const uint8_t *src, const RearrangeParams &p, unsigned gy, unsigned gz, cudaStream_t stream) {
    const uint64_t total = (uint64_t)p.n0 * p.n1 * p.vecs;
    // four loads in flight per thread once that still leaves every SM several CTAs; small tensors keep
    // one vector per thread so the grid stays wide (they are latency-, not bandwidth-bound)
    const uint64_t ctas4 = (total + RE_THREADS * 4 - 1) / (RE_THREADS * 4) * gy * gz;
    if (ctas4 >= 148 * 8) {
        dim3 grid((unsigned)((total + RE_THREADS * 4 - 1) / (RE_THREADS * 4)), gy, gz);
        return launch_pdl(rearrange_kernel<W, 4>, grid, RE_THREADS, 0, stream, dst, src, p);
    }
    dim3 grid((unsigned)((total + RE_THREADS - 1) / RE_THREADS), gy, gz);
    return launch_pdl(rearrange_kernel<W, 1>, grid, RE_THREADS, 0, stream, dst, src, p);
}

// dims: innermost first.  The grid enumerates four of them (two per thread index, two per blockIdx.y/z);
// anything beyond that, or beyond the grid limits, is looped over on the host.
cudaError_t launch_dims(uint8_t *dst, const uint8_t *src, std::vector<Dim> dims, uint64_t run, int w, cudaStream_t stream, uint64_t *launches) {
    constexpr uint64_t MAX_YZ = 65535, MAX_FLAT = 0xFFFFFFFFull;
    const uint64_t vecs = run / (uint64_t)w;
    // loop over dim `d` on the host in chunks of `per`
    auto host_loop = [&](size_t d, uint64_t per) -> cudaError_t {
        for (uint64_t c = 0; c < dims[d].n; c += per) {
            std::vector<Dim> d2 = dims;
            d2[d].n = std::min(per, dims[d].n - c);
            if (per == 1) d2.erase(d2.begin() + d);
            cudaError_t e = launch_dims(dst + (int64_t)c * dims[d].ds, src + (int64_t)c * dims[d].ss, d2, run, w, stream, launches);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    if (dims.size() > 4) return host_loop(dims.size() - 1, 1);
    while (dims.size() < 4) dims.push_back({1, 0, 0});
    if (dims[3].n > MAX_YZ) return host_loop(3, MAX_YZ);
    if (dims[2].n > MAX_YZ) return host_loop(2, MAX_YZ);
    if (dims[0].n * vecs > MAX_FLAT) return host_loop(0, std::max<uint64_t>(1, MAX_FLAT / vecs));  // vecs <= 2^26 (caller)
    if (dims[0].n * dims[1].n * vecs > MAX_FLAT) return host_loop(1, std::max<uint64_t>(1, MAX_FLAT / (dims[0].n * vecs)));
    RearrangeParams p;
    p.n0 = (uint32_t)dims[0].n;
    p.n1 = (uint32_t)dims[1].n;
    p.vecs = (uint32_t)vecs;
    for (int i = 0; i < 4; i++) { p.ds[i] = dims[i].ds; p.ss[i] = dims[i].ss; }
    const unsigned gy = (unsigned)dims[2].n, gz = (unsigned)dims[3].n;
    ++*launches;
    switch (w) {
        case 16: return launch_w<16>(dst, src, p, gy, gz, stream);
        case 8: return launch_w<8>(dst, src, p, gy, gz, stream);
        case 4: return launch_w<4>(dst, src, p, gy, gz, stream);
        case 2: return launch_w<2>(dst, src, p, gy, gz, stream);
    }
    return launch_w<1>(dst, src, p, gy, gz, stream);
}

}  // namespace

cudaError_t rearrange_strided(void *dst_base, const StridedLayout &dl, const void *src_base, const StridedLayout &sl, size_t unit,
                              cudaStream_t stream, uint64_t *launches) {
    *launches = 0;
    std::vector<Dim> dims;
    for (int i = 0; i < dl.ndim; i++) {
        if (dl.shape[i] == 0) return cudaSuccess;  // empty tensor
        if (dl.shape[i] > 1) dims.push_back({dl.shape[i], dl.strides[i], sl.strides[i]});
    }
    uint8_t *dst = static_cast<uint8_t *>(dst_base) + dl.offset;
    const uint8_t *src = static_cast<const uint8_t *>(src_base) + sl.offset;
    // innermost = smallest destination stride: consecutive threads then write consecutive bytes
    std::stable_sort(dims.begin(), dims.end(), [](const Dim &a, const Dim &b) { return std::llabs(a.ds) < std::llabs(b.ds); });
    // merge neighbours that are contiguous in both layouts
    for (size_t i = 0; i + 1 < dims.size();) {
        if (dims[i + 1].ds == dims[i].ds * (int64_t)dims[i].n && dims[i + 1].ss == dims[i].ss * (int64_t)dims[i].n) {
            dims[i].n *= dims[i + 1].n;
            dims.erase(dims.begin() + i + 1);
        } else {
            i++;
        }
    }
    uint64_t run = unit;
    if (!dims.empty() && dims[0].ds == (int64_t)unit && dims[0].ss == (int64_t)unit) {
        run = unit * dims[0].n;
        dims.erase(dims.begin());
    }
    // a run longer than the 32-bit vector index allows becomes an extra dim of 1 GiB pieces (+ a tail call)
    constexpr uint64_t PIECE = uint64_t(1) << 30;
    if (run > PIECE) {
        const uint64_t whole = run / PIECE, tail = run % PIECE;
        if (tail) {
            std::vector<Dim> td = dims;  // tail piece of every run: same outer dims, shorter run
            uint64_t l2 = 0;
            int w = 16;
            auto misaligned = [&](int ww) {
                if (tail % ww || (reinterpret_cast<uintptr_t>(dst + whole * PIECE) % ww) || (reinterpret_cast<uintptr_t>(src + whole * PIECE) % ww)) return true;
                for (const Dim &d : td) if (d.ds % ww || d.ss % ww) return true;
                return false;
            };
            while (w > 1 && misaligned(w)) w >>= 1;
            cudaError_t e = launch_dims(dst + whole * PIECE, src + whole * PIECE, td, tail, w, stream, &l2);
            if (e != cudaSuccess) return e;
            *launches += l2;
        }
        dims.insert(dims.begin(), Dim{whole, (int64_t)PIECE, (int64_t)PIECE});
        run = PIECE;
    }
    int w = 16;
    auto misaligned = [&](int ww) {
        if (run % ww || (reinterpret_cast<uintptr_t>(dst) % ww) || (reinterpret_cast<uintptr_t>(src) % ww)) return true;
        for (const Dim &d : dims) if (d.ds % ww || d.ss % ww) return true;
        return false;
    };
    while (w > 1 && misaligned(w)) w >>= 1;
    uint64_t l = 0;
    cudaError_t e = launch_dims(dst, src, dims, run, w, stream, &l);
    *launches += l;
    return e;
}

}  // namespace ggq
