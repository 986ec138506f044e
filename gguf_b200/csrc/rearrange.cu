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
//   grid.y  enumerates the two outer dims,   grid.x × threads enumerate (innermost outer dim, vector)
// so a thread pays one 32-bit division per 16 bytes and every warp writes consecutive vectors of a
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
    uint32_t n0;    // extent of the innermost outer dim (decomposed per thread)
    uint32_t n1;    // blockIdx.y = i1 + n1 * i2
    uint32_t vecs;  // W-byte vectors per run
    int64_t ds0, ds1, ds2, ss0, ss1, ss2;  // byte strides of the three outer dims
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
    const uint32_t i2 = blockIdx.y / p.n1, i1 = blockIdx.y - i2 * p.n1;
    dst += (int64_t)i1 * p.ds1 + (int64_t)i2 * p.ds2;
    src += (int64_t)i1 * p.ss1 + (int64_t)i2 * p.ss2;
    const uint64_t total = (uint64_t)p.n0 * p.vecs;  // < 2^32 (host splits larger extents)
    const uint64_t base = (uint64_t)blockIdx.x * (RE_THREADS * UNROLL) + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    V v[UNROLL];
    int64_t doff[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
        const uint64_t idx = base + (uint64_t)u * RE_THREADS;
        if (idx < total) {
            const uint32_t i0 = (uint32_t)idx / p.vecs, k = (uint32_t)idx - i0 * p.vecs;
            v[u] = __ldg(reinterpret_cast<const V *>(src + (int64_t)i0 * p.ss0 + (int64_t)k * W));
            doff[u] = (int64_t)i0 * p.ds0 + (int64_t)k * W;
        }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++)
        if (base + (uint64_t)u * RE_THREADS < total) *reinterpret_cast<V *>(dst + doff[u]) = v[u];
}

namespace {

struct Dim { uint64_t n; int64_t ds, ss; };

template <int W>
cudaError_t launch_w(uint8_t *dst, const uint8_t *src, const RearrangeParams &p, uint32_t n2, cudaStream_t stream) {
    const uint64_t total = (uint64_t)p.n0 * p.vecs;
    const unsigned gy = p.n1 * n2;
    if (total >= (uint64_t)RE_THREADS * 4) {
        dim3 grid((unsigned)((total + RE_THREADS * 4 - 1) / (RE_THREADS * 4)), gy);
        return launch_pdl(rearrange_kernel<W, 4>, grid, RE_THREADS, 0, stream, dst, src, p);
    }
    dim3 grid((unsigned)((total + RE_THREADS - 1) / RE_THREADS), gy);
    return launch_pdl(rearrange_kernel<W, 1>, grid, RE_THREADS, 0, stream, dst, src, p);
}

// dims: innermost first, at most 3 after the recursion below has peeled the rest
cudaError_t launch_dims(uint8_t *dst, const uint8_t *src, std::vector<Dim> dims, uint64_t run, int w, cudaStream_t stream, uint64_t *launches) {
    constexpr uint64_t MAX_Y = 65535, MAX_FLAT = 0xFFFFFFFFull;
    const uint64_t vecs = run / (uint64_t)w;
    // peel dims the grid cannot enumerate: host loop over the outermost one
    auto peel_last = [&]() -> cudaError_t {
        const Dim d = dims.back();
        dims.pop_back();
        for (uint64_t i = 0; i < d.n; i++) {
            cudaError_t e = launch_dims(dst + (int64_t)i * d.ds, src + (int64_t)i * d.ss, dims, run, w, stream, launches);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    if (dims.size() > 3) return peel_last();
    while (dims.size() < 3) dims.push_back({1, 0, 0});
    if (dims[1].n * dims[2].n > MAX_Y) {
        if (dims[2].n > 1) return peel_last();
        // one huge middle dim: chunks of MAX_Y
        for (uint64_t c = 0; c < dims[1].n; c += MAX_Y) {
            std::vector<Dim> d2 = dims;
            d2[1].n = std::min(MAX_Y, dims[1].n - c);
            cudaError_t e = launch_dims(dst + (int64_t)c * dims[1].ds, src + (int64_t)c * dims[1].ss, d2, run, w, stream, launches);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    if (dims[0].n * vecs > MAX_FLAT) {
        const uint64_t per = std::max<uint64_t>(1, MAX_FLAT / vecs);  // vecs <= MAX_FLAT is ensured by the caller
        for (uint64_t c = 0; c < dims[0].n; c += per) {
            std::vector<Dim> d2 = dims;
            d2[0].n = std::min(per, dims[0].n - c);
            cudaError_t e = launch_dims(dst + (int64_t)c * dims[0].ds, src + (int64_t)c * dims[0].ss, d2, run, w, stream, launches);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    RearrangeParams p;
    p.n0 = (uint32_t)dims[0].n;
    p.n1 = (uint32_t)dims[1].n;
    p.vecs = (uint32_t)vecs;
    p.ds0 = dims[0].ds; p.ss0 = dims[0].ss;
    p.ds1 = dims[1].ds; p.ss1 = dims[1].ss;
    p.ds2 = dims[2].ds; p.ss2 = dims[2].ss;
    const uint32_t n2 = (uint32_t)dims[2].n;
    ++*launches;
    switch (w) {
        case 16: return launch_w<16>(dst, src, p, n2, stream);
        case 8: return launch_w<8>(dst, src, p, n2, stream);
        case 4: return launch_w<4>(dst, src, p, n2, stream);
        case 2: return launch_w<2>(dst, src, p, n2, stream);
    }
    return launch_w<1>(dst, src, p, n2, stream);
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
