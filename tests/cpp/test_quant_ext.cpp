// test_quant_ext.cpp — the reference's own unit tests, restated against the C++ host mirror
// (gguf_b200/host/ggml_quants.hpp) so they read like ggml-quants/src/lib.rs:264-370
// (`generate_tests!(Q8_0, 4.5e-3)`) and the per-struct round-trip tests (structs/*.rs, thresholds in
// SURVEY.md App. C.3).  `--no-gpu` runs only the checks that need no device (error order, layouts).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../../gguf_b200/host/ggml_quants.hpp"

using namespace ggml_quants;
static int failures = 0;
#define CHECK(...) do { if (!(__VA_ARGS__)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #__VA_ARGS__); failures++; } } while (0)

// lib.rs:293-331
static void test_error_order() {
    std::vector<float> src31(31), src64(64);
    std::vector<Q8_0> one(1), three(3);
    CHECK(QuantExt<Q8_0, float>::quantize_slice(one.data(), 1, src31.data(), 31).unwrap_err() == QuantizeError::Indivisible);
    CHECK(QuantExt<Q8_0, float>::quantize_slice(three.data(), 3, src64.data(), 64).unwrap_err() == QuantizeError::LengthMismatch);
    CHECK(QuantExt<Q8_0, float>::dequantize_slice(src31.data(), 31, one.data(), 1).unwrap_err() == QuantizeError::Indivisible);
    CHECK(QuantExt<Q8_0, float>::dequantize_slice(src64.data(), 64, three.data(), 3).unwrap_err() == QuantizeError::LengthMismatch);
    CHECK(Q8_0::COUNT == 32 && Q8K::COUNT == 256 && Q4K::COUNT == 256);
}

// lib.rs:277-291: 64 random values in [-1,1), slice round trip within 4.5e-3
static void test_q8_0_slice_roundtrip(std::mt19937 &rng) {
    std::uniform_real_distribution<float> u(-1.f, 1.f);
    std::vector<float> x(64), y(64);
    for (auto &v : x) v = u(rng);
    std::vector<Q8_0> q(2);
    QuantExt<Q8_0, float>::quantize_slice(q.data(), 2, x.data(), 64).unwrap();
    QuantExt<Q8_0, float>::dequantize_slice(y.data(), 64, q.data(), 2).unwrap();
    for (int i = 0; i < 64; i++) CHECK(std::fabs(x[i] - y[i]) <= 4.5e-3f);
}

// lib.rs:333-338: the zero block dequantises to zeros
static void test_zero_block() {
    Q8_0 z = Q8_0::ZEROS();
    auto y = Quantize<Q8_0, float>::dequantize(z);
    for (float v : y) CHECK(v == 0.f);
    std::array<float, 32> zeros{};
    Q4_0 b = Quantize<Q4_0, float>::quantize(zeros);
    Q4_0 zz = Q4_0::ZEROS();
    CHECK(std::memcmp(&b, &zz, sizeof b) == 0);  // reference returns Self::ZEROS (q4_0.rs:30-33)
}

// structs/*.rs `test_utils::test::<N, T>(abs, 0.)`: one random [0,1) block, quantize -> dequantize
template <class Blk> static void test_block_tolerance(std::mt19937 &rng, float tol, const char *name) {
    constexpr size_t N = Blk::COUNT;
    std::uniform_real_distribution<float> u(0.f, 1.f);
    std::array<float, N> x;
    for (auto &v : x) v = u(rng);
    Blk q = Quantize<Blk, float>::quantize(x);
    auto y = Quantize<Blk, float>::dequantize(q);
    float worst = 0;
    for (size_t i = 0; i < N; i++) worst = std::fmax(worst, std::fabs(x[i] - y[i]));
    if (worst > tol) { std::printf("FAIL %s: max abs err %g > %g\n", name, worst, tol); failures++; }
}

// the multi-GPU split as the library reports it (pure host function): covers every element once, cuts on 2^20 elements
static void test_plan_shards() {
    std::vector<ggq_slice_job> jobs;
    const size_t n0 = size_t(4096) * 14336, n1 = size_t(4096) * 4096;
    jobs.push_back(dequantize_job<Q4_0, f16>(nullptr, n0, nullptr, n0 / 32));
    jobs.push_back(quantize_job<Q8_0, f16>(nullptr, n1 / 32, nullptr, n1));
    for (int ndev : {1, 2, 8}) {
        std::vector<ggq_shard_piece> p(64);
        const size_t n = plan_shards(jobs.data(), jobs.size(), ndev, p.data(), p.size());
        CHECK(n >= (size_t)ndev && n <= p.size());
        size_t covered[2] = {0, 0};
        for (size_t i = 0; i < n; i++) {
            CHECK(p[i].device >= 0 && p[i].device < ndev && p[i].elem_begin % (size_t(1) << 20) == 0);
            covered[p[i].job] += p[i].elem_end - p[i].elem_begin;
        }
        CHECK(covered[0] == n0 && covered[1] == n1);
    }
    ggq_slice_job bad = quantize_job<Q8_0, float>(nullptr, 3, nullptr, 33);
    CHECK(plan_shards(&bad, 1, 2, nullptr, 0) == 0);
    CHECK(slices(&bad, 1).unwrap_err() == QuantizeError::Indivisible);
}

int main(int argc, char **argv) {
    const bool no_gpu = argc > 1 && std::strcmp(argv[1], "--no-gpu") == 0;
    test_error_order();
    test_plan_shards();
    if (!no_gpu) {
        if (ggq_device_count() < 1) { std::printf("no CUDA device\n"); return 2; }
        std::mt19937 rng(1234);
        test_q8_0_slice_roundtrip(rng);
        test_zero_block();
        test_block_tolerance<Q4_0>(rng, 8e-2f, "Q4_0");
        test_block_tolerance<Q4_1>(rng, 4e-2f, "Q4_1");
        test_block_tolerance<Q5_0>(rng, 4e-2f, "Q5_0");
        test_block_tolerance<Q5_1>(rng, 2e-2f, "Q5_1");
        test_block_tolerance<Q8_0>(rng, 4.5e-3f, "Q8_0");
        test_block_tolerance<Q8_1>(rng, 4.5e-3f, "Q8_1");
        test_block_tolerance<Q8K>(rng, 4.5e-3f, "Q8K");
        // the K-quants the reference leaves as todo!(): thresholds scaled from their bit widths
        test_block_tolerance<Q6K>(rng, 2e-2f, "Q6K");
        test_block_tolerance<Q5K>(rng, 4e-2f, "Q5K");
        test_block_tolerance<Q4K>(rng, 8e-2f, "Q4K");
        test_block_tolerance<Q3K>(rng, 2e-1f, "Q3K");
        test_block_tolerance<Q2K>(rng, 4e-1f, "Q2K");
        // f16 adapter (lib.rs:340-351): f16 in, f16 out through the same block type
        std::vector<f16> h(64), hb(64);
        std::vector<float> x(64);
        std::uniform_real_distribution<float> u(-1.f, 1.f);
        for (auto &v : x) v = u(rng);
        QuantExt<f16, float>::quantize_slice(h.data(), 64, x.data(), 64).unwrap();   // f32 -> f16 cast
        std::vector<Q8_0> q(2);
        QuantExt<Q8_0, f16>::quantize_slice(q.data(), 2, h.data(), 64).unwrap();
        QuantExt<Q8_0, f16>::dequantize_slice(hb.data(), 64, q.data(), 2).unwrap();
        std::vector<float> y(64);
        QuantExt<f16, float>::dequantize_slice(y.data(), 64, hb.data(), 64).unwrap(); // f16 -> f32 cast
        for (int i = 0; i < 64; i++) CHECK(std::fabs(x[i] - y[i]) <= 4.5e-3f);
    }
    if (!no_gpu) {
        // cast.rs:132-135 leaves Q8_0 -> F16 as todo!(); the library implements every pair
        std::vector<float> x(64, 0.25f), y(64);
        std::vector<Q8_0> q(2);
        QuantExt<Q8_0, float>::quantize_slice(q.data(), 2, x.data(), 64).unwrap();
        std::vector<f16> h(64);
        cast({GGQ_Q8_0, GGQ_F16}, h.data(), q.data(), 64).unwrap();
        cast({GGQ_F16, GGQ_F32}, y.data(), h.data(), 64).unwrap();
        for (int i = 0; i < 64; i++) CHECK(std::fabs(y[i] - 0.25f) <= 4.5e-3f);
        CHECK(cast({GGQ_Q8_0, GGQ_F16}, h.data(), q.data(), 31).unwrap_err() == QuantizeError::Indivisible);
        CHECK(type_nbytes(GGQ_Q4K, 512) == 288 && type_nbytes(GGQ_Q4K, 100) == 0);
    }
    {
        // layout algebra of permute_qk.rs:57-60 (documented ndarray-layout examples) — no GPU needed
        using ndl::ArrayLayout;
        const ArrayLayout t = ArrayLayout::new_contiguous({6, 3, 2}, 1);      // strides 1, 6, 18
        const ArrayLayout tl = t.tile_le(0, {2, 3});
        CHECK((tl.shape == std::vector<uint64_t>{2, 3, 3, 2}) && (tl.strides == std::vector<int64_t>{1, 2, 6, 18}));
        const ArrayLayout tr = t.transpose({1, 0});
        CHECK((tr.shape == std::vector<uint64_t>{3, 6, 2}) && (tr.strides == std::vector<int64_t>{6, 1, 18}));
        const auto parts = ArrayLayout::new_contiguous({4, 6}, 2).split(1, {1, 2, 3});
        CHECK(parts.size() == 3 && parts[1].offset == 8 && parts[2].offset == 24 && parts[2].shape[1] == 3);
        CHECK(parts[0].is_dense(2) && ArrayLayout::new_contiguous({4, 6, 2}, 2).split(1, {2, 4})[0].is_dense(2) == false);
        // shape mismatch = SchemeError::ShapeMismatch, rejected before any CUDA call
        uint8_t buf[64] = {0};
        CHECK(rearrange(buf, ArrayLayout::new_contiguous({4, 4}, 2), buf, ArrayLayout::new_contiguous({4, 3}, 2), 2).unwrap_err() == QuantizeError::LengthMismatch);
    }
    if (!no_gpu) {
        // permute_qk.rs:46-69 on 8 rows of 4 bytes, 2 heads: within a head, row 2*i+j <- row j*half+i
        using ndl::ArrayLayout;
        uint8_t src[32], dst[32];
        for (int i = 0; i < 32; i++) src[i] = (uint8_t)i;
        const ArrayLayout sl = ArrayLayout::new_contiguous({4, 8}, 1).tile_le(1, {2, 2, 2}).transpose({2, 1});
        const ArrayLayout dl = ArrayLayout::new_contiguous(sl.shape, 1);
        rearrange(dst, dl, src, sl, 1).unwrap();
        const int want_row[8] = {0, 2, 1, 3, 4, 6, 5, 7};
        for (int r = 0; r < 8; r++)
            for (int b = 0; b < 4; b++) CHECK(dst[r * 4 + b] == want_row[r] * 4 + b);
    }
    std::printf(failures ? "%d FAILURES\n" : "all ok%.0d\n", failures);
    return failures ? 1 : 0;
}
