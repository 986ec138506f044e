// array_layout.hpp — the slice of the crate `ndarray-layout` 0.2.1 (`ArrayLayout<N>`) the reference's
// tensor operators use, restated for the host side of libggq (the crate is a crates.io dependency of
// the reference, Cargo.lock:340-343, not vendored).  Call sites mirrored:
//   new_contiguous(shape, LittleEndian, unit)   xtask/src/utils/operator/merge.rs:359-364, permute_qk.rs:57,60
//   tile_le(axis, tiles)                        permute_qk.rs:58
//   transpose(perm)                             permute_qk.rs:59
//   split(axis, parts)                          merge.rs:310, 337
// Shapes count elements of `unit` bytes; strides and offset are bytes.  "Little endian" = the first
// dim is the fastest varying one, which is ggml's ne[0]-first order.
#pragma once

#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "../../include/ggq.h"

namespace ndl {

struct ArrayLayout {
    std::vector<uint64_t> shape;
    std::vector<int64_t> strides;
    int64_t offset = 0;

    // ArrayLayout::new_contiguous(shape, Endian::LittleEndian, element_size)
    static ArrayLayout new_contiguous(const std::vector<uint64_t> &shape, uint64_t unit) {
        ArrayLayout l;
        l.shape = shape;
        int64_t mul = (int64_t)unit;
        for (uint64_t d : shape) {
            l.strides.push_back(mul);
            mul *= (int64_t)d;
        }
        return l;
    }

    // tile_le: one dim becomes several, the first tile varying fastest; product(tiles) == shape[axis]
    ArrayLayout tile_le(size_t axis, const std::vector<uint64_t> &tiles) const {
        if (axis >= shape.size()) throw std::invalid_argument("tile_le: axis out of range");
        uint64_t prod = 1;
        for (uint64_t t : tiles) prod *= t;
        if (prod != shape[axis]) throw std::invalid_argument("tile_le: tiles do not multiply to the dim");
        ArrayLayout l;
        l.offset = offset;
        for (size_t i = 0; i < shape.size(); i++) {
            if (i != axis) {
                l.shape.push_back(shape[i]);
                l.strides.push_back(strides[i]);
                continue;
            }
            int64_t st = strides[axis];
            for (uint64_t t : tiles) {
                l.shape.push_back(t);
                l.strides.push_back(st);
                st *= (int64_t)t;
            }
        }
        return l;
    }

    // transpose(perm): the dims listed in `perm` are placed, in that order, at the sorted positions of
    // `perm`; every other dim stays where it is.  transpose({2, 1}) swaps dims 1 and 2.
    ArrayLayout transpose(const std::vector<size_t> &perm) const {
        std::vector<size_t> pos = perm;
        std::sort(pos.begin(), pos.end());
        if (std::adjacent_find(pos.begin(), pos.end()) != pos.end()) throw std::invalid_argument("transpose: repeated dim");
        if (!pos.empty() && pos.back() >= shape.size()) throw std::invalid_argument("transpose: dim out of range");
        ArrayLayout l = *this;
        for (size_t k = 0; k < perm.size(); k++) {
            l.shape[pos[k]] = shape[perm[k]];
            l.strides[pos[k]] = strides[perm[k]];
        }
        return l;
    }

    // split(axis, parts): consecutive sub-ranges of one dim, each with its own offset
    std::vector<ArrayLayout> split(size_t axis, const std::vector<uint64_t> &parts) const {
        if (axis >= shape.size()) throw std::invalid_argument("split: axis out of range");
        uint64_t sum = 0;
        for (uint64_t p : parts) sum += p;
        if (sum != shape[axis]) throw std::invalid_argument("split: parts do not add up to the dim");
        std::vector<ArrayLayout> out;
        uint64_t start = 0;
        for (uint64_t p : parts) {
            ArrayLayout l = *this;
            l.shape[axis] = p;
            l.offset = offset + (int64_t)start * strides[axis];
            out.push_back(std::move(l));
            start += p;
        }
        return out;
    }

    uint64_t count() const { uint64_t n = 1; for (uint64_t d : shape) n *= d; return n; }

    // true when the layout addresses one gap-free ascending byte range (so a part of a concat can be
    // written in place, without a rearranging copy)
    bool is_dense(uint64_t unit) const {
        int64_t expect = (int64_t)unit;
        for (size_t i = 0; i < shape.size(); i++) {
            if (shape[i] == 1) continue;
            if (strides[i] != expect) return false;
            expect *= (int64_t)shape[i];
        }
        return true;
    }

    ggq_layout c() const {
        if (shape.size() > GGQ_MAX_NDIM) throw std::invalid_argument("layout has more than GGQ_MAX_NDIM dims");
        ggq_layout l{};
        l.ndim = (uint32_t)shape.size();
        for (size_t i = 0; i < shape.size(); i++) {
            l.shape[i] = shape[i];
            l.strides[i] = strides[i];
        }
        l.offset = offset;
        return l;
    }
};

}  // namespace ndl
