// tensor_ops.hpp — host-side planning of the reference's tensor operators, header only, no CUDA and no
// I/O: what `Content::apply` (xtask/src/utils/operator/mod.rs:59-72) does to the tensor list for
//   cast:<rules>                    operator/cast.rs:11-90     (rule grammar, per-architecture classes)
//   merge-linear / split-linear     operator/merge.rs:22-357   (grouping, naming, order, concat / split)
//   permute-qk                      operator/permute_qk.rs:11-69
// Like the reference's `DataPromise::lazy` (utils/mod.rs:104-138) an operator does not move data: each
// tensor is an expression (`Node`) over byte ranges of the input files that convert.cpp evaluates when
// it writes the tensor.  Expressions are normalised while they are built — casts and splits commute
// with whole-row moves (a block never straddles a row, ggus/src/tensor.rs:92) — so that only real row
// permutations and 3-D expert merges are left for the device-resident path.
// Where the reference panics (assert / unwrap / todo!) these functions throw StepError.
#pragma once

#include <cctype>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <map>
#include <memory>
#include <string>
#include <string_view>
#include <vector>

#include "../../include/ggq.h"
#include "array_layout.hpp"
#include "gguf.hpp"

namespace tensor_ops {

inline bool ends_with(std::string_view s, std::string_view suf) { return s.size() >= suf.size() && s.substr(s.size() - suf.size()) == suf; }

inline std::string upper(std::string s) { for (char &c : s) c = (char)std::toupper((unsigned char)c); return s; }

// cast.rs:179-216 `parse`
inline bool parse_type(const std::string &name, uint32_t *ty) {
    static const std::map<std::string, uint32_t> M = {
        {"F32", 0}, {"F16", 1}, {"Q4_0", 2}, {"Q4_1", 3}, {"Q5_0", 6}, {"Q5_1", 7}, {"Q8_0", 8}, {"Q8_1", 9}, {"Q2K", 10},
        {"Q3K", 11}, {"Q4K", 12}, {"Q5K", 13}, {"Q6K", 14}, {"Q8K", 15}, {"BF16", 30}};
    auto it = M.find(upper(name));
    if (it == M.end()) return false;
    *ty = it->second;
    return true;
}

struct CastRule { bool has[4] = {false, false, false, false}; uint32_t ty[4] = {0, 0, 0, 0}; };  // linear, embd, norm, else
enum { LINEAR = 0, EMBD = 1, NORM = 2, ELSE = 3 };

// `Operator::cast("k:v k:v")` — cast.rs:11-26 (regex (\w+):(\w+))
inline bool parse_cast_step(const std::string &spec, CastRule *r, std::string *err) {
    size_t i = 0;
    while (i < spec.size()) {
        while (i < spec.size() && !(std::isalnum((unsigned char)spec[i]) || spec[i] == '_')) i++;
        size_t k0 = i;
        while (i < spec.size() && (std::isalnum((unsigned char)spec[i]) || spec[i] == '_')) i++;
        if (i >= spec.size() || spec[i] != ':') continue;
        std::string key = spec.substr(k0, i - k0);
        size_t v0 = ++i;
        while (i < spec.size() && (std::isalnum((unsigned char)spec[i]) || spec[i] == '_')) i++;
        std::string val = spec.substr(v0, i - v0);
        if (key.empty() || val.empty()) continue;
        uint32_t ty;
        if (!parse_type(val, &ty)) { *err = "unknown tensor type '" + val + "'"; return false; }
        int slot = key == "linear" ? LINEAR : key == "embd" ? EMBD : key == "norm" ? NORM : key == "else" ? ELSE : -1;
        if (slot < 0) continue;  // the reference keeps unknown keys in the map and never reads them
        r->has[slot] = true;
        r->ty[slot] = ty;
    }
    return true;
}

// cast.rs:28-71: which rule applies to a tensor, by architecture
inline int classify(const std::string &arch, std::string_view name, size_t ndim) {
    if (arch == "clip") {
        if (name.substr(0, 2) == "v.") {
            std::string_view n = name.substr(2);
            if (n.find("embd") != n.npos) return EMBD;
            if (n.find("ln") != n.npos) return NORM;
            return LINEAR;
        }
        if (name.substr(0, 10) == "resampler.") return name.substr(10, 3) == "ln_" ? NORM : LINEAR;
        return ELSE;
    }
    if (name == "token_embd.weight" || name == "output.weight") return EMBD;
    if (ends_with(name, "_norm.weight") || ends_with(name, "_norm.bias")) return NORM;
    if (ndim > 1 || ends_with(name, ".bias")) return LINEAR;
    return ELSE;
}

struct StepError { int code; std::string msg; };  // thrown while applying operators (the reference panics)

// ---- tensor expressions ---------------------------------------------------------------------------
struct Node;
using NodeP = std::shared_ptr<const Node>;
struct Node {
    enum Kind { SOURCE, CAST, CONCAT, SPLIT, PERMUTE } kind = SOURCE;
    uint32_t type = 0;
    std::vector<uint64_t> shape;   // ggml order: shape[0] is the contiguous row
    std::vector<NodeP> in;
    int file = -1;                 // SOURCE: input file and absolute byte offset
    uint64_t file_off = 0;
    std::vector<uint32_t> chain;   // CAST: in[0]->type ... type
    size_t axis = 0;               // CONCAT / SPLIT
    uint64_t start = 0;            // SPLIT: first element along `axis`
    uint64_t nh = 0;               // PERMUTE: heads
};

inline uint64_t count(const std::vector<uint64_t> &shape) { uint64_t n = 1; for (uint64_t d : shape) n *= d; return n; }
// GGmlTypeSize::elements_to_bytes (ggus/src/tensor.rs:83-96)
inline uint64_t nbytes_of(uint32_t type, const std::vector<uint64_t> &shape) {
    uint64_t be = 1, bb = 1;
    gguf::type_size(type, &be, &bb);
    if (shape.empty()) return bb;
    return count(shape) / be * bb;
}
inline uint64_t nbytes_of(const Node &n) { return nbytes_of(n.type, n.shape); }
// merge.rs:359-364 `layout(ty, shape)`: shape[0] in blocks, element = one block
inline ndl::ArrayLayout block_layout(uint32_t type, std::vector<uint64_t> shape, uint64_t *unit) {
    uint64_t be = 1, bb = 1;
    gguf::type_size(type, &be, &bb);
    if (!shape.empty()) shape[0] /= be;
    *unit = bb;
    return ndl::ArrayLayout::new_contiguous(shape, bb);
}
// parts of a concat / split along `axis` sit next to each other in memory when no slower dim has extent > 1
inline bool axis_is_slowest(const std::vector<uint64_t> &shape, size_t axis) {
    for (size_t i = axis + 1; i < shape.size(); i++) if (shape[i] != 1) return false;
    return true;
}

inline NodeP make_cast(const NodeP &x, uint32_t to, const std::string &name) {
    if (x->type == to) return x;
    uint64_t be, bb;
    if (!gguf::type_size(to, &be, &bb) || ggq_type_nbytes(to, be) == 0 || ggq_type_nbytes(x->type, 256) == 0)
        throw StepError{GGQ_ERR_UNSUPPORTED, "cast chain of " + name + " has an unsupported type"};
    if (x->shape.empty() || x->shape[0] % be)  // cast.rs:142-143 `assert_eq!(row % N, 0)`
        throw StepError{GGQ_ERR_INDIVISIBLE, "row of " + name + " is not a multiple of the target block size"};
    auto n = std::make_shared<Node>();
    if (x->kind == Node::CAST) {  // one chain: intermediates stay on the device
        *n = *x;
        n->chain.push_back(to);
    } else if (x->kind == Node::CONCAT && axis_is_slowest(x->shape, x->axis)) {  // blocks never straddle rows
        *n = *x;
        for (auto &c : n->in) c = make_cast(c, to, name);
    } else {
        n->kind = Node::CAST;
        n->shape = x->shape;
        n->in = {x};
        n->chain = {x->type, to};
    }
    n->type = to;
    return n;
}

// merge.rs:288-325 `concat(axis, tensors)`
inline NodeP make_concat(size_t axis, const std::vector<NodeP> &parts) {
    auto n = std::make_shared<Node>();
    n->kind = Node::CONCAT;
    n->type = parts[0]->type;
    n->shape = parts[0]->shape;
    if (n->shape.size() == 1) axis = 0;
    if (axis >= n->shape.size()) throw StepError{GGQ_ERR_INVALID, "concat: tensor has no axis " + std::to_string(axis)};
    for (size_t k = 1; k < parts.size(); k++) {
        const Node &t = *parts[k];
        if (t.type != n->type) throw StepError{GGQ_ERR_INVALID, "concat: tensors of different types"};
        if (t.shape.size() != n->shape.size()) throw StepError{GGQ_ERR_INVALID, "concat: tensors of different rank"};
        for (size_t i = 0; i < n->shape.size(); i++) {
            if (i == axis) n->shape[i] += t.shape[i];
            else if (n->shape[i] != t.shape[i]) throw StepError{GGQ_ERR_INVALID, "concat: shapes differ off the concat axis"};
        }
    }
    n->axis = axis;
    n->in = parts;
    return n;
}

// one part of merge.rs:327-357 `split(axis, tensor, parts)`
inline NodeP make_split(const NodeP &x, size_t axis, uint64_t start, uint64_t len) {
    std::vector<uint64_t> shape = x->shape;
    shape[axis] = len;
    if (x->kind == Node::SOURCE && axis_is_slowest(x->shape, axis)) {  // a byte range of the file
        auto n = std::make_shared<Node>(*x);
        std::vector<uint64_t> before = x->shape;
        before[axis] = start;
        n->file_off += nbytes_of(x->type, before);
        n->shape = shape;
        return n;
    }
    if (x->kind == Node::CAST) {  // split rows first, cast only what is kept
        bool ok = axis > 0;
        if (!ok) {
            ok = true;
            for (uint32_t t : x->chain) {
                uint64_t be = 1, bb;
                gguf::type_size(t, &be, &bb);
                ok &= start % be == 0 && len % be == 0;
            }
        }
        if (ok) {
            auto n = std::make_shared<Node>(*x);
            n->in = {make_split(x->in[0], axis, start, len)};
            n->shape = shape;
            return n;
        }
    }
    if (x->kind == Node::CONCAT && x->axis == axis) {  // undoing a merge
        uint64_t at = 0;
        for (const NodeP &c : x->in) {
            if (at == start && c->shape[axis] == len) return c;
            at += c->shape[axis];
        }
    }
    auto n = std::make_shared<Node>();
    n->kind = Node::SPLIT;
    n->type = x->type;
    n->shape = shape;
    n->in = {x};
    n->axis = axis;
    n->start = start;
    return n;
}
inline std::vector<NodeP> split_parts(const NodeP &x, size_t axis, const std::vector<uint64_t> &parts) {
    if (x->shape.size() == 1) axis = 0;
    uint64_t sum = 0;
    for (uint64_t p : parts) sum += p;
    if (axis >= x->shape.size() || x->shape[axis] != sum) throw StepError{GGQ_ERR_INVALID, "split: parts do not add up to the axis"};  // merge.rs:333
    uint64_t be = 1, bb;
    gguf::type_size(x->type, &be, &bb);
    std::vector<NodeP> out;
    uint64_t at = 0;
    for (uint64_t p : parts) {
        if (axis == 0 && (at % be || p % be)) throw StepError{GGQ_ERR_INDIVISIBLE, "split: part is not a whole number of blocks"};
        out.push_back(make_split(x, axis, at, p));
        at += p;
    }
    return out;
}

// merge.rs:279-286 `distruct`
inline void distruct(const Node &t, uint64_t *c, uint64_t *r) {
    if (t.shape.size() == 1) { *c = 1; *r = t.shape[0]; }
    else if (t.shape.size() == 2) { *c = t.shape[0]; *r = t.shape[1]; }
    else throw StepError{GGQ_ERR_INVALID, "invalid tensor shape for a qkv operator (rank " + std::to_string(t.shape.size()) + ")"};
}
// merge.rs:239-250
inline NodeP merge_qkv(const NodeP &q, const NodeP &k, const NodeP &v) {
    uint64_t c, qr, kr, vr;
    distruct(*q, &c, &qr); distruct(*k, &c, &kr); distruct(*v, &c, &vr);
    if (kr == 0 || qr % kr != 0 || qr < kr || kr != vr) throw StepError{GGQ_ERR_INVALID, "merge-linear: q/k/v row counts do not fit"};
    return make_concat(1, {q, k, v});
}
// merge.rs:267-272
inline std::vector<NodeP> split_qkv(const NodeP &t, uint64_t nh, uint64_t nkvh) {
    uint64_t c, r;
    distruct(*t, &c, &r);
    const uint64_t dh = r / (nh + nkvh * 2);
    return split_parts(t, 1, {nh * dh, nkvh * dh, nkvh * dh});
}
// permute_qk.rs:46-69
inline NodeP make_permute(const NodeP &x, uint64_t nh) {
    uint64_t be = 1, bb;
    gguf::type_size(x->type, &be, &bb);
    uint64_t c, r;
    if (x->shape.size() == 1) { c = 1; r = x->shape[0]; }
    else if (x->shape.size() == 2) { c = x->shape[0]; r = x->shape[1]; }
    else throw StepError{GGQ_ERR_UNSUPPORTED, "permute-qk of a tensor of rank > 2 (todo!() in the reference)"};
    if (c % be) throw StepError{GGQ_ERR_INDIVISIBLE, "permute-qk: row is not a whole number of blocks"};  // tensor.rs:92
    if (nh == 0 || r % (nh * 2)) throw StepError{GGQ_ERR_INVALID, "permute-qk: rows are not a multiple of 2 * heads"};
    auto n = std::make_shared<Node>();
    n->kind = Node::PERMUTE;
    n->type = x->type;
    n->shape = x->shape;
    n->in = {x};
    n->nh = nh;
    return n;
}

inline uint64_t cast_elems_of(const Node &n) {
    uint64_t e = n.kind == Node::CAST ? count(n.shape) : 0;
    for (const NodeP &c : n.in) e += cast_elems_of(*c);
    return e;
}

struct Tensor {                 // one tensor of the content being built (utils/mod.rs:97-101)
    std::string name;
    NodeP node;
    uint64_t out_nbytes = 0;
    int shard = 0;
    uint64_t out_off = 0;
};

// `NAME.(weight|bias)$` with NAME one of `alts` (the MERGE / SPLIT / QK regexes of merge.rs:8-10, permute_qk.rs:24-25)
inline bool match_linear(const std::string &name, std::initializer_list<const char *> alts, std::string *pre, std::string *which, std::string *wb) {
    std::string_view n(name), tail;
    if (ends_with(n, ".weight")) tail = "weight";
    else if (ends_with(n, ".bias")) tail = "bias";
    else return false;
    n.remove_suffix(tail.size() + 1);
    for (const char *a : alts)
        if (ends_with(n, a)) {
            *pre = std::string(n.substr(0, n.size() - std::string_view(a).size()));
            *which = a;
            *wb = std::string(tail);
            return true;
        }
    return false;
}

// GGufMetaMapExt::get_usize (ggus/src/metadata/collection.rs:40-72): any integer type, must fit usize
inline bool get_usize(const std::vector<const gguf::MetaKV *> &kvs, const std::string &key, uint64_t *out, bool *exists) {
    *exists = false;
    for (const gguf::MetaKV *kv : kvs) {
        if (kv->key != key) continue;
        *exists = true;
        int64_t sv = 0;
        switch (kv->type) {
            case gguf::U8: *out = kv->value[0]; return true;
            case gguf::U16: { uint16_t v; memcpy(&v, kv->value, 2); *out = v; return true; }
            case gguf::U32: { uint32_t v; memcpy(&v, kv->value, 4); *out = v; return true; }
            case gguf::U64: { uint64_t v; memcpy(&v, kv->value, 8); *out = v; return true; }
            case gguf::I8: sv = (int8_t)kv->value[0]; break;
            case gguf::I16: { int16_t v; memcpy(&v, kv->value, 2); sv = v; break; }
            case gguf::I32: { int32_t v; memcpy(&v, kv->value, 4); sv = v; break; }
            case gguf::I64: { int64_t v; memcpy(&v, kv->value, 8); sv = v; break; }
            default: return false;  // TypeMismatch
        }
        if (sv < 0) return false;  // OutOfRange
        *out = (uint64_t)sv;
        return true;
    }
    return false;
}

// `llm_attention_head_count` / `_kv` as merge.rs:41-46 and permute_qk.rs:13-18 read them
inline void head_counts(const std::vector<const gguf::MetaKV *> &kvs, const std::string &arch, uint64_t *nh, uint64_t *nkvh) {
    bool exists;
    if (!get_usize(kvs, arch + ".attention.head_count", nh, &exists))
        throw StepError{GGQ_ERR_INVALID, exists ? "bad type for " + arch + ".attention.head_count" : "NotExist: " + arch + ".attention.head_count"};
    if (!get_usize(kvs, arch + ".attention.head_count_kv", nkvh, &exists)) {
        if (exists) throw StepError{GGQ_ERR_INVALID, "bad type for " + arch + ".attention.head_count_kv"};
        *nkvh = *nh;
    }
    if (*nh == 0 || *nkvh == 0) throw StepError{GGQ_ERR_INVALID, "attention head count is zero"};
}

// merge.rs:22-39 + the collectors of merge.rs:106-237.  Parts are gathered per name prefix and per
// (layer kind, weight|bias); a group is emitted where its LAST part stood.  Incomplete groups are
// appended unmerged at the end (the reference iterates a HashMap there, i.e. in arbitrary order; here:
// first-seen order).
inline void apply_merge(std::vector<Tensor> &tensors) {
    struct Group { std::string pre, wb; int layer; NodeP part[3]; std::string name[3]; bool open = true; };
    std::vector<Group> groups;
    std::vector<Tensor> out;
    for (Tensor &t : tensors) {
        std::string pre, which, wb;
        if (!match_linear(t.name, {"attn_q", "attn_k", "attn_v", "ffn_gate_exps", "ffn_up_exps", "ffn_gate", "ffn_up"}, &pre, &which, &wb)) {
            out.push_back(std::move(t));
            continue;
        }
        const int layer = which.rfind("attn_", 0) == 0 ? 0 : ends_with(which, "_exps") ? 2 : 1;
        const int idx = which == "attn_q" || which == "ffn_gate" || which == "ffn_gate_exps" ? 0 : which == "attn_v" ? 2 : 1;
        Group *g = nullptr;
        for (Group &c : groups) if (c.open && c.pre == pre && c.wb == wb && c.layer == layer) g = &c;
        if (!g) {
            groups.push_back(Group{pre, wb, layer, {}, {}, true});
            g = &groups.back();
        }
        g->part[idx] = t.node;
        g->name[idx] = t.name;
        const bool done = layer == 0 ? (g->part[0] && g->part[1] && g->part[2]) : (g->part[0] && g->part[1]);
        if (!done) continue;
        g->open = false;
        Tensor m;
        if (layer == 0) {
            m.name = pre + "attn_qkv." + wb;
            m.node = merge_qkv(g->part[0], g->part[1], g->part[2]);
        } else {
            if (layer == 1 && (g->part[0]->shape.size() < 2 || g->part[0]->shape[1] != g->part[1]->shape[1]))  // merge.rs:256
                throw StepError{GGQ_ERR_INVALID, "merge-linear: ffn_gate and ffn_up differ in rows (" + g->name[0] + ")"};
            m.name = pre + (layer == 1 ? "ffn_gate_up." : "ffn_gate_up_exps.") + wb;
            m.node = make_concat(1, {g->part[0], g->part[1]});
        }
        out.push_back(std::move(m));
    }
    for (Group &g : groups)
        if (g.open)
            for (int i = 0; i < 3; i++)
                if (g.part[i]) { Tensor t; t.name = g.name[i]; t.node = g.part[i]; out.push_back(std::move(t)); }
    tensors = std::move(out);
}

// merge.rs:40-81.  SPLIT = (attn_qkv|ffn_gate_up).(weight|bias)$ — expert tensors are not matched.
inline void apply_split(std::vector<Tensor> &tensors, uint64_t nh, uint64_t nkvh) {
    std::vector<Tensor> out;
    auto put = [&](const std::string &name, const NodeP &node) { Tensor t; t.name = name; t.node = node; out.push_back(std::move(t)); };
    for (Tensor &t : tensors) {
        std::string pre, which, wb;
        if (!match_linear(t.name, {"attn_qkv", "ffn_gate_up"}, &pre, &which, &wb)) {
            out.push_back(std::move(t));
            continue;
        }
        if (which == "attn_qkv") {
            auto p = split_qkv(t.node, nh, nkvh);
            put(pre + "attn_q." + wb, p[0]);
            put(pre + "attn_k." + wb, p[1]);
            put(pre + "attn_v." + wb, p[2]);
        } else {
            if (t.node->shape.size() < 2) throw StepError{GGQ_ERR_INVALID, "split-linear: " + t.name + " has no second axis"};  // merge.rs:275
            const uint64_t r = t.node->shape[1] / 2;
            auto p = split_parts(t.node, 1, {r, r});
            put(pre + "ffn_gate." + wb, p[0]);
            put(pre + "ffn_up." + wb, p[1]);
        }
    }
    tensors = std::move(out);
}

// permute_qk.rs:11-44
inline void apply_permute(std::vector<Tensor> &tensors, uint64_t nh, uint64_t nkvh) {
    for (Tensor &t : tensors) {
        std::string pre, which, wb;
        if (!match_linear(t.name, {"attn_qkv", "attn_q", "attn_k"}, &pre, &which, &wb)) continue;
        if (which == "attn_q") t.node = make_permute(t.node, nh);
        else if (which == "attn_k") t.node = make_permute(t.node, nkvh);
        else {
            auto p = split_qkv(t.node, nh, nkvh);
            t.node = merge_qkv(make_permute(p[0], nh), make_permute(p[1], nkvh), p[2]);
        }
    }
}

}  // namespace tensor_ops
