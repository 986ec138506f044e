// gguf.hpp — zero-copy GGUF v3 parse and the writer-side layout planning, C++17, header only.
//
// Mirrors the part of the reference's `ggus` crate the convert path needs (paths under
// /root/reference/): header (ggus/src/header.rs:7-64), meta KVs with `general.alignment`
// (ggus/src/file.rs:82-98, metadata/mod.rs:13-15), tensor infos (ggus/src/tensor.rs:193-204),
// data-region location (file.rs:100-136), `pad` (ggus/src/lib.rs:28-31), and the writer's layout
// (ggus/src/write/file_writer.rs:96-136, writer.rs:42-105).  Typed metadata accessors, file-name
// grammar and sharding are out of scope.
#pragma once

#include <cstdint>
#include <cstring>
#include <set>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

namespace gguf {

constexpr uint64_t DEFAULT_ALIGNMENT = 32;            // ggus/src/metadata/mod.rs:13-15
constexpr const char *GENERAL_ALIGNMENT = "general.alignment";

inline uint64_t pad(uint64_t pos, uint64_t align) { return (align - pos % align) % align; }  // ggus/src/lib.rs:28-31

// GGufError (ggus/src/file.rs:25-40)
struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

enum MetaType : uint32_t { U8 = 0, I8, U16, I16, U32, I32, F32, BOOL, STRING, ARRAY, U64, I64, F64 };

// GGmlType::size() (ggus/src/tensor.rs:102-144): {block elements, block bytes} — the whole table, so that a
// tensor of a type this library has no codec for still passes through untouched (the reference hands such
// tensors on as borrowed bytes).  Sizes are size_of of the REFERENCE's repr(C) structs
// (ggml-quants/src/structs/iq*.rs), which is what its reader computes nbytes from; IQ3XXS / IQ4NL / IQ4XS are
// wider there than in upstream ggml (194 / 34 / 264 vs 98 / 18 / 136 bytes).  Q4_0_4_4 / _4_8 / _8_8 (31..33)
// are `todo!()` in the reference's table and stay unsupported.
inline bool type_size(uint32_t ty, uint64_t *elems, uint64_t *bytes) {
    switch (ty) {
        case 0: *elems = 1; *bytes = 4; return true;     // F32
        case 1: *elems = 1; *bytes = 2; return true;     // F16
        case 2: *elems = 32; *bytes = 18; return true;   // Q4_0
        case 3: *elems = 32; *bytes = 20; return true;   // Q4_1
        case 6: *elems = 32; *bytes = 22; return true;   // Q5_0
        case 7: *elems = 32; *bytes = 24; return true;   // Q5_1
        case 8: *elems = 32; *bytes = 34; return true;   // Q8_0
        case 9: *elems = 32; *bytes = 36; return true;   // Q8_1
        case 10: *elems = 256; *bytes = 84; return true;  // Q2K
        case 11: *elems = 256; *bytes = 110; return true; // Q3K
        case 12: *elems = 256; *bytes = 144; return true; // Q4K
        case 13: *elems = 256; *bytes = 176; return true; // Q5K
        case 14: *elems = 256; *bytes = 210; return true; // Q6K
        case 15: *elems = 256; *bytes = 290; return true; // Q8K (reference layout)
        case 16: *elems = 256; *bytes = 66; return true;  // IQ2XXS  f16 + [u16; 32]
        case 17: *elems = 256; *bytes = 74; return true;  // IQ2XS   f16 + [u16; 32] + [u8; 8]
        case 18: *elems = 256; *bytes = 194; return true; // IQ3XXS  f16 + [u16; 96]
        case 19: *elems = 256; *bytes = 50; return true;  // IQ1S    f16 + [u8; 32] + [u16; 8]
        case 20: *elems = 32; *bytes = 34; return true;   // IQ4NL   f16 + [u16; 16]
        case 21: *elems = 256; *bytes = 110; return true; // IQ3S    f16 + [u8; 64] + [u8; 8] + [u8; 32] + [u8; 4]
        case 22: *elems = 256; *bytes = 82; return true;  // IQ2S    f16 + [u8; 64] + [u8; 8] + [u8; 8]
        case 23: *elems = 256; *bytes = 264; return true; // IQ4XS   f16 + u16 + [u8; 4] + [u16; 128]
        case 24: *elems = 1; *bytes = 1; return true;    // I8
        case 25: *elems = 1; *bytes = 2; return true;    // I16
        case 26: *elems = 1; *bytes = 4; return true;    // I32
        case 27: *elems = 1; *bytes = 8; return true;    // I64
        case 28: *elems = 1; *bytes = 8; return true;    // F64
        case 29: *elems = 256; *bytes = 56; return true;  // IQ1M    [u8; 32] + [u8; 16] + [u8; 8]
        case 30: *elems = 1; *bytes = 2; return true;    // BF16
    }
    return false;
}

struct MetaKV {
    std::string_view key;
    uint32_t type;
    const uint8_t *value;   // value bytes (after the type tag)
    uint64_t value_len;
    const uint8_t *raw;     // whole record: key string + type + value, as it sits in the file
    uint64_t raw_len;
};

struct TensorInfo {
    std::string_view name;
    std::vector<uint64_t> shape;  // shape[0] is ggml ne[0], the contiguous row
    uint32_t type;
    uint64_t offset;              // relative to the data region
    uint64_t nbytes;
    uint64_t n_elems() const { uint64_t n = 1; for (uint64_t d : shape) n *= d; return n; }
};

class Reader {
   public:
    Reader(const uint8_t *p, uint64_t n) : p_(p), n_(n), pos_(0) {}
    template <class T> T read() {
        need(sizeof(T));
        T v;
        std::memcpy(&v, p_ + pos_, sizeof(T));
        pos_ += sizeof(T);
        return v;
    }
    std::string_view read_str() {
        const uint64_t len = read<uint64_t>();
        need(len);
        std::string_view s(reinterpret_cast<const char *>(p_ + pos_), len);
        pos_ += len;
        return s;
    }
    void skip(uint64_t n) { need(n); pos_ += n; }
    uint64_t pos() const { return pos_; }
    uint64_t remaining() const { return n_ - pos_; }
    const uint8_t *here() const { return p_ + pos_; }

   private:
    void need(uint64_t n) const { if (n > n_ - pos_) throw Error("Reading(Eos): unexpected end of file"); }
    const uint8_t *p_;
    uint64_t n_, pos_;
};

inline uint64_t scalar_size(uint32_t ty) {
    switch (ty) {
        case U8: case I8: case BOOL: return 1;
        case U16: case I16: return 2;
        case U32: case I32: case F32: return 4;
        case U64: case I64: case F64: return 8;
    }
    return 0;
}
// skip one metadata value of type `ty` (ggus/src/read.rs:10-92 `read_meta_kv` value walk)
inline void skip_value(Reader &r, uint32_t ty) {
    if (ty == STRING) { r.read_str(); return; }
    if (ty == ARRAY) {
        const uint32_t et = r.read<uint32_t>();
        const uint64_t n = r.read<uint64_t>();
        if (et == STRING) { for (uint64_t i = 0; i < n; i++) r.read_str(); return; }
        if (et == ARRAY) { for (uint64_t i = 0; i < n; i++) skip_value(r, ARRAY); return; }
        const uint64_t sz = scalar_size(et);
        if (!sz) throw Error("unknown metadata array element type");
        r.skip(sz * n);
        return;
    }
    const uint64_t sz = scalar_size(ty);
    if (!sz) throw Error("unknown metadata value type");
    r.skip(sz);
}

// `GGuf::new` (ggus/src/file.rs:66-145): borrows `data`
struct File {
    uint32_t version = 0;
    uint64_t alignment = DEFAULT_ALIGNMENT;
    std::vector<MetaKV> meta_kvs;
    std::vector<TensorInfo> tensors;
    const uint8_t *data = nullptr;  // start of the tensor data region
    uint64_t data_len = 0;

    static File parse(const uint8_t *bytes, uint64_t len) {
        File f;
        Reader r(bytes, len);
        char magic[4];
        for (char &c : magic) c = (char)r.read<uint8_t>();
        if (std::memcmp(magic, "GGUF", 4) != 0) throw Error("MagicMismatch");
        f.version = r.read<uint32_t>();
        if (f.version == 0x03000000u) throw Error("EndianNotSupport");
        if (f.version != 3) throw Error("VersionNotSupport");
        const uint64_t n_tensors = r.read<uint64_t>(), n_kvs = r.read<uint64_t>();
        std::set<std::string_view> seen;
        for (uint64_t i = 0; i < n_kvs; i++) {
            MetaKV kv;
            kv.raw = r.here();
            kv.key = r.read_str();
            kv.type = r.read<uint32_t>();
            kv.value = r.here();
            skip_value(r, kv.type);
            kv.value_len = (uint64_t)(r.here() - kv.value);
            kv.raw_len = (uint64_t)(r.here() - kv.raw);
            if (kv.key == GENERAL_ALIGNMENT) {
                if (kv.type == U32) { uint32_t a; std::memcpy(&a, kv.value, 4); f.alignment = a; }
                else if (kv.type == U64) { uint64_t a; std::memcpy(&a, kv.value, 8); f.alignment = a; }
                else throw Error("AlignmentTypeMismatch");
            }
            if (!seen.insert(kv.key).second) throw Error("DuplicateMetaKey(" + std::string(kv.key) + ")");
            f.meta_kvs.push_back(kv);
        }
        if (f.alignment == 0) throw Error("general.alignment is zero");
        std::set<std::string_view> names;
        for (uint64_t i = 0; i < n_tensors; i++) {
            TensorInfo t;
            t.name = r.read_str();
            const uint32_t ndim = r.read<uint32_t>();
            for (uint32_t d = 0; d < ndim; d++) t.shape.push_back(r.read<uint64_t>());
            t.type = r.read<uint32_t>();
            t.offset = r.read<uint64_t>();
            uint64_t be, bb;
            if (!type_size(t.type, &be, &bb)) throw Error("unsupported tensor type " + std::to_string(t.type) + " for " + std::string(t.name));
            // elements_to_bytes (tensor.rs:83-96): blocks run along shape[0]
            if (t.shape.empty()) { if (be != 1) throw Error("scalar tensor of a block type"); t.nbytes = bb; }
            else {
                if (t.shape[0] % be) throw Error("shape[0] is not a multiple of the block size for " + std::string(t.name));
                t.nbytes = t.n_elems() / be * bb;
            }
            if (t.offset + t.nbytes > f.data_len) f.data_len = t.offset + t.nbytes;
            if (!names.insert(t.name).second) throw Error("DuplicateTensorName(" + std::string(t.name) + ")");
            f.tensors.push_back(std::move(t));
        }
        if (!f.tensors.empty()) r.skip(pad(r.pos(), f.alignment));
        if (r.remaining() < f.data_len) throw Error("Reading(Eos): tensor data is truncated");
        f.data = r.here();
        return f;
    }

    const MetaKV *find(std::string_view key) const {
        for (const auto &kv : meta_kvs) if (kv.key == key) return &kv;
        return nullptr;
    }
    // string-valued KV or empty
    std::string_view get_str(std::string_view key) const {
        const MetaKV *kv = find(key);
        if (!kv || kv->type != STRING || kv->value_len < 8) return {};
        uint64_t n;
        std::memcpy(&n, kv->value, 8);
        return std::string_view(reinterpret_cast<const char *>(kv->value + 8), n);
    }
};

// ---- writer: byte sink + the reference's layout --------------------------------------------------
class Sink {
   public:
    explicit Sink(uint8_t *base = nullptr) : base_(base), pos_(0) {}
    void bytes(const void *p, uint64_t n) { if (base_) std::memcpy(base_ + pos_, p, n); pos_ += n; }
    template <class T> void put(T v) { bytes(&v, sizeof v); }
    void str(std::string_view s) { put<uint64_t>(s.size()); bytes(s.data(), s.size()); }
    void zeros(uint64_t n) { if (base_) std::memset(base_ + pos_, 0, n); pos_ += n; }
    uint64_t pos() const { return pos_; }

   private:
    uint8_t *base_;  // nullptr = simulate (ggus/src/write/simulator.rs)
    uint64_t pos_;
};

struct OutTensor {
    std::string_view name;
    const std::vector<uint64_t> *shape;
    uint32_t type;
    uint64_t nbytes;
    uint64_t file_offset = 0;  // absolute position of the tensor's bytes in the output file
};

// Emits header + `general.alignment` + KVs + tensor infos exactly as write.rs:75-90 /
// file_writer.rs:96-108 do for a single shard, fills in every tensor's absolute file offset and
// returns the total file size.  With a null sink it only plans (the reference's simulator).
inline uint64_t write_front(Sink &s, uint64_t alignment, const std::vector<const MetaKV *> &kvs, std::vector<OutTensor> &tensors) {
    s.bytes("GGUF", 4);
    s.put<uint32_t>(3);
    s.put<uint64_t>(tensors.size());
    s.put<uint64_t>(kvs.size() + 1);
    s.str(GENERAL_ALIGNMENT);            // writer.rs:42-49
    s.put<uint32_t>(U32);
    s.put<uint32_t>((uint32_t)alignment);
    for (const MetaKV *kv : kvs) s.bytes(kv->raw, kv->raw_len);
    uint64_t off = 0;
    std::vector<uint64_t> rel(tensors.size());
    for (size_t i = 0; i < tensors.size(); i++) {  // file_writer.rs:96-108
        off += pad(off, alignment);
        rel[i] = off;
        s.str(tensors[i].name);
        s.put<uint32_t>((uint32_t)tensors[i].shape->size());
        for (uint64_t d : *tensors[i].shape) s.put<uint64_t>(d);
        s.put<uint32_t>(tensors[i].type);
        s.put<uint64_t>(off);
        off += tensors[i].nbytes;
    }
    // data: each tensor is preceded by padding of the FILE position (file_writer.rs:121-126)
    uint64_t pos = s.pos();
    for (size_t i = 0; i < tensors.size(); i++) {
        pos += pad(pos, alignment);
        tensors[i].file_offset = pos;
        pos += tensors[i].nbytes;
    }
    return pos;
}

}  // namespace gguf
