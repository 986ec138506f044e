// convert.cpp — whole-file GGUF conversion on the GPUs: the `xtask convert --steps "a -> b -> c"`
// path (/root/reference/xtask/src/convert.rs:24-58 → utils/mod.rs:36-59 → operator/*.rs →
// utils/write.rs:6-100) for the operators that touch tensor bytes: `cast:` (operator/cast.rs:28-138),
// `merge-linear` / `split-linear` (operator/merge.rs) and `permute-qk` (operator/permute_qk.rs).
//
// Like the reference, operators do not move data when they are applied: each tensor becomes a small
// expression (source bytes of a file, cast chain, concat, split, row permutation) that is evaluated
// when the tensor is written (the reference's `DataPromise::lazy`, utils/mod.rs:104-138).  Expressions
// are normalised while they are built — casts and splits commute with whole-row moves, so a split
// of a file tensor is a sub-range of the file, a cast of a concat is a concat of casts, and a concat
// along the slowest axis writes its parts next to each other — which leaves exactly two ways to
// produce bytes: the streaming pipeline (file → pinned → H2D → cast kernels → D2H → file) and, for
// tensors whose rows really are rearranged (permute-qk, 3-D expert merges), a device-resident
// evaluation: upload, strided-copy kernels (rearrange.cu) and cast kernels, download.
//
// What changes against the reference pipeline:
//   * all casts of a tensor are one device-resident chain (F16→Q8_0→F32→F16 never returns to the host
//     in between; the reference materialises every intermediate in an anonymous mmap);
//   * tensor offsets of the output are planned up front (the reference's simulator does the same,
//     write.rs:23-51), so every tensor is written at its final place with pwrite, in any order;
//   * tensors are spread over worker threads (largest first, several per GPU), no collective.
// The operator planning itself (expressions, grouping, naming) is host/tensor_ops.hpp.
// Output bytes are identical to what the reference writer emits for a single shard: header,
// `general.alignment` first, the other KVs in input order (minus `split.*`), infos, padded data.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ggq.h"
#include "../host/array_layout.hpp"
#include "../host/gguf.hpp"
#include "../host/tensor_ops.hpp"
#include "ggq_internal.h"

namespace {

using namespace tensor_ops;

thread_local std::string t_cerr;

struct Mapping {
    uint8_t *p = nullptr;
    size_t len = 0;
    ~Mapping() { if (p && len) munmap(p, len); }
};

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct InFile {
    Mapping map;
    gguf::File f;
    int fd = -1;
    int direct_fd = -1;  // the same file opened O_DIRECT (ggq_convert_options.direct_io), or -1
    uint64_t data_off = 0, size = 0;
    ~InFile() { if (fd >= 0) close(fd); if (direct_fd >= 0) close(direct_fd); }
};

thread_local uint64_t t_read_ns = 0, t_write_ns = 0;  // this worker's time inside pread / pwrite (ggq_convert_stats)
struct ScopedNs {
    uint64_t &acc;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit ScopedNs(uint64_t &a) : acc(a) {}
    ~ScopedNs() { acc += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); }
};

bool pread_all(int fd, void *buf, size_t n, uint64_t off) {
    ScopedNs timer(t_read_ns);
    char *p = static_cast<char *>(buf);
    while (n) {
        ssize_t r = pread(fd, p, n, (off_t)off);
        if (r <= 0) return false;
        p += r; off += (uint64_t)r; n -= (size_t)r;
    }
    return true;
}
bool pwrite_all(int fd, const void *buf, size_t n, uint64_t off) {
    ScopedNs timer(t_write_ns);
    const char *p = static_cast<const char *>(buf);
    while (n) {
        ssize_t r = pwrite(fd, p, n, (off_t)off);
        if (r <= 0) return false;
        p += r; off += (uint64_t)r; n -= (size_t)r;
    }
    return true;
}

// bytes of one tensor-info record (writer.rs:74-86)
uint64_t info_bytes(const Tensor &t) { return 8 + t.name.size() + 4 + 8 * t.node->shape.size() + 4 + 8; }
constexpr uint64_t ALIGNMENT_KV_BYTES = 8 + 17 + 4 + 4;  // "general.alignment": key string + type + u32

// GGufFileSimulator / GGufTensorSimulator (ggus/src/write/simulator.rs:26-96)
struct Simulator {
    uint64_t alignment, written, n = 0;
    std::vector<uint64_t> data;
    Simulator(uint64_t align, uint64_t kv_bytes) : alignment(align), written(24 + ALIGNMENT_KV_BYTES + kv_bytes) {}
    void write_tensor(const Tensor &t) { written += info_bytes(t); data.push_back(t.out_nbytes); }
    uint64_t written_bytes() const {
        uint64_t total = written;
        for (uint64_t len : data) { total += gguf::pad(total, alignment); total += len; }
        return total;
    }
};

struct Step {
    enum Kind { CAST, MERGE, SPLIT, PERMUTE } kind;
    CastRule rule;
};

// ---- producing a tensor's bytes ---------------------------------------------------------------------

// Device memory the resident evaluation of `n` holds at its peak: the result plus the temporaries of
// the deepest operand chain (each operator frees its operand as soon as it has consumed it).
uint64_t resident_peak_bytes(const Node &n) {
    if (n.kind == Node::SOURCE) return 0;  // uploaded straight into the caller's buffer
    uint64_t worst = 0;
    for (const NodeP &c : n.in) worst = std::max(worst, nbytes_of(*c) + resident_peak_bytes(*c));
    if (n.kind == Node::CAST) {  // intermediates of a multi-hop chain: two live at a time, F32 is the widest
        if (n.chain.size() > 2) worst += 2 * count(n.shape) * 4;
    }
    return worst;
}

// Worker threads evaluate different tensors at the same time; a per-device budget keeps the sum of their
// peaks inside the GPU's free memory (a 256-expert merge is tens of GB).  A tensor that is larger than
// the whole budget runs alone.
class ResidentBudget {
   public:
    void acquire(int dev, uint64_t bytes) {
        std::unique_lock<std::mutex> lk(mu_);
        if (limit_.count(dev) == 0) {
            const size_t free_b = ggq::device_free_bytes();
            limit_[dev] = free_b ? (uint64_t)(free_b * 0.8) : UINT64_MAX;
            if (const char *env = getenv("GGQ_RESIDENT_BUDGET_MB")) limit_[dev] = (uint64_t)atoll(env) << 20;  // tests / tuning
        }
        cv_.wait(lk, [&] { return used_[dev] == 0 || used_[dev] + bytes <= limit_[dev]; });
        used_[dev] += bytes;
    }
    void release(int dev, uint64_t bytes) {
        { std::lock_guard<std::mutex> lk(mu_); used_[dev] -= bytes; }
        cv_.notify_all();
    }

   private:
    std::mutex mu_;
    std::condition_variable cv_;
    std::map<int, uint64_t> limit_, used_;
};

// O_DIRECT read of [off, off + n): the enclosing 4 KiB-aligned range goes straight into the (page-aligned) pinned buffer;
// returns where `off` landed inside it, or -1.  Reads past the end of the file come back short, which is fine as long as
// the requested bytes arrived.
long pread_direct(int fd, void *pinned, size_t cap, size_t n, uint64_t off) {
    ScopedNs timer(t_read_ns);
    constexpr uint64_t A = 4096;
    const uint64_t lo = off & ~(A - 1), hi = (off + n + A - 1) & ~(A - 1);
    if (hi - lo > cap) return -1;
    char *p = static_cast<char *>(pinned);
    uint64_t got = 0;
    while (lo + got < off + n) {
        const ssize_t r = pread(fd, p + got, (size_t)(hi - lo - got), (off_t)(lo + got));
        if (r <= 0) return -1;
        got += (uint64_t)r;
        if (got % A) break;  // short read at the end of the file: nothing more to get
    }
    return lo + got >= off + n ? (long)(off - lo) : -1;
}

struct IoCtx {
    const std::vector<int> *in_fds;
    const std::vector<int> *in_direct_fds = nullptr;
    int out_fd;
    std::vector<uint8_t> *copy_buf;
    ggq::Resident *res = nullptr;  // created on first use by the worker
    std::function<ggq::Resident *()> resident;
    bool used_resident = false;  // set by emit(): the tensor went through the device-resident path
    ResidentBudget *budget = nullptr;
    int device = 0;
};

// device-resident evaluation: the node's bytes end up, contiguous, at d_dst
int eval_into(const Node &n, void *d_dst, IoCtx &io) {
    ggq::Resident *res = io.resident();
    if (!res) return GGQ_ERR_CUDA;
    int rc = GGQ_OK;
    auto with_child = [&](const Node &c, void **tmp) {
        int r = res->alloc(nbytes_of(c), tmp);
        return r != GGQ_OK ? r : eval_into(c, *tmp, io);
    };
    switch (n.kind) {
        case Node::SOURCE: {
            const int fd = (*io.in_fds)[n.file];
            const uint64_t base = n.file_off;
            return res->upload(d_dst, nbytes_of(n), [fd, base](void *pinned, size_t off, size_t len) { return pread_all(fd, pinned, len, base + off); });
        }
        case Node::CAST: {
            void *tmp;
            if ((rc = with_child(*n.in[0], &tmp)) != GGQ_OK) return rc;
            rc = res->cast(n.chain.data(), (int)n.chain.size(), count(n.shape), d_dst, tmp);
            res->free(tmp);
            return rc;
        }
        case Node::CONCAT: {  // merge.rs:303-321
            uint64_t unit;
            const ndl::ArrayLayout whole = block_layout(n.type, n.shape, &unit);
            uint64_t be = 1, bb;
            gguf::type_size(n.type, &be, &bb);
            std::vector<uint64_t> parts;
            for (const NodeP &c : n.in) parts.push_back(n.axis == 0 ? c->shape[0] / be : c->shape[n.axis]);
            const auto views = whole.split(n.axis, parts);
            for (size_t k = 0; k < n.in.size(); k++) {
                if (views[k].is_dense(unit)) {  // the part is one byte range of the result: produce it in place
                    if ((rc = eval_into(*n.in[k], static_cast<char *>(d_dst) + views[k].offset, io)) != GGQ_OK) return rc;
                    continue;
                }
                void *tmp;
                if ((rc = with_child(*n.in[k], &tmp)) != GGQ_OK) return rc;
                uint64_t u2;
                rc = res->rearrange(d_dst, views[k].c(), tmp, block_layout(n.in[k]->type, n.in[k]->shape, &u2).c(), unit);
                res->free(tmp);
                if (rc != GGQ_OK) return rc;
            }
            return GGQ_OK;
        }
        case Node::SPLIT: {  // merge.rs:335-356
            const Node &c = *n.in[0];
            uint64_t unit, be = 1, bb;
            gguf::type_size(n.type, &be, &bb);
            ndl::ArrayLayout src = block_layout(c.type, c.shape, &unit);
            const uint64_t start = n.axis == 0 ? n.start / be : n.start, len = n.axis == 0 ? n.shape[0] / be : n.shape[n.axis];
            src.offset += (int64_t)start * src.strides[n.axis];
            src.shape[n.axis] = len;
            void *tmp;
            if ((rc = with_child(c, &tmp)) != GGQ_OK) return rc;
            rc = res->rearrange(d_dst, block_layout(n.type, n.shape, &unit).c(), tmp, src.c(), unit);
            res->free(tmp);
            return rc;
        }
        case Node::PERMUTE: {  // permute_qk.rs:46-60: rows (bytes) tiled [r/nh/2, 2, nh], the two inner tiles swapped
            const uint64_t c = n.shape.size() == 1 ? nbytes_of(n.type, {1}) : nbytes_of(n.type, {n.shape[0]});
            const uint64_t r = n.shape.size() == 1 ? n.shape[0] : n.shape[1];
            const ndl::ArrayLayout src = ndl::ArrayLayout::new_contiguous({c, r}, 1).tile_le(1, {r / n.nh / 2, 2, n.nh}).transpose({2, 1});
            const ndl::ArrayLayout dst = ndl::ArrayLayout::new_contiguous(src.shape, 1);
            void *tmp;
            if ((rc = with_child(*n.in[0], &tmp)) != GGQ_OK) return rc;
            rc = res->rearrange(d_dst, dst.c(), tmp, src.c(), 1);
            res->free(tmp);
            return rc;
        }
    }
    return GGQ_ERR_INVALID;
}

// write the node's bytes at out_off of the output file
int emit(const Node &n, uint64_t out_off, IoCtx &io) {
    const uint64_t nbytes = nbytes_of(n);
    const int ofd = io.out_fd;
    if (nbytes == 0) return GGQ_OK;
    if (n.kind == Node::SOURCE) {  // untouched bytes: file to file
        const int ifd = (*io.in_fds)[n.file];
        constexpr size_t CH = size_t(8) << 20;
        io.copy_buf->resize(std::min<uint64_t>(CH, nbytes));
        for (uint64_t off = 0; off < nbytes; off += CH) {
            const size_t len = (size_t)std::min<uint64_t>(CH, nbytes - off);
            if (!pread_all(ifd, io.copy_buf->data(), len, n.file_off + off) || !pwrite_all(ofd, io.copy_buf->data(), len, out_off + off)) {
                t_cerr = "I/O error while copying tensor bytes";
                return GGQ_ERR_INVALID;
            }
        }
        return GGQ_OK;
    }
    if (n.kind == Node::CAST && n.in[0]->kind == Node::SOURCE) {  // the streaming pipeline
        const Node &src = *n.in[0];
        const int ifd = (*io.in_fds)[src.file];
        const uint64_t src_off = src.file_off;
        ggq::ChainIO cio;
        cio.read = [&](void *pinned, size_t off, size_t len) { return pread_all(ifd, pinned, len, src_off + off); };
        const int dfd = io.in_direct_fds ? (*io.in_direct_fds)[src.file] : -1;
        if (dfd >= 0) cio.read_shift = [=](void *pinned, size_t cap, size_t off, size_t len) { return pread_direct(dfd, pinned, cap, len, src_off + off); };
        cio.write = [&](const void *pinned, size_t off, size_t len) { return pwrite_all(ofd, pinned, len, out_off + off); };
        const int rc = ggq::cast_chain_io(n.chain.data(), (int)n.chain.size(), count(n.shape), cio);
        if (rc != GGQ_OK) t_cerr = ggq_last_error();
        return rc;
    }
    if (n.kind == Node::CONCAT && axis_is_slowest(n.shape, n.axis)) {  // parts are consecutive byte ranges
        uint64_t at = out_off;
        for (const NodeP &c : n.in) {
            const int rc = emit(*c, at, io);
            if (rc != GGQ_OK) return rc;
            at += nbytes_of(*c);
        }
        return GGQ_OK;
    }
    ggq::Resident *res = io.resident();
    if (!res) return GGQ_ERR_CUDA;
    io.used_resident = true;
    const uint64_t peak = nbytes + resident_peak_bytes(n);
    if (io.budget) io.budget->acquire(io.device, peak);
    void *d = nullptr;
    int rc = res->alloc(nbytes, &d);
    if (rc == GGQ_OK) rc = eval_into(n, d, io);
    if (rc == GGQ_OK) rc = res->download(d, nbytes, [ofd, out_off](const void *pinned, size_t off, size_t len) { return pwrite_all(ofd, pinned, len, out_off + off); });
    if (d) res->free(d);
    if (io.budget) io.budget->release(io.device, peak);
    if (rc != GGQ_OK && t_cerr.empty()) t_cerr = ggq_last_error();
    return rc;
}
bool needs_gpu(const Node &n) {
    if (n.kind == Node::SOURCE) return false;
    if (n.kind == Node::CONCAT && axis_is_slowest(n.shape, n.axis)) {
        for (const NodeP &c : n.in) if (needs_gpu(*c)) return true;
        return false;
    }
    return true;
}

}  // namespace

extern "C" {

const char *ggq_convert_last_error(void) { return t_cerr.c_str(); }

int ggq_convert_gguf_ex(const char *const *in_paths, size_t n_in, const char *out_path, const char *steps,
                        const struct ggq_convert_options *opts, struct ggq_convert_stats *stats) {
    auto failc = [&](int code, const std::string &m) { t_cerr = m; return code; };
    try {
        const double t0 = now();
        ggq_convert_options o{};
        if (opts) o = *opts;
        const uint64_t max_tensors = o.max_tensors ? o.max_tensors : UINT64_MAX;
        const uint64_t max_bytes = o.max_bytes ? o.max_bytes : UINT64_MAX;
        if (!in_paths || n_in == 0 || !out_path) return failc(GGQ_ERR_INVALID, "no input or output path");
        // ---- parse `--steps "a -> b -> c"` (convert.rs:38-51) ----
        std::vector<Step> ops;
        {
            std::string s(steps ? steps : "");
            size_t pos = 0;
            while (pos <= s.size()) {
                size_t nx = s.find("->", pos);
                std::string step = s.substr(pos, nx == s.npos ? s.npos : nx - pos);
                size_t a = step.find_first_not_of(" \t"), b = step.find_last_not_of(" \t");
                step = a == step.npos ? "" : step.substr(a, b - a + 1);
                if (!step.empty()) {
                    Step st{};
                    if (step == "merge-linear") st.kind = Step::MERGE;
                    else if (step == "split-linear" || step == "!merge-linear") st.kind = Step::SPLIT;
                    else if (step == "permute-qk") st.kind = Step::PERMUTE;
                    else if (step.rfind("cast:", 0) == 0) {
                        st.kind = Step::CAST;
                        std::string err;
                        if (!parse_cast_step(step.substr(5), &st.rule, &err)) return failc(GGQ_ERR_UNSUPPORTED, err);
                    } else
                        return failc(GGQ_ERR_UNSUPPORTED, "only `cast:`, `merge-linear`, `split-linear` and `permute-qk` steps are implemented (got '" + step + "')");
                    ops.push_back(st);
                }
                if (nx == s.npos) break;
                pos = nx + 2;
            }
        }
        // ---- map + parse every input shard and merge them (utils/mod.rs:42-46, read.rs:5-62) ----
        std::vector<std::unique_ptr<InFile>> files;
        int n_direct = 0;
        uint64_t alignment = 0, bytes_in = 0;
        std::vector<const gguf::MetaKV *> kvs;
        std::set<std::string_view> kv_seen, name_seen;
        std::vector<Tensor> tensors;
        for (size_t i = 0; i < n_in; i++) {
            auto in = std::make_unique<InFile>();
            in->fd = open(in_paths[i], O_RDONLY);
            if (in->fd < 0) return failc(GGQ_ERR_INVALID, std::string("cannot open ") + in_paths[i]);
            struct stat st;
            if (fstat(in->fd, &st) != 0) return failc(GGQ_ERR_INVALID, std::string("cannot stat ") + in_paths[i]);
            in->map.len = (size_t)st.st_size;
            in->size = (uint64_t)st.st_size;
            if (o.direct_io) {
                in->direct_fd = open(in_paths[i], O_RDONLY | O_DIRECT);  // -1 (EINVAL) where the file system has no O_DIRECT: buffered reads then
                n_direct += in->direct_fd >= 0;
            }
            in->map.p = (uint8_t *)mmap(nullptr, in->map.len, PROT_READ, MAP_PRIVATE, in->fd, 0);
            if (in->map.p == MAP_FAILED) { in->map.p = nullptr; return failc(GGQ_ERR_INVALID, "mmap of the input failed"); }
            in->f = gguf::File::parse(in->map.p, in->map.len);
            in->data_off = (uint64_t)(in->f.data - in->map.p);
            bytes_in += in->map.len;
            alignment = std::max(alignment, in->f.alignment);                          // read.rs:34
            for (const auto &kv : in->f.meta_kvs) {
                if (kv.key == gguf::GENERAL_ALIGNMENT || kv.key.substr(0, 6) == "split.") continue;  // read.rs:37-39
                if (!kv_seen.insert(kv.key).second) return failc(GGQ_ERR_INVALID, "DuplicateMetaKey(" + std::string(kv.key) + ")");
                kvs.push_back(&kv);
            }
            for (const auto &t : in->f.tensors) {
                if (!name_seen.insert(t.name).second) return failc(GGQ_ERR_INVALID, "DuplicateTensorName(" + std::string(t.name) + ")");
                auto node = std::make_shared<Node>();
                node->type = t.type;
                node->shape = t.shape;
                node->file = (int)i;
                node->file_off = in->data_off + t.offset;
                Tensor vt;
                vt.name = std::string(t.name);
                vt.node = node;
                tensors.push_back(std::move(vt));
            }
            files.push_back(std::move(in));
        }
        std::string arch;
        for (const auto &f : files) if (arch.empty()) arch = std::string(f->f.get_str("general.architecture"));

        // ---- apply the operators in order (utils/mod.rs:48-53): they only rewrite the expressions ----
        try {
            for (const Step &op : ops) {
                switch (op.kind) {
                    case Step::CAST:
                        if (arch != "llama" && arch != "gpt2" && arch != "qwen2" && arch != "clip")
                            throw StepError{GGQ_ERR_UNSUPPORTED, "Unsupported architecture: " + arch};  // cast.rs:69
                        for (Tensor &t : tensors) {  // cast.rs:73-90
                            const int cls = classify(arch, t.name, t.node->shape.size());
                            if (op.rule.has[cls]) t.node = make_cast(t.node, op.rule.ty[cls], t.name);
                        }
                        break;
                    case Step::MERGE: apply_merge(tensors); break;
                    case Step::SPLIT:
                    case Step::PERMUTE: {
                        uint64_t nh, nkvh;
                        head_counts(kvs, arch, &nh, &nkvh);
                        if (op.kind == Step::SPLIT) apply_split(tensors, nh, nkvh);
                        else apply_permute(tensors, nh, nkvh);
                        break;
                    }
                }
            }
        } catch (const StepError &e) {
            return failc(e.code, e.msg);
        }
        const size_t nt = tensors.size();
        for (Tensor &t : tensors) t.out_nbytes = nbytes_of(*t.node);
        // ---- plan the output shards (write.rs:23-51, with the simulator's byte accounting) ----
        uint64_t kv_bytes = 0;
        for (const auto *kv : kvs) kv_bytes += kv->raw_len;
        std::vector<std::vector<size_t>> shards(1);
        {
            Simulator sim(alignment, kv_bytes);
            for (size_t i = 0; i < nt; i++) {
                if (shards.size() == 1 && o.no_tensor_first) {
                    sim = Simulator(alignment, 0);
                    sim.write_tensor(tensors[i]);
                    shards.push_back({i});
                    continue;
                }
                sim.write_tensor(tensors[i]);
                if (shards.back().size() < max_tensors && sim.written_bytes() < max_bytes) {
                    shards.back().push_back(i);
                } else {
                    sim = Simulator(alignment, 0);
                    sim.write_tensor(tensors[i]);
                    shards.push_back({i});
                }
            }
        }
        // ---- create the shard files: name-00001-of-0000N.gguf when N > 1 (ggus/src/name/shard.rs:30-39) ----
        std::string base(out_path);
        if (base.size() > 5 && base.substr(base.size() - 5) == ".gguf") base.resize(base.size() - 5);
        struct OutFile { int fd = -1; uint64_t len = 0; ~OutFile() { if (fd >= 0) close(fd); } };
        std::vector<std::unique_ptr<OutFile>> outs;
        uint64_t bytes_out = 0;
        for (size_t si = 0; si < shards.size(); si++) {
            char suffix[64] = "";
            if (shards.size() > 1) snprintf(suffix, sizeof suffix, "-%05zu-of-%05zu", si + 1, shards.size());
            const std::string path = base + suffix + ".gguf";
            std::vector<gguf::OutTensor> ot;
            for (size_t i : shards[si]) ot.push_back({tensors[i].name, &tensors[i].node->shape, tensors[i].node->type, tensors[i].out_nbytes, 0});
            const std::vector<const gguf::MetaKV *> none;
            const auto &shard_kvs = si == 0 ? kvs : none;           // write.rs:72, 77-81
            gguf::Sink sim;
            uint64_t total = gguf::write_front(sim, alignment, shard_kvs, ot);
            const uint64_t front_len = sim.pos();
            if (o.no_data) total = front_len;                       // file_writer.rs:104-106: no data queued
            std::vector<uint8_t> front(front_len);
            gguf::Sink sink(front.data());
            gguf::write_front(sink, alignment, shard_kvs, ot);
            auto of = std::make_unique<OutFile>();
            unlink(path.c_str());
            of->fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
            if (of->fd < 0) return failc(GGQ_ERR_INVALID, "cannot create " + path);
            if (ftruncate(of->fd, (off_t)total) != 0) return failc(GGQ_ERR_INVALID, "ftruncate failed");  // gaps = zero padding
            if (!pwrite_all(of->fd, front.data(), front.size(), 0)) return failc(GGQ_ERR_INVALID, "write of the header failed");
            of->len = total;
            bytes_out += total;
            for (size_t k = 0; k < shards[si].size(); k++) {
                tensors[shards[si][k]].shard = (int)si;
                tensors[shards[si][k]].out_off = ot[k].file_offset;
            }
            outs.push_back(std::move(of));
        }
        const double t1 = now();

        // ---- convert: largest tensors first; WORKERS_PER_DEVICE threads per GPU, each with its own
        // stream pipeline, so one tensor's pread overlaps another's kernels / D2H / pwrite ----
        // File I/O is the bound, not the GPU: 8 workers per GPU saturate one output file; big inputs (where the
        // one-off cost of more pinned pipelines is noise) get up to 16 when the host has the cores, which pays
        // off with sharded output (70B-shaped 31 GB file, -s 1G: 1.56 s with 8, 1.17 s with 16).  GGQ_CONVERT_WORKERS overrides.
        int WORKERS_PER_DEVICE = 8;
        if (bytes_in > (uint64_t(16) << 30)) {
            const int hw = (int)std::thread::hardware_concurrency(), nd = std::max(1, o.n_devices > 0 ? o.n_devices : ggq_device_count());
            WORKERS_PER_DEVICE = std::min(16, std::max(8, hw / nd));
        }
        if (const char *wenv = getenv("GGQ_CONVERT_WORKERS")) { const int v = atoi(wenv); if (v >= 1 && v <= 32) WORKERS_PER_DEVICE = v; }
        std::vector<size_t> order(nt);
        for (size_t i = 0; i < nt; i++) order[i] = i;
        std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return tensors[a].out_nbytes > tensors[b].out_nbytes; });
        std::atomic<size_t> next{0};
        std::atomic<int> rc_all{GGQ_OK};
        std::atomic<uint64_t> cast_elems{0}, cast_tensors{0}, rearranged{0};
        std::atomic<uint64_t> read_ns{0}, write_ns{0}, gpu_wait_ns{0}, h2d_bytes{0}, d2h_bytes{0};
        std::string first_err;
        std::mutex err_mu;
        const int ndev_avail = ggq_device_count();
        int ndev = o.n_devices <= 0 ? ndev_avail : std::min(o.n_devices, ndev_avail);
        bool need_gpu = false;
        for (const auto &t : tensors) need_gpu |= needs_gpu(*t.node);
        if (o.no_data) need_gpu = false;
        if (need_gpu && ndev < 1) return failc(GGQ_ERR_CUDA, "no CUDA device (libggq has no CPU fallback)");
        if (ndev < 1) ndev = 1;
        auto set_err = [&](int rc, const std::string &m) {
            std::lock_guard<std::mutex> lk(err_mu);
            int ok = GGQ_OK;
            if (rc_all.compare_exchange_strong(ok, rc)) first_err = m;
        };
        std::vector<int> in_fds, in_direct_fds;
        for (const auto &f : files) { in_fds.push_back(f->fd); in_direct_fds.push_back(f->direct_fd); }
        ResidentBudget budget;
        auto worker_body = [&](int dev) {
            if (need_gpu && ggq_set_device(dev) != GGQ_OK) { set_err(GGQ_ERR_CUDA, ggq_last_error()); return; }
            std::vector<uint8_t> copy_buf;
            std::unique_ptr<ggq::Resident> res;
            IoCtx io;
            io.in_fds = &in_fds;
            io.in_direct_fds = n_direct ? &in_direct_fds : nullptr;
            io.out_fd = -1;
            io.copy_buf = &copy_buf;
            io.budget = &budget;
            io.device = dev;
            io.resident = [&]() -> ggq::Resident * {
                if (!res) {
                    res = std::make_unique<ggq::Resident>();
                    if (res->status() != GGQ_OK) { t_cerr = ggq_last_error(); res.reset(); }
                }
                return res.get();
            };
            for (;;) {
                const size_t k = next.fetch_add(1);
                if (k >= nt || rc_all.load() != GGQ_OK) return;
                const Tensor &t = tensors[order[k]];
                io.out_fd = outs[t.shard]->fd;
                t_cerr.clear();
                io.used_resident = false;
                const int rc = emit(*t.node, t.out_off, io);
                if (rc != GGQ_OK) { set_err(rc, t.name + ": " + t_cerr); return; }
                const uint64_t ce = cast_elems_of(*t.node);
                if (ce) { cast_elems += ce; cast_tensors += 1; }
                if (io.used_resident) rearranged += 1;
            }
        };
        // Every worker, the first included, is a spawned thread: ggq_set_device() pins the thread it is called on,
        // and the caller's thread must come back from this call as it went in (device, sharding state).  The layout
        // helpers and allocations below emit() can throw; an exception must become a status, not std::terminate.
        auto worker = [&](int dev) {
            t_read_ns = t_write_ns = 0;
            ggq::take_pipe_counters();
            struct Flush {  // whatever way the worker leaves, its counters reach the totals
                std::atomic<uint64_t> &r, &w, &g, &h, &d;
                ~Flush() {
                    const ggq::PipeCounters c = ggq::take_pipe_counters();
                    r += t_read_ns; w += t_write_ns; g += c.gpu_wait_ns; h += c.h2d_bytes; d += c.d2h_bytes;
                }
            } flush{read_ns, write_ns, gpu_wait_ns, h2d_bytes, d2h_bytes};
            try {
                worker_body(dev);
            } catch (const std::exception &e) {
                set_err(GGQ_ERR_INVALID, e.what());
            } catch (...) {
                set_err(GGQ_ERR_INVALID, "unknown exception in a convert worker");
            }
        };
        const int nworkers = ndev * WORKERS_PER_DEVICE;  // plain copies are file I/O: they want the threads too
        if (!o.no_data) {
            std::vector<std::thread> th;
            for (int w = 0; w < nworkers; w++) th.emplace_back(worker, w % ndev);
            for (auto &x : th) x.join();
            if (rc_all.load() != GGQ_OK) return failc(rc_all.load(), first_err);
        }
        const double t2 = now();
        if (stats) {
            stats->n_tensors = nt;
            stats->n_cast_tensors = cast_tensors.load();
            stats->cast_elems = cast_elems.load();
            stats->bytes_in = bytes_in;
            stats->bytes_out = bytes_out;
            stats->seconds_plan = t1 - t0;
            stats->seconds_convert = t2 - t1;
            stats->seconds_sync = 0.0;  // like the reference writer, no fsync: the page cache owns the rest
            stats->n_devices = ndev;
            stats->n_out_files = (int)outs.size();
            stats->n_rearranged_tensors = rearranged.load();
            stats->n_workers = o.no_data ? 0 : nworkers;
            stats->worker_seconds_read = read_ns.load() * 1e-9;
            stats->worker_seconds_write = write_ns.load() * 1e-9;
            stats->worker_seconds_gpu_wait = gpu_wait_ns.load() * 1e-9;
            stats->h2d_bytes = h2d_bytes.load();
            stats->d2h_bytes = d2h_bytes.load();
            stats->n_direct_inputs = n_direct;
        }
        return GGQ_OK;
    } catch (const std::exception &e) {
        return failc(GGQ_ERR_INVALID, e.what());
    }
}

int ggq_convert_gguf(const char *in_path, const char *out_path, const char *steps, int n_devices, struct ggq_convert_stats *stats) {
    ggq_convert_options o{};
    o.n_devices = n_devices;
    const char *ins[1] = {in_path};
    return ggq_convert_gguf_ex(ins, in_path ? 1 : 0, out_path, steps, &o, stats);
}

}  // extern "C"
