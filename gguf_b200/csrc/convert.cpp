// convert.cpp — whole-file GGUF conversion on the GPUs: the `xtask convert --steps "cast:…->cast:…"`
// path (/root/reference/xtask/src/convert.rs:24-58 → utils/mod.rs:36-59 → operator/cast.rs:28-138 →
// utils/write.rs:6-100), restricted to `cast:` steps, the only operator that touches tensor values.
//
// What changes against the reference pipeline:
//   * all casts of a tensor are one device-resident chain (F16→Q8_0→F32→F16 never returns to the host
//     in between; the reference materialises every intermediate in an anonymous mmap);
//   * tensor offsets of the output are planned up front (the reference's simulator does the same,
//     write.rs:23-51), the output file is mapped once and every tensor's D2H lands at its final place;
//   * tensors are spread over `n_devices` worker threads (largest first), one GPU each, no collective.
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
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ggq.h"
#include "../host/gguf.hpp"
#include "ggq_internal.h"

namespace {

thread_local std::string t_cerr;

std::string upper(std::string s) { for (char &c : s) c = (char)std::toupper((unsigned char)c); return s; }

// cast.rs:179-216 `parse`
bool parse_type(const std::string &name, uint32_t *ty) {
    static const std::map<std::string, uint32_t> M = {
        {"F32", 0}, {"F16", 1}, {"Q4_0", 2}, {"Q4_1", 3}, {"Q5_0", 6}, {"Q5_1", 7}, {"Q8_0", 8}, {"Q8_1", 9}, {"Q2K", 10},
        {"Q3K", 11}, {"Q4K", 12}, {"Q5K", 13}, {"Q6K", 14}, {"Q8K", 15}, {"BF16", 30}};
    auto it = M.find(upper(name));
    if (it == M.end()) return false;
    *ty = it->second;
    return true;
}

bool ends_with(std::string_view s, std::string_view suf) { return s.size() >= suf.size() && s.substr(s.size() - suf.size()) == suf; }

struct CastRule { bool has[4] = {false, false, false, false}; uint32_t ty[4] = {0, 0, 0, 0}; };  // linear, embd, norm, else
enum { LINEAR = 0, EMBD = 1, NORM = 2, ELSE = 3 };

// `Operator::cast("k:v k:v")` — cast.rs:11-26 (regex (\w+):(\w+))
bool parse_cast_step(const std::string &spec, CastRule *r, std::string *err) {
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
int classify(const std::string &arch, std::string_view name, size_t ndim) {
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
    uint64_t data_off = 0;
    ~InFile() { if (fd >= 0) close(fd); }
};

struct Tensor {                 // one tensor of the merged content (read.rs:49-59)
    const gguf::TensorInfo *info;
    int file;                   // index into the input files
    std::vector<uint32_t> chain;
    uint64_t out_nbytes;
    int shard = 0;
    uint64_t out_off = 0;
};

bool pread_all(int fd, void *buf, size_t n, uint64_t off) {
    char *p = static_cast<char *>(buf);
    while (n) {
        ssize_t r = pread(fd, p, n, (off_t)off);
        if (r <= 0) return false;
        p += r; off += (uint64_t)r; n -= (size_t)r;
    }
    return true;
}
bool pwrite_all(int fd, const void *buf, size_t n, uint64_t off) {
    const char *p = static_cast<const char *>(buf);
    while (n) {
        ssize_t r = pwrite(fd, p, n, (off_t)off);
        if (r <= 0) return false;
        p += r; off += (uint64_t)r; n -= (size_t)r;
    }
    return true;
}

// bytes of one tensor-info record (writer.rs:74-86)
uint64_t info_bytes(const gguf::TensorInfo &t) { return 8 + t.name.size() + 4 + 8 * t.shape.size() + 4 + 8; }
constexpr uint64_t ALIGNMENT_KV_BYTES = 8 + 17 + 4 + 4;  // "general.alignment": key string + type + u32

// GGufFileSimulator / GGufTensorSimulator (ggus/src/write/simulator.rs:26-96)
struct Simulator {
    uint64_t alignment, written, n = 0;
    std::vector<uint64_t> data;
    Simulator(uint64_t align, uint64_t kv_bytes) : alignment(align), written(24 + ALIGNMENT_KV_BYTES + kv_bytes) {}
    void write_tensor(const gguf::TensorInfo &t, uint64_t nbytes) { written += info_bytes(t); data.push_back(nbytes); }
    uint64_t written_bytes() const {
        uint64_t total = written;
        for (uint64_t len : data) { total += gguf::pad(total, alignment); total += len; }
        return total;
    }
};

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
        std::vector<CastRule> rules;
        {
            std::string s(steps ? steps : "");
            size_t pos = 0;
            while (pos <= s.size()) {
                size_t nx = s.find("->", pos);
                std::string step = s.substr(pos, nx == s.npos ? s.npos : nx - pos);
                size_t a = step.find_first_not_of(" \t"), b = step.find_last_not_of(" \t");
                step = a == step.npos ? "" : step.substr(a, b - a + 1);
                if (!step.empty()) {
                    if (step.rfind("cast:", 0) != 0) return failc(GGQ_ERR_UNSUPPORTED, "only `cast:` steps are implemented (got '" + step + "')");
                    CastRule r;
                    std::string err;
                    if (!parse_cast_step(step.substr(5), &r, &err)) return failc(GGQ_ERR_UNSUPPORTED, err);
                    rules.push_back(r);
                }
                if (nx == s.npos) break;
                pos = nx + 2;
            }
        }
        // ---- map + parse every input shard and merge them (utils/mod.rs:42-46, read.rs:5-62) ----
        std::vector<std::unique_ptr<InFile>> files;
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
                tensors.push_back(Tensor{&t, (int)i, {}, 0});
            }
            files.push_back(std::move(in));
        }
        std::string arch;
        for (const auto &f : files) if (arch.empty()) arch = std::string(f->f.get_str("general.architecture"));
        if (!rules.empty() && arch != "llama" && arch != "gpt2" && arch != "qwen2" && arch != "clip")
            return failc(GGQ_ERR_UNSUPPORTED, "Unsupported architecture: " + arch);  // cast.rs:69

        // ---- per-tensor type chains (cast.rs:73-90, applied once per step) and output sizes ----
        const size_t nt = tensors.size();
        for (auto &t : tensors) {
            const auto &ti = *t.info;
            t.chain.push_back(ti.type);
            const int cls = classify(arch, ti.name, ti.shape.size());
            for (const CastRule &r : rules)
                if (r.has[cls] && r.ty[cls] != t.chain.back()) t.chain.push_back(r.ty[cls]);
            t.out_nbytes = ti.nbytes;
            if (t.chain.size() > 1) {
                uint64_t be, bb;
                if (!gguf::type_size(t.chain.back(), &be, &bb)) return failc(GGQ_ERR_UNSUPPORTED, "unsupported target type");
                if (ti.shape.empty() || ti.shape[0] % be)  // cast.rs:142-143 `assert_eq!(row % N, 0)`
                    return failc(GGQ_ERR_INDIVISIBLE, "row of " + std::string(ti.name) + " is not a multiple of the target block size");
                t.out_nbytes = ti.n_elems() / be * bb;
                for (uint32_t c : t.chain)
                    if (ggq_type_nbytes(c, ti.n_elems()) == 0) return failc(GGQ_ERR_UNSUPPORTED, "cast chain of " + std::string(ti.name) + " has an unsupported type");
            }
        }
        // ---- plan the output shards (write.rs:23-51, with the simulator's byte accounting) ----
        uint64_t kv_bytes = 0;
        for (const auto *kv : kvs) kv_bytes += kv->raw_len;
        std::vector<std::vector<size_t>> shards(1);
        {
            Simulator sim(alignment, kv_bytes);
            for (size_t i = 0; i < nt; i++) {
                if (shards.size() == 1 && o.no_tensor_first) {
                    sim = Simulator(alignment, 0);
                    sim.write_tensor(*tensors[i].info, tensors[i].out_nbytes);
                    shards.push_back({i});
                    continue;
                }
                sim.write_tensor(*tensors[i].info, tensors[i].out_nbytes);
                if (shards.back().size() < max_tensors && sim.written_bytes() < max_bytes) {
                    shards.back().push_back(i);
                } else {
                    sim = Simulator(alignment, 0);
                    sim.write_tensor(*tensors[i].info, tensors[i].out_nbytes);
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
            for (size_t i : shards[si]) ot.push_back({tensors[i].info->name, &tensors[i].info->shape, tensors[i].chain.back(), tensors[i].out_nbytes, 0});
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
        int WORKERS_PER_DEVICE = 8;  // GGQ_CONVERT_WORKERS overrides (file I/O is the bound, not the GPU)
        if (const char *wenv = getenv("GGQ_CONVERT_WORKERS")) { const int v = atoi(wenv); if (v >= 1 && v <= 32) WORKERS_PER_DEVICE = v; }
        std::vector<size_t> order(nt);
        for (size_t i = 0; i < nt; i++) order[i] = i;
        std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return tensors[a].info->nbytes > tensors[b].info->nbytes; });
        std::atomic<size_t> next{0};
        std::atomic<int> rc_all{GGQ_OK};
        std::atomic<uint64_t> cast_elems{0}, cast_tensors{0};
        std::string first_err;
        std::mutex err_mu;
        const int ndev_avail = ggq_device_count();
        int ndev = o.n_devices <= 0 ? ndev_avail : std::min(o.n_devices, ndev_avail);
        bool need_gpu = false;
        for (const auto &t : tensors) need_gpu |= t.chain.size() > 1;
        if (o.no_data) need_gpu = false;
        if (need_gpu && ndev < 1) return failc(GGQ_ERR_CUDA, "no CUDA device (libggq has no CPU fallback)");
        if (ndev < 1) ndev = 1;
        auto set_err = [&](int rc, const std::string &m) {
            std::lock_guard<std::mutex> lk(err_mu);
            int ok = GGQ_OK;
            if (rc_all.compare_exchange_strong(ok, rc)) first_err = m;
        };
        auto worker = [&](int dev) {
            if (need_gpu && ggq_set_device(dev) != GGQ_OK) { set_err(GGQ_ERR_CUDA, ggq_last_error()); return; }
            std::vector<uint8_t> copy_buf;
            for (;;) {
                const size_t k = next.fetch_add(1);
                if (k >= nt || rc_all.load() != GGQ_OK) return;
                const Tensor &t = tensors[order[k]];
                const auto &ti = *t.info;
                const int ifd = files[t.file]->fd, ofd = outs[t.shard]->fd;
                const uint64_t src_off = files[t.file]->data_off + ti.offset, dst_off = t.out_off;
                if (t.chain.size() == 1) {  // untouched tensor: byte copy
                    constexpr size_t CH = size_t(8) << 20;
                    copy_buf.resize(std::min<uint64_t>(CH, ti.nbytes));
                    for (uint64_t off = 0; off < ti.nbytes; off += CH) {
                        const size_t n = (size_t)std::min<uint64_t>(CH, ti.nbytes - off);
                        if (!pread_all(ifd, copy_buf.data(), n, src_off + off) || !pwrite_all(ofd, copy_buf.data(), n, dst_off + off)) {
                            set_err(GGQ_ERR_INVALID, "I/O error copying " + std::string(ti.name));
                            return;
                        }
                    }
                    continue;
                }
                ggq::ChainIO io;
                io.read = [&](void *pinned, size_t off, size_t n) { return pread_all(ifd, pinned, n, src_off + off); };
                io.write = [&](const void *pinned, size_t off, size_t n) { return pwrite_all(ofd, pinned, n, dst_off + off); };
                const int rc = ggq::cast_chain_io(t.chain.data(), (int)t.chain.size(), ti.n_elems(), io);
                if (rc != GGQ_OK) { set_err(rc, std::string(ti.name) + ": " + ggq_last_error()); return; }
                cast_elems += ti.n_elems();
                cast_tensors += 1;
            }
        };
        if (!o.no_data) {
            std::vector<std::thread> th;
            const int nworkers = ndev * (need_gpu ? WORKERS_PER_DEVICE : 1);
            for (int w = 1; w < nworkers; w++) th.emplace_back(worker, w % ndev);
            worker(0);
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
