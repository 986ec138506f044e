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
#include <mutex>
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

}  // namespace

extern "C" {

const char *ggq_convert_last_error(void) { return t_cerr.c_str(); }

int ggq_convert_gguf(const char *in_path, const char *out_path, const char *steps, int n_devices, struct ggq_convert_stats *stats) {
    auto failc = [&](int code, const std::string &m) { t_cerr = m; return code; };
    try {
        const double t0 = now();
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
        // ---- map + parse the input (utils/mod.rs:42-46, read.rs:5-31) ----
        int fd = open(in_path, O_RDONLY);
        if (fd < 0) return failc(GGQ_ERR_INVALID, std::string("cannot open ") + in_path);
        struct stat st;
        fstat(fd, &st);
        Mapping in;
        in.len = (size_t)st.st_size;
        in.p = (uint8_t *)mmap(nullptr, in.len, PROT_READ, MAP_PRIVATE, fd, 0);
        close(fd);
        if (in.p == MAP_FAILED) { in.p = nullptr; return failc(GGQ_ERR_INVALID, "mmap of the input failed"); }
        gguf::File f = gguf::File::parse(in.p, in.len);
        const std::string arch(f.get_str("general.architecture"));
        if (!rules.empty() && arch != "llama" && arch != "gpt2" && arch != "qwen2" && arch != "clip")
            return failc(GGQ_ERR_UNSUPPORTED, "Unsupported architecture: " + arch);  // cast.rs:69

        // ---- per-tensor type chains (cast.rs:73-90, applied once per step) ----
        const size_t nt = f.tensors.size();
        std::vector<std::vector<uint32_t>> chains(nt);
        for (size_t i = 0; i < nt; i++) {
            chains[i].push_back(f.tensors[i].type);
            const int cls = classify(arch, f.tensors[i].name, f.tensors[i].shape.size());
            for (const CastRule &r : rules)
                if (r.has[cls] && r.ty[cls] != chains[i].back()) chains[i].push_back(r.ty[cls]);
        }
        // ---- plan the output (write.rs:23-51 simulator, single shard) ----
        std::vector<const gguf::MetaKV *> kvs;
        for (const auto &kv : f.meta_kvs)
            if (kv.key != gguf::GENERAL_ALIGNMENT && kv.key.substr(0, 6) != "split.") kvs.push_back(&kv);  // read.rs:37-39
        std::vector<gguf::OutTensor> outs(nt);
        for (size_t i = 0; i < nt; i++) {
            const auto &t = f.tensors[i];
            const uint32_t ty = chains[i].back();
            uint64_t nbytes = t.nbytes;
            if (chains[i].size() > 1) {
                uint64_t be, bb;
                if (!gguf::type_size(ty, &be, &bb)) return failc(GGQ_ERR_UNSUPPORTED, "unsupported target type");
                if (t.shape.empty() || t.shape[0] % be)  // cast.rs:142-143 `assert_eq!(row % N, 0)`
                    return failc(GGQ_ERR_INDIVISIBLE, "row of " + std::string(t.name) + " is not a multiple of the target block size");
                nbytes = t.n_elems() / be * bb;
                for (uint32_t c : chains[i])
                    if (ggq_type_nbytes(c, t.n_elems()) == 0) return failc(GGQ_ERR_UNSUPPORTED, "cast chain of " + std::string(t.name) + " has an unsupported type");
            }
            outs[i] = {t.name, &t.shape, ty, nbytes, 0};
        }
        gguf::Sink sim;
        const uint64_t out_len = gguf::write_front(sim, f.alignment, kvs, outs);
        const uint64_t front_len = sim.pos();

        // ---- output file: front matter with one write, tensor bytes with pwrite at their final
        // offsets; the gaps between tensors are the zero padding of writer.rs:95-100 (ftruncate) ----
        unlink(out_path);
        int ofd = open(out_path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (ofd < 0) return failc(GGQ_ERR_INVALID, std::string("cannot create ") + out_path);
        struct FdGuard { int fd; ~FdGuard() { if (fd >= 0) close(fd); } } og{ofd};
        if (ftruncate(ofd, (off_t)out_len) != 0) return failc(GGQ_ERR_INVALID, "ftruncate failed");
        {
            std::vector<uint8_t> front(front_len);
            gguf::Sink sink(front.data());
            gguf::write_front(sink, f.alignment, kvs, outs);
            if (pwrite(ofd, front.data(), front.size(), 0) != (ssize_t)front.size()) return failc(GGQ_ERR_INVALID, "write of the header failed");
        }
        int ifd = open(in_path, O_RDONLY);
        if (ifd < 0) return failc(GGQ_ERR_INVALID, std::string("cannot reopen ") + in_path);
        FdGuard ig{ifd};
        const uint64_t in_data_off = (uint64_t)(f.data - in.p);
        auto pread_all = [](int fd, void *buf, size_t n, uint64_t off) {
            char *p = static_cast<char *>(buf);
            while (n) {
                ssize_t r = pread(fd, p, n, (off_t)off);
                if (r <= 0) return false;
                p += r; off += (uint64_t)r; n -= (size_t)r;
            }
            return true;
        };
        auto pwrite_all = [](int fd, const void *buf, size_t n, uint64_t off) {
            const char *p = static_cast<const char *>(buf);
            while (n) {
                ssize_t r = pwrite(fd, p, n, (off_t)off);
                if (r <= 0) return false;
                p += r; off += (uint64_t)r; n -= (size_t)r;
            }
            return true;
        };
        const double t1 = now();

        // ---- convert: largest tensors first; WORKERS_PER_DEVICE threads per GPU, each with its own
        // stream pipeline, so one tensor's pread overlaps another's kernels / D2H / pwrite ----
        int WORKERS_PER_DEVICE = 4;  // GGQ_CONVERT_WORKERS overrides (file I/O is the bound, not the GPU)
        if (const char *wenv = getenv("GGQ_CONVERT_WORKERS")) { const int v = atoi(wenv); if (v >= 1 && v <= 32) WORKERS_PER_DEVICE = v; }
        std::vector<size_t> order(nt);
        for (size_t i = 0; i < nt; i++) order[i] = i;
        std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return f.tensors[a].nbytes > f.tensors[b].nbytes; });
        std::atomic<size_t> next{0};
        std::atomic<int> rc_all{GGQ_OK};
        std::atomic<uint64_t> cast_elems{0}, cast_tensors{0};
        std::string first_err;
        std::mutex err_mu;
        const int ndev_avail = ggq_device_count();
        int ndev = n_devices <= 0 ? ndev_avail : std::min(n_devices, ndev_avail);
        bool need_gpu = false;
        for (size_t i = 0; i < nt; i++) need_gpu |= chains[i].size() > 1;
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
                const size_t i = order[k];
                const auto &t = f.tensors[i];
                const uint64_t src_off = in_data_off + t.offset, dst_off = outs[i].file_offset;
                if (chains[i].size() == 1) {  // untouched tensor: byte copy
                    constexpr size_t CH = size_t(8) << 20;
                    copy_buf.resize(std::min<uint64_t>(CH, t.nbytes));
                    for (uint64_t o = 0; o < t.nbytes; o += CH) {
                        const size_t n = (size_t)std::min<uint64_t>(CH, t.nbytes - o);
                        if (!pread_all(ifd, copy_buf.data(), n, src_off + o) || !pwrite_all(ofd, copy_buf.data(), n, dst_off + o)) {
                            set_err(GGQ_ERR_INVALID, "I/O error copying " + std::string(t.name));
                            return;
                        }
                    }
                    continue;
                }
                ggq::ChainIO io;
                io.read = [&](void *pinned, size_t off, size_t n) { return pread_all(ifd, pinned, n, src_off + off); };
                io.write = [&](const void *pinned, size_t off, size_t n) { return pwrite_all(ofd, pinned, n, dst_off + off); };
                const int rc = ggq::cast_chain_io(chains[i].data(), (int)chains[i].size(), t.n_elems(), io);
                if (rc != GGQ_OK) { set_err(rc, std::string(t.name) + ": " + ggq_last_error()); return; }
                cast_elems += t.n_elems();
                cast_tensors += 1;
            }
        };
        std::vector<std::thread> th;
        const int nworkers = ndev * (need_gpu ? WORKERS_PER_DEVICE : 1);
        for (int w = 1; w < nworkers; w++) th.emplace_back(worker, w % ndev);
        worker(0);
        for (auto &x : th) x.join();
        if (rc_all.load() != GGQ_OK) return failc(rc_all.load(), first_err);
        const double t2 = now();
        const double t3 = t2;  // like the reference writer, no fsync: the page cache owns the rest
        if (stats) {
            stats->n_tensors = nt;
            stats->n_cast_tensors = cast_tensors.load();
            stats->cast_elems = cast_elems.load();
            stats->bytes_in = in.len;
            stats->bytes_out = out_len;
            stats->seconds_plan = t1 - t0;
            stats->seconds_convert = t2 - t1;
            stats->seconds_sync = t3 - t2;
            stats->n_devices = ndev;
        }
        return GGQ_OK;
    } catch (const std::exception &e) {
        return failc(GGQ_ERR_INVALID, e.what());
    }
}

}  // extern "C"
