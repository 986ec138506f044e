// CPU unit test of gguf_b200/host/tensor_ops.hpp: the planning of merge-linear / split-linear /
// permute-qk / cast (xtask/src/utils/operator/{merge,permute_qk,cast}.rs) — names, order, shapes and
// the normal forms convert.cpp relies on.  No GPU, no files.
#include <cstdio>
#include <string>
#include <vector>

#include "../../gguf_b200/host/tensor_ops.hpp"

using namespace tensor_ops;

static int failures = 0;
#define CHECK(...) do { if (!(__VA_ARGS__)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #__VA_ARGS__); failures++; } } while (0)

static Tensor src(const std::string &name, uint32_t type, std::vector<uint64_t> shape, uint64_t off) {
    auto n = std::make_shared<Node>();
    n->type = type;
    n->shape = std::move(shape);
    n->file = 0;
    n->file_off = off;
    Tensor t;
    t.name = name;
    t.node = n;
    return t;
}

static std::vector<std::string> names(const std::vector<Tensor> &ts) {
    std::vector<std::string> v;
    for (const auto &t : ts) v.push_back(t.name);
    return v;
}

int main() {
    // ---- merge-linear: grouping, naming, "stands where its last part stood", leftovers at the end ----
    std::vector<Tensor> ts;
    ts.push_back(src("token_embd.weight", GGQ_F16, {64, 100}, 0));
    ts.push_back(src("blk.0.attn_q.weight", GGQ_F16, {64, 64}, 1000));
    ts.push_back(src("blk.0.attn_k.weight", GGQ_F16, {64, 16}, 2000));
    ts.push_back(src("blk.0.ffn_gate.weight", GGQ_F16, {64, 96}, 3000));
    ts.push_back(src("blk.0.attn_q.bias", GGQ_F32, {64}, 4000));
    ts.push_back(src("blk.0.attn_v.weight", GGQ_F16, {64, 16}, 5000));
    ts.push_back(src("blk.0.attn_norm.weight", GGQ_F32, {64}, 6000));
    ts.push_back(src("blk.0.ffn_up.weight", GGQ_F16, {64, 96}, 7000));
    ts.push_back(src("blk.1.ffn_gate_exps.weight", GGQ_Q8_0, {64, 32, 4}, 8000));
    ts.push_back(src("blk.1.ffn_up_exps.weight", GGQ_Q8_0, {64, 32, 4}, 9000));
    ts.push_back(src("output.weight", GGQ_F16, {64, 100}, 10000));
    apply_merge(ts);
    CHECK(names(ts) == std::vector<std::string>{"token_embd.weight", "blk.0.attn_qkv.weight", "blk.0.attn_norm.weight", "blk.0.ffn_gate_up.weight",
                                                "blk.1.ffn_gate_up_exps.weight", "output.weight", "blk.0.attn_q.bias"});
    const Node &qkv = *ts[1].node;
    CHECK(qkv.kind == Node::CONCAT && qkv.axis == 1 && qkv.shape == std::vector<uint64_t>{64, 96} && qkv.in.size() == 3);
    CHECK(axis_is_slowest(qkv.shape, qkv.axis));                      // written part by part, no kernel
    const Node &exps = *ts[4].node;
    CHECK(exps.kind == Node::CONCAT && exps.shape == std::vector<uint64_t>{64, 64, 4} && !axis_is_slowest(exps.shape, exps.axis));
    CHECK(nbytes_of(exps) == 64 / 32 * 34 * 64 * 4);

    // ---- cast of a slowest-axis concat is a concat of casts; of the expert concat it stays one cast ----
    NodeP c = make_cast(ts[1].node, GGQ_Q8_0, "qkv");
    CHECK(c->kind == Node::CONCAT && c->type == GGQ_Q8_0 && c->in[0]->kind == Node::CAST && c->in[0]->in[0]->kind == Node::SOURCE);
    NodeP c2 = make_cast(c, GGQ_F32, "qkv");                          // chains extend: F16 -> Q8_0 -> F32
    CHECK(c2->in[1]->chain == std::vector<uint32_t>{GGQ_F16, GGQ_Q8_0, GGQ_F32});
    NodeP ce = make_cast(ts[4].node, GGQ_F16, "exps");
    CHECK(ce->kind == Node::CAST && ce->in[0]->kind == Node::CONCAT);
    CHECK(make_cast(ts[0].node, GGQ_F16, "same") == ts[0].node);      // no-op cast returns the node itself
    try { make_cast(src("x", GGQ_F16, {48, 2}, 0).node, GGQ_Q4K, "x"); CHECK(false); } catch (const StepError &e) { CHECK(e.code == GGQ_ERR_INDIVISIBLE); }

    // ---- split-linear undoes the merge: the parts are the original nodes again ----
    std::vector<Tensor> back = ts;
    apply_split(back, 4, 1);
    CHECK(names(back) == std::vector<std::string>{"token_embd.weight", "blk.0.attn_q.weight", "blk.0.attn_k.weight", "blk.0.attn_v.weight",
                                                  "blk.0.attn_norm.weight", "blk.0.ffn_gate.weight", "blk.0.ffn_up.weight",
                                                  "blk.1.ffn_gate_up_exps.weight", "output.weight", "blk.0.attn_q.bias"});
    CHECK(back[1].node->kind == Node::SOURCE && back[1].node->file_off == 1000 && back[3].node->file_off == 5000);

    // ---- a split of a file tensor is a byte range of the file; through a cast it splits first ----
    Tensor merged = src("blk.2.attn_qkv.weight", GGQ_Q8_0, {64, 96}, 500);   // row = 2 blocks = 68 bytes
    auto parts = split_qkv(merged.node, 4, 1);
    CHECK(parts.size() == 3 && parts[0]->kind == Node::SOURCE && parts[0]->shape == std::vector<uint64_t>{64, 64});
    CHECK(parts[1]->file_off == 500 + 64 * 68 && parts[2]->file_off == 500 + 80 * 68 && nbytes_of(*parts[2]) == 16 * 68);
    NodeP casted = make_cast(merged.node, GGQ_F16, "m");
    auto cparts = split_qkv(casted, 4, 1);
    CHECK(cparts[1]->kind == Node::CAST && cparts[1]->in[0]->kind == Node::SOURCE && cparts[1]->in[0]->file_off == 500 + 64 * 68);
    // 1-D bias in blocks: parts must be whole blocks
    Tensor qb = src("b.attn_qkv.bias", GGQ_Q8_0, {96}, 0);
    try { split_parts(qb.node, 1, {48, 48}); CHECK(false); } catch (const StepError &e) { CHECK(e.code == GGQ_ERR_INDIVISIBLE); }

    // ---- permute-qk: q with nh, k with nkvh, qkv = split / permute / merge; errors instead of panics ----
    std::vector<Tensor> pq;
    pq.push_back(src("blk.0.attn_q.weight", GGQ_F16, {64, 64}, 0));
    pq.push_back(src("blk.0.attn_k.bias", GGQ_F32, {16}, 0));
    pq.push_back(src("blk.0.attn_qkv.weight", GGQ_F16, {64, 96}, 100));
    pq.push_back(src("blk.0.attn_v.weight", GGQ_F16, {64, 16}, 0));
    apply_permute(pq, 4, 1);
    CHECK(pq[0].node->kind == Node::PERMUTE && pq[0].node->nh == 4 && pq[1].node->kind == Node::PERMUTE && pq[1].node->nh == 1);
    CHECK(pq[3].node->kind == Node::SOURCE);
    const Node &pm = *pq[2].node;
    CHECK(pm.kind == Node::CONCAT && pm.in[0]->kind == Node::PERMUTE && pm.in[1]->kind == Node::PERMUTE && pm.in[2]->kind == Node::SOURCE);
    CHECK(pm.in[0]->in[0]->kind == Node::SOURCE && pm.in[1]->in[0]->file_off == 100 + 64 * 128);
    try { make_permute(src("x", GGQ_F16, {64, 30}, 0).node, 4); CHECK(false); } catch (const StepError &e) { CHECK(e.code == GGQ_ERR_INVALID); }
    try { make_permute(src("x", GGQ_F16, {64, 8, 2}, 0).node, 2); CHECK(false); } catch (const StepError &e) { CHECK(e.code == GGQ_ERR_UNSUPPORTED); }
    try { make_permute(src("x", GGQ_Q8_0, {64}, 0).node, 2); CHECK(false); } catch (const StepError &e) { CHECK(e.code == GGQ_ERR_INDIVISIBLE); }

    // ---- name matching (merge.rs:8-10): suffix match, experts are merged but never split ----
    std::string pre, which, wb;
    CHECK(match_linear("blk.3.ffn_gate_exps.weight", {"ffn_gate_exps", "ffn_gate"}, &pre, &which, &wb) && pre == "blk.3." && which == "ffn_gate_exps");
    CHECK(!match_linear("blk.3.ffn_gate_inp.weight", {"ffn_gate_exps", "ffn_gate"}, &pre, &which, &wb));
    CHECK(!match_linear("blk.3.attn_qkv.weight", {"attn_q", "attn_k"}, &pre, &which, &wb));
    CHECK(!match_linear("blk.3.ffn_gate_up_exps.weight", {"attn_qkv", "ffn_gate_up"}, &pre, &which, &wb));
    CHECK(!match_linear("blk.3.attn_q.scale", {"attn_q"}, &pre, &which, &wb));

    // ---- cast rule grammar and classes (cast.rs:11-71) ----
    CastRule r;
    std::string err;
    CHECK(parse_cast_step("linear:q8_0 embd:F16, norm:f32 bogus:q4k", &r, &err) && r.has[LINEAR] && r.ty[LINEAR] == GGQ_Q8_0 && r.ty[EMBD] == GGQ_F16 && !r.has[ELSE]);
    CHECK(!parse_cast_step("linear:q9_9", &r, &err));
    CHECK(classify("llama", "token_embd.weight", 2) == EMBD && classify("llama", "blk.0.attn_norm.weight", 1) == NORM);
    CHECK(classify("llama", "blk.0.attn_q.bias", 1) == LINEAR && classify("llama", "rope_freqs.weight", 1) == ELSE);
    CHECK(classify("clip", "v.patch_embd.weight", 4) == EMBD && classify("clip", "v.blk.0.ln1.weight", 1) == NORM && classify("clip", "resampler.ln_q.weight", 1) == NORM);

    std::printf(failures ? "%d FAILURES\n" : "all ok%.0d\n", failures);
    return failures ? 1 : 0;
}
