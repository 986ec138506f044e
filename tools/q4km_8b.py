"""BASELINE.json configs[2]: Llama-3-8B-shaped F16 -> Q4_K_M mix (Q4_K everywhere, Q6_K for output.weight
and for attn_v / ffn_down on the `use_more_bits` layers — upstream llama.cpp's rule; the reference has no
mix rule), device-resident on one B200.  Reports GB/s per type and a code-level diff against the CPU
oracle on a strided sample of super-blocks, for Gaussian and heavy-tailed (Student-t, nu=3) inputs."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gguf_b200 as g
from oracle import oracle as O

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
st = torch.cuda.current_stream().cuda_stream
def use_more_bits(i, n): return i < n // 8 or i >= 7 * n // 8 or (i - n // 8) % 3 == 2
tensors = [("token_embd.weight", 4096 * 128256, g.Q4K), ("output.weight", 4096 * 128256, g.Q6K)]
for l in range(layers):
    more = use_more_bits(l, layers)
    tensors += [(f"blk.{l}.attn_q", 4096 * 4096, g.Q4K), (f"blk.{l}.attn_k", 4096 * 1024, g.Q4K),
                (f"blk.{l}.attn_v", 4096 * 1024, g.Q6K if more else g.Q4K), (f"blk.{l}.attn_output", 4096 * 4096, g.Q4K),
                (f"blk.{l}.ffn_gate", 4096 * 14336, g.Q4K), (f"blk.{l}.ffn_up", 4096 * 14336, g.Q4K),
                (f"blk.{l}.ffn_down", 14336 * 4096, g.Q6K if more else g.Q4K)]
res = {"config": "Llama-3-8B-shaped F16 -> Q4_K_M mix (BASELINE configs[2])", "layers": layers, "variants": {}}
for variant in ("gaussian", "student_t3"):
    gen = torch.Generator(device="cuda"); gen.manual_seed(2)
    tot = {g.Q4K: [0, 0.0, 0], g.Q6K: [0, 0.0, 0]}   # elems, seconds, bytes
    diff_codes = diff_scales = sampled_blocks = 0
    max_abs_err = 0.0
    for name, n, ty in tensors:
        if variant == "gaussian":
            x = (torch.randn(n, device="cuda", generator=gen) * 0.02).to(torch.float16)
        else:
            z = torch.randn(n, device="cuda", generator=gen)
            chi = (torch.randn(n, device="cuda", generator=gen) ** 2 + torch.randn(n, device="cuda", generator=gen) ** 2 + torch.randn(n, device="cuda", generator=gen) ** 2) / 3
            x = (z / chi.sqrt() * 0.02).clamp(-60000, 60000).to(torch.float16); del z, chi
        e, b = g.block_info(ty)
        packed = torch.empty(n // e * b, dtype=torch.uint8, device="cuda")
        g.quantize_slice_device(ty, g.F16, packed, n // e, x, n, st)   # warm
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.quantize_slice_device(ty, g.F16, packed, n // e, x, n, st); e1.record(); torch.cuda.synchronize()
        tot[ty][0] += n; tot[ty][1] += e0.elapsed_time(e1) * 1e-3; tot[ty][2] += n * 2 + n // e * b
        # parity sample: every 1009th super-block re-quantised by the oracle
        idx = torch.arange(0, n // 256, 1009, device="cuda")
        xs = x.view(-1, 256)[idx].contiguous().cpu().numpy().view(np.uint16).reshape(-1)
        got = packed.view(-1, b)[idx].contiguous().cpu().numpy()
        want = O.quantize(ty, O.F16, xs, threads=os.cpu_count()).reshape(-1, b)
        hdr = slice(0, 16) if ty == g.Q4K else slice(192, 210)
        mask = np.ones(b, bool); mask[hdr] = False
        diff_codes += int((got[:, mask] != want[:, mask]).sum()); diff_scales += int((got[:, hdr] != want[:, hdr]).sum())
        sampled_blocks += len(idx)
        y = g.dequantize(ty, got.reshape(-1), g.F32)
        max_abs_err = max(max_abs_err, float(np.abs(y - xs.view(np.float16).astype(np.float32)).max()))
        del x, packed
    res["variants"][variant] = {
        "Q4_K": {"elements": tot[g.Q4K][0], "seconds": tot[g.Q4K][1], "GBps": tot[g.Q4K][2] / tot[g.Q4K][1] / 1e9},
        "Q6_K": {"elements": tot[g.Q6K][0], "seconds": tot[g.Q6K][1], "GBps": tot[g.Q6K][2] / tot[g.Q6K][1] / 1e9},
        "whole_model_seconds": tot[g.Q4K][1] + tot[g.Q6K][1], "sampled_super_blocks": sampled_blocks,
        "differing_code_bytes": diff_codes, "differing_scale_bytes": diff_scales, "max_abs_reconstruction_error": max_abs_err}
    print(variant, json.dumps(res["variants"][variant]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "q4km_8b.json"), "w"), indent=1)
