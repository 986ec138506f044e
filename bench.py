#!/usr/bin/env python
"""bench.py — headline benchmark of the ggml block codec hot path on B200 (see BASELINE.json).

Workload (N=1): BASELINE.json configs[1] — Llama-3-8B-shaped synthetic tensors (4096x14336 FFN,
4096x4096 attention), dequantize Q4_0 / Q8_0 / Q4_K / Q6_K -> f16.  One *step* = one pass of
`dequantize_slice` over all 8 (type, shape) tensors.  Metric: algorithmic GB/s
(bytes = packed bytes read + f16 bytes written), whole job.

  value     device-resident: inputs already in HBM, kernels launched through the C ABI
            (`ggq_dequantize_slice_device`) on torch's current stream, CUDA-event timed.
  e2e       the same step through the host C ABI (`ggq_dequantize_slice`, the drop-in for
            QuantExt::dequantize_slice) with pinned HOST buffers: H2D + kernel + D2H inside the timing.
  roofline  the dominant kernel (largest share of the step): algorithmic bytes per launch / average
            launch duration from CUDA events recorded around that launch inside the timed region.
  cpu_baseline  the CPU oracle port (oracle/, the reference's algorithm restated in C; the Rust
            reference cannot be built in this image) on the host cores, bounded sample.

`--impl reference` times that CPU port on the same workload config (bounded sample per step).
Multi-GPU (`torchrun ... bench.py --gpus N`): tensors are independent, each rank processes its own
copy of the workload on its own GPU, no collective on the data path => "scaling": "weak".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

F32, F16, BF16 = 0, 1, 30
Q4_0, Q8_0, Q4K, Q6K = 2, 8, 12, 14
NAMES = {2: "Q4_0", 8: "Q8_0", 12: "Q4_K", 14: "Q6_K"}
BLOCK = {2: (32, 18), 8: (32, 34), 12: (256, 144), 14: (256, 210)}
SHAPES = {"ffn": (4096, 14336), "attn": (4096, 4096)}
TYPES = [Q4_0, Q8_0, Q4K, Q6K]
WORKLOAD = "llama3-8b-shaped dequant Q4_0/Q8_0/Q4_K/Q6_K->f16: 4096x14336 + 4096x4096 per type (BASELINE configs[1])"
METRIC = "dequant_GBps_algorithmic"
UNIT = "GB/s"


def algo_bytes(ty, n_elems):
    e, b = BLOCK[ty]
    return n_elems // e * b + n_elems * 2


def tensors():
    out = []
    for ty in TYPES:
        for sname, (r, c) in SHAPES.items():
            out.append((ty, sname, r * c))
    return out


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa(local_rank):
    """Pin this rank to the CPUs nearest its GPU (NVML affinity) BEFORE any pinned allocation, so the
    staging buffers are first-touched on the local NUMA node.  Best effort; returns a description."""
    if os.environ.get("GGQ_BENCH_NO_BIND"):
        return "disabled"
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [i * 64 + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            return f"{len(allowed)} cpus [{allowed[0]}..{allowed[-1]}]"
        return "no narrower affinity reported"
    except Exception as e:  # noqa: BLE001
        return f"unavailable ({type(e).__name__})"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
def run_cpu_port(steps, warmup, sample_elems, threads):
    """The oracle port on host cores: dequantize `sample_elems` elements of every type per step."""
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    packed = {}
    for ty in TYPES:
        e, b = BLOCK[ty]
        x = (rng.standard_normal(sample_elems) * 0.02).astype(np.float32).astype(np.float16)
        packed[ty] = O.quantize(ty, O.F16, x, threads=threads)
    outs = {ty: np.empty(sample_elems, np.uint16) for ty in TYPES}
    L = O.lib()

    def step():
        for ty in TYPES:
            e, b = BLOCK[ty]
            rc = L.ggo_dequantize_slice(ty, O.F16, outs[ty].ctypes.data, sample_elems, packed[ty].ctypes.data, sample_elems // e, threads)
            assert rc == 0
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    nbytes = sum(algo_bytes(ty, sample_elems) for ty in TYPES)
    return nbytes * steps / dt / 1e9, dt / steps * 1e3


def reference_arm(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 4096 * 4096
    gbs, ms = run_cpu_port(args.steps, max(args.warmup, 1), sample, threads)
    sample_desc = f"4096x4096 elements per type x {len(TYPES)} types per step (attention-shaped slice of the workload), {threads} pthreads over contiguous block ranges"
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/f32->f16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU port of the reference algorithm (oracle/ggq_oracle.c); the Rust reference cannot be compiled here (no cargo/rustc)"},
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_desc},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import gguf_b200 as g
    from gguf_b200._lib import lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(local_rank) if world > 1 else "single rank"
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its "NCCL version ..." banner there at NCCL_DEBUG=VERSION and
        # WARN (NCCL_DEBUG_FILE does not move it), so those two levels are switched off; INFO / TRACE are left to a
        # user who asked for them.  NCCL is only this bench's barrier and max-reduce of the timing, not the data path.
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"
        dist.init_process_group("nccl", device_id=dev)
    L = lib()
    stream = torch.cuda.current_stream().cuda_stream

    # ---- synthetic tensors: Gaussian f16 weights, quantised ON THE GPU by this library ----
    work = []
    gen = torch.Generator(device=dev)
    gen.manual_seed(1 + rank)
    for ty, sname, n in tensors():
        e, b = BLOCK[ty]
        x = (torch.randn(n, device=dev, generator=gen, dtype=torch.float32) * 0.02).to(torch.float16)
        packed = torch.empty(n // e * b, dtype=torch.uint8, device=dev)
        g.quantize_slice_device(ty, F16, packed, n // e, x, n, stream)
        out = torch.empty(n, dtype=torch.float16, device=dev)
        work.append({"ty": ty, "shape": sname, "n": n, "nb": n // e, "packed": packed, "out": out, "bytes": algo_bytes(ty, n)})
        del x
    torch.cuda.synchronize()
    step_bytes = sum(w["bytes"] for w in work)

    def step():
        for w in work:
            g.dequantize_slice_device(w["ty"], F16, w["out"], w["n"], w["packed"], w["nb"], stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    # ---- timed region 1: K steps back to back, events only at the ends -> `value` ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = L.ggq_launch_count()
    barrier()
    t_start.record()
    for k in range(args.steps):
        step()
    t_end.record()
    barrier()
    launches = L.ggq_launch_count() - launches0
    elapsed_ms = t_start.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = step_bytes * args.steps * world / (elapsed_ms * 1e-3) / 1e9

    # ---- timed region 2: per-kernel average launch duration.  R back-to-back launches of ONE kernel
    # between two CUDA events (an event pair around every single launch adds ~3.7 us of drain +
    # timestamp per launch and was measured to under-report these 8-30 us kernels by 10-30 %), rotating
    # over NSETS distinct (packed, out) buffer sets so the footprint exceeds L2. ----
    per_kernel = []
    NSETS, R = 5, max(10, min(args.steps, 50))
    for w in work:
        sets = [(w["packed"], w["out"])] + [(w["packed"].clone(), torch.empty_like(w["out"])) for _ in range(NSETS - 1)]
        for i in range(NSETS):
            g.dequantize_slice_device(w["ty"], F16, sets[i][1], w["n"], sets[i][0], w["nb"], stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(R):
            pk, out = sets[i % NSETS]
            g.dequantize_slice_device(w["ty"], F16, out, w["n"], pk, w["nb"], stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / R
        launches += R + NSETS
        per_kernel.append({"kernel": f"dequant_kernel<{NAMES[w['ty']]},f16>", "shape": w["shape"], "us": ms * 1e3,
                           "GBps": w["bytes"] / (ms * 1e-3) / 1e9, "bytes": w["bytes"], "launches_timed": R})
        del sets
    # ---- informational: the quantize direction of the same four types on the FFN shape (f16 -> packed), same method.
    # BASELINE.json's metric names both directions; `value` and `roofline` stay the dequantize workload's. ----
    quant_per_kernel = []
    if rank == 0:
        ffn = [w for w in work if w["shape"] == "ffn"]
        xs = [(torch.randn(ffn[0]["n"], device=dev) * 0.02).to(torch.float16) for _ in range(3)]
        for w in ffn:
            outs = [torch.empty_like(w["packed"]) for _ in range(3)]
            rq = 6 if w["ty"] in (g.Q4K, g.Q6K) else R
            for i in range(3):
                g.quantize_slice_device(w["ty"], F16, outs[i], w["nb"], xs[i], w["n"], stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for i in range(rq):
                g.quantize_slice_device(w["ty"], F16, outs[i % 3], w["nb"], xs[i % 3], w["n"], stream)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / rq
            launches += rq + 3
            quant_per_kernel.append({"kernel": f"quantize<{NAMES[w['ty']]},f16>", "shape": "ffn", "us": ms * 1e3,
                                     "GBps": w["bytes"] / (ms * 1e-3) / 1e9, "bytes": w["bytes"], "launches_timed": rq,
                                     "bound": "fp32 issue (bit-faithful scale search)" if w["ty"] in (g.Q4K, g.Q6K) else "hbm"})
            del outs
        del xs
    clocks = sampler.stop()
    dom = max(per_kernel, key=lambda r: r["us"])
    peak, peak_src = measured_peak()
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        dom_ty = [w["ty"] for w in work if f"dequant_kernel<{NAMES[w['ty']]},f16>" == dom["kernel"]][0]
        if dom["shape"] == "ffn":
            traffic = tj["dequant_ffn_f16"].get(str(dom_ty))
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": f"{dom['kernel']} {dom['shape']}", "achieved": dom["GBps"], "peak": peak, "unit": "GB/s",
                "frac": dom["GBps"] / peak, "traffic": traffic, "peak_source": peak_src, "frac_of_nominal_8000_GBps": dom["GBps"] / 8000.0, "algorithmic_bytes_per_launch": dom["bytes"],
                "avg_launch_us": dom["us"], "method": f"{R} back-to-back launches between two CUDA events on the launch stream, {NSETS} rotating buffer sets"}

    # ---- e2e: host C ABI with pinned host buffers (H2D + kernel + D2H in the timing) ----
    e2e = None
    if not args.no_e2e:
        host = []
        for w in work:
            pin_in, pin_out = g.PinnedBuffer(w["packed"].numel()), g.PinnedBuffer(w["n"] * 2)
            pin_in.array[:] = w["packed"].cpu().numpy()
            host.append((w, pin_in, pin_out))

        from gguf_b200._lib import SliceJob
        jobs = (SliceJob * len(host))()
        for i, (w, pi, po) in enumerate(host):
            jobs[i] = SliceJob(w["ty"], F16, 0, po.ptr, w["n"], pi.ptr, w["nb"])

        def host_step():          # ONE ggq_slices call per step: the 8 tensors stream through one pipeline
            rc = L.ggq_slices(jobs, len(host))
            assert rc == 0, L.ggq_last_error()

        def host_step_per_call():  # the same step as 8 separate synchronous slice calls
            for w, pi, po in host:
                rc = L.ggq_dequantize_slice(w["ty"], F16, po.ptr, w["n"], pi.ptr, w["nb"])
                assert rc == 0, L.ggq_last_error()
        for _ in range(max(1, min(args.warmup, 2))):
            host_step()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        t1 = time.perf_counter()
        for _ in range(e2e_steps):
            host_step_per_call()
        dt_per_call = time.perf_counter() - t1
        if world > 1:
            t = torch.tensor([dt, dt_per_call], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, dt_per_call = float(t[0].item()), float(t[1].item())
        # spot check: the host path and the device path agree byte for byte
        w, pi, po = host[0]
        assert np.array_equal(po.view(np.uint16), w["out"].cpu().numpy().view(np.uint16)), "host/device path mismatch"
        e2e = {"value": step_bytes * e2e_steps * world / dt / 1e9, "unit": UNIT, "steps": e2e_steps,
               "h2d_bytes_per_step": int(sum(w["packed"].numel() for w in work)), "d2h_bytes_per_step": int(sum(w["n"] * 2 for w in work)),
               "api": "ggq_slices: one synchronous call per step over the 8 tensors (host pointers, pinned)", "ms_per_step": dt / e2e_steps * 1e3,
               "per_call_value": step_bytes * e2e_steps * world / dt_per_call / 1e9,
               "per_call_api": "8 separate ggq_dequantize_slice calls per step (pipeline drains between tensors)"}
        for _, pi, po in host:
            pi.free(); po.free()

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on a bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        sample = 4096 * 4096
        gbs, _ = run_cpu_port(3, 1, sample, threads)
        cpu = {"value": gbs, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"4096x4096 elements per type x {len(TYPES)} types, 3 passes, oracle/ggq_oracle.c with {threads} pthreads"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/f32->f16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "bytes_per_step_per_gpu": step_bytes,
                       "l2": "per-step footprint (packed+f16 of 8 tensors = %.0f MB) exceeds the 126 MB L2; no flush needed" % (step_bytes / 1e6),
                       "sharding": "by tensor, one replica of the workload per GPU, no collective", "cpu_binding_rank0": numa},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "frac_of_peak_whole_step": value / world / peak, "per_kernel": per_kernel, "quant_per_kernel": quant_per_kernel,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:])
    ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
