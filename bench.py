#!/usr/bin/env python
"""bench.py — headline benchmark of the ggml block codec hot path on B200 (see BASELINE.json).

Workload (N=1): BASELINE.json configs[1] — Llama-3-8B-shaped synthetic tensors (4096x14336 FFN,
4096x4096 attention), dequantize Q4_0 / Q8_0 / Q4_K / Q6_K -> f16.  One *step* = one pass of
`dequantize_slice` over all 8 (type, shape) tensors.  Metric: algorithmic GB/s
(bytes = packed bytes read + f16 bytes written), whole job.

  value     device-resident: inputs already in HBM; the step is ONE `ggq_slices_device` call (the eight
            dequantize_slice jobs share one descriptor-table grid) on torch's current stream, CUDA-event timed.
  e2e       the same step through the host C ABI (`ggq_slices`, the drop-in for a writer thread's
            QuantExt::dequantize_slice calls) with pinned HOST buffers: H2D + kernels + D2H inside the timing.
            Beside it, measured in the same run on the same box:
              pcie_floor   concurrent raw cudaMemcpyAsync of exactly the step's H2D and D2H bytes from / to pinned
                           memory on every rank (no kernels): what the host link allows; `frac_of_pcie_floor`;
              pageable     the step with the buffers the reference's caller really passes (pageable source, output
                           in a fresh anonymous mmap per step = "cold", cast.rs:158-161; or a reused one = "warm").
  e2e_strong  (N > 1, rank 0, the other ranks idle) ONE process, `ggq_set_shard_devices(N)` + `ggq_slices` over
            ONE copy of the 8 tensors: the library's own by-tensor / by-block-range split (ggq_plan_shards),
            bytes asserted equal to the single-device result.
  roofline  the dominant kernel (the batched dequantize grid: 100 % of the step's launches): algorithmic bytes per
            launch / average launch duration from the CUDA events of the timed region.
  quant_roofline  the K-quant quantize kernels against the FP32-pipe and instruction-issue rooflines that bound them.
  cpu_baseline  the CPU oracle port (oracle/, the reference's algorithm restated in C; the Rust
            reference cannot be built in this image) on the host cores, the whole workload.

`--impl reference` times that CPU port on the same workload (all 8 tensors per step).
Multi-GPU (`torchrun ... bench.py --gpus N`): tensors are independent, each rank processes its own
copy of the workload on its own GPU, no collective on the data path => "scaling": "weak".
"""
import argparse
import json
import mmap
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


STEP_BYTES = sum(algo_bytes(ty, n) for ty, _, n in tensors())
# identical in both arms, so the driver compares like with like
CONFIG = {"workload": WORKLOAD, "bytes_per_step_per_gpu": STEP_BYTES, "tensors_per_step": len(tensors()),
          "l2": "inputs larger than L2: every step touches %.0f MB (packed + f16 of the 8 tensors) against 126 MB of L2, no flush needed" % (STEP_BYTES / 1e6)}


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


def load_profile_json(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
class CpuPort:
    """The oracle port on the host cores over the WHOLE workload: per step, dequantize_slice of all 8 tensors
    (4 types x {4096x14336, 4096x4096}) -> f16, `threads` pthreads over contiguous block ranges."""

    def __init__(self, threads):
        from oracle import oracle as O
        self.O, self.L, self.threads = O, O.lib(), threads
        x = (np.random.default_rng(0).standard_normal(4096 * 14336, dtype=np.float32) * np.float32(0.02)).astype(np.float16)
        self.work, self._maps = [], []
        for ty, sname, n in tensors():
            q = O.quantize(ty, O.F16, x[:n], threads=threads)
            packed = self._huge(q.nbytes, np.uint8)
            packed[:] = q
            self.work.append((ty, n, packed, self._huge(n * 2, np.uint16)))

    def _huge(self, nbytes, dtype):
        """Warm buffers in explicitly huge-page-advised, pre-faulted anonymous memory: whether numpy's own allocations get
        transparent huge pages depends on how fragmented the box is when the process starts, and moved this baseline
        between 39 and 57 GB/s on one box within a minute (profiles/r02_cpu_port_probe.txt).  The CPU gets its best case."""
        m = mmap.mmap(-1, nbytes, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        try:
            m.madvise(mmap.MADV_HUGEPAGE)
        except (AttributeError, OSError):
            pass
        a = np.frombuffer(m, dtype)
        a[:] = 0  # first touch outside every timed region
        self._maps.append(m)
        return a

    def step(self, cold=False):
        for ty, n, packed, out in self.work:
            e, _ = BLOCK[ty]
            if cold:  # cast.rs:158-161: the caller maps a fresh anonymous region for every tensor it produces
                m = mmap.mmap(-1, n * 2, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)  # what memmap2's map_anon maps
                dst = np.frombuffer(m, np.uint16)
            else:
                dst = out
            rc = self.L.ggo_dequantize_slice(ty, self.O.F16, dst.ctypes.data, n, packed.ctypes.data, n // e, self.threads)
            assert rc == 0
            if cold:
                del dst
                m.close()

    def run(self, steps, warmup, cold=False, repeats=1):
        """`warmup` untimed passes, then `steps` timed ones; with repeats > 1 the timed block runs that many times and the
        FASTEST block is reported (the baseline is never penalised for a noisy neighbour on the host)."""
        for _ in range(warmup):
            self.step(cold)
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            for _ in range(steps):
                self.step(cold)
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        return STEP_BYTES * steps / best / 1e9, best / steps * 1e3


def reference_arm(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    port = CpuPort(threads)
    warm = max(args.warmup, 1)
    gbs, ms = port.run(args.steps, warm, repeats=3)
    cold_gbs, cold_ms = port.run(max(1, min(args.steps, 5)), 1, cold=True, repeats=2)
    sample = (f"the whole workload per step: all 8 tensors (4 types x 4096x14336 + 4096x4096, {STEP_BYTES / 1e6:.0f} MB algorithmic), "
              f"{threads} pthreads over contiguous block ranges, buffers in huge-page-advised pre-faulted memory (warm); "
              f"fastest of 3 blocks of {args.steps} steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/f32->f16", "data": "synthetic", "config": CONFIG,
        "note": "CPU port of the reference algorithm (oracle/ggq_oracle.c); the Rust reference cannot be compiled here (no cargo/rustc)",
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "cold": {"value": cold_gbs, "ms_per_step": cold_ms,
                 "what": "output of every tensor in a fresh anonymous mmap (first-touch page faults inside the timing), as the reference's caller allocates it (cast.rs:158-161)"},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import gguf_b200 as g
    from gguf_b200._lib import SliceJob, lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(local_rank) if world > 1 else "single rank"
    cpu_group = None
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its "NCCL version ..." banner there at NCCL_DEBUG=VERSION and
        # WARN (NCCL_DEBUG_FILE does not move it), so those two levels are switched off; INFO / TRACE are left to a
        # user who asked for them.  NCCL is only this bench's barrier and max-reduce of the timing, not the data path.
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")  # host-side waits while rank 0 runs the strong-scaling leg
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
    assert step_bytes == STEP_BYTES

    dev_jobs = (SliceJob * len(work))()
    for i, w in enumerate(work):
        dev_jobs[i] = SliceJob(w["ty"], F16, 0, w["out"].data_ptr(), w["n"], w["packed"].data_ptr(), w["nb"])

    def step():  # ONE call: the eight dequantize_slice jobs share one descriptor-table grid
        rc = L.ggq_slices_device(dev_jobs, len(work), stream)
        assert rc == 0, L.ggq_last_error()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor([float(v) for v in vals], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    for _ in range(args.warmup):
        step()
    # ---- timed region 1: K steps back to back, events only at the ends -> `value` and the roofline ----
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
    (elapsed_ms,) = max_over_ranks([t_start.elapsed_time(t_end)])
    value = step_bytes * args.steps * world / (elapsed_ms * 1e-3) / 1e9
    launches_per_step = launches / args.steps

    # ---- informational: every tensor as its own launch (the per-call device API).  R back-to-back launches of ONE
    # kernel between two CUDA events (an event pair around every single launch adds ~3.7 us of drain + timestamp per
    # launch and under-reports these 8-30 us kernels by 10-30 %), rotating over NSETS buffer sets (> L2). ----
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
        per_kernel.append({"kernel": f"dequant_kernel<{NAMES[w['ty']]},f16>", "shape": w["shape"], "us": ms * 1e3,
                           "GBps": w["bytes"] / (ms * 1e-3) / 1e9, "bytes": w["bytes"], "launches_timed": R})
        del sets
    # ---- informational: the quantize direction of the same four types on the FFN shape (f16 -> packed), same method.
    # BASELINE.json's metric names both directions; `value` and `roofline` stay the dequantize workload's. ----
    quant_per_kernel = []
    quant_roofline = None
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
            quant_per_kernel.append({"kernel": f"quantize<{NAMES[w['ty']]},f16>", "shape": "ffn", "us": ms * 1e3,
                                     "GBps": w["bytes"] / (ms * 1e-3) / 1e9, "bytes": w["bytes"], "launches_timed": rq,
                                     "bound": "fp32 issue (bit-faithful scale search)" if w["ty"] in (g.Q4K, g.Q6K) else "hbm"})
            del outs
        del xs
        # K-quant quantize is bound by the SM's arithmetic, not by HBM.  Two rooflines, both from the committed
        # ncu capture of the shipped kernels (tools/quant_roofline.py writes the JSON): issue = warp instructions /
        # (schedulers x clock); fp32 pipe = the time the FP32 pipe alone needs (a packed FADD2 / FMUL2 / FFMA2 holds it two
        # cycles, so this counts lane operations).  `frac` is against the tighter (larger) of the two.
        qr = load_profile_json("r02_quant_k_roofline.json")
        if qr:
            quant_roofline = {"bound": "fp32 pipe / issue", "source": "profiles/r02_quant_k_roofline.json (smsp__inst_executed and sm__pipe_fma_cycles_active from ncu, shipped kernels)", "kernels": []}
            sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
            for row in quant_per_kernel:
                key = {"quantize<Q4_K,f16>": "Q4K", "quantize<Q6_K,f16>": "Q6K"}.get(row["kernel"])
                if key and key in qr.get("instr_per_elem", {}):
                    ipe = qr["instr_per_elem"][key]   # thread-level instructions per element
                    n_el = ffn[0]["n"]
                    issue_us = ipe * n_el / 32.0 / (sm_count * 4 * qr.get("sm_clock_mhz", 1965.0) * 1e6) * 1e6
                    fp32_us = qr.get("kernels", {}).get(key, {}).get("fp32_pipe_limit_us")
                    limit = max(issue_us, fp32_us or 0.0)
                    quant_roofline["kernels"].append({"kernel": row["kernel"], "instr_per_elem": ipe, "issue_limit_us": issue_us,
                                                      "fp32_pipe_limit_us": fp32_us, "achieved_us": row["us"],
                                                      "frac_issue": issue_us / row["us"],
                                                      "frac_fp32_pipe": (fp32_us / row["us"]) if fp32_us else None,
                                                      "frac": limit / row["us"]})
    clocks = sampler.stop()
    peak, peak_src = measured_peak()
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
    tj = load_profile_json("r02_traffic.json")
    if tj:
        traffic = tj.get("dequant_batch_step_f16")
    step_us = elapsed_ms / args.steps * 1e3
    batched = launches_per_step < 1.5
    roofline = {"bound": "hbm",
                "kernel": "dequant_batch_kernel<f16,16384> (the step's 8 tensors in one grid)" if batched else "step of %.1f launches" % launches_per_step,
                "achieved": step_bytes / (step_us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": step_bytes / (step_us * 1e-6) / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                "frac_of_nominal_8000_GBps": step_bytes / (step_us * 1e-6) / 1e9 / 8000.0, "algorithmic_bytes_per_launch": step_bytes,
                "avg_launch_us": step_us, "launches_per_step": launches_per_step,
                "method": f"{args.steps} back-to-back launches between two CUDA events on the launch stream (the timed region of `value`); "
                          f"each launch touches {step_bytes / 1e6:.0f} MB > 126 MB L2"}

    # ---- e2e legs ----
    e2e = None
    e2e_strong = None
    if not args.no_e2e:
        h2d_bytes = int(sum(w["packed"].numel() for w in work))
        d2h_bytes = int(sum(w["n"] * 2 for w in work))
        e2e_steps = max(1, min(args.steps, args.e2e_steps))

        # (a) same-run raw floor: exactly the step's bytes, pinned, both directions at once, no kernels
        h_in = torch.empty(h2d_bytes, dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty(d2h_bytes, dtype=torch.uint8, pin_memory=True)
        d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
        d_out = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
        s_up, s_down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def raw_step(up=True, down=True):
            if up:
                with torch.cuda.stream(s_up):
                    d_in.copy_(h_in, non_blocking=True)
            if down:
                with torch.cuda.stream(s_down):
                    h_out.copy_(d_out, non_blocking=True)

        def timed(fn, reps):
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return time.perf_counter() - t0
        t_both = min(timed(raw_step, e2e_steps) for _ in range(2))
        t_up = timed(lambda: raw_step(True, False), e2e_steps)
        t_down = timed(lambda: raw_step(False, True), e2e_steps)
        t_both, t_up, t_down = max_over_ranks([t_both, t_up, t_down])
        # The floor is what full duplex allows: the slower direction alone (measured), the other hidden behind it.  The
        # measured concurrent pair is reported beside it: the copy engines do not always overlap two single large
        # copies (one run of this bench measured 44 GB/s for the pair and 71 GB/s for the chunked pipeline on the same box).
        t_floor = max(t_up, t_down)
        floor = {"value": step_bytes * e2e_steps * world / t_floor / 1e9, "unit": UNIT, "ms_per_step": t_floor / e2e_steps * 1e3,
                 "h2d_alone_GBps_per_gpu": h2d_bytes * e2e_steps / t_up / 1e9, "d2h_alone_GBps_per_gpu": d2h_bytes * e2e_steps / t_down / 1e9,
                 "concurrent_pair_measured": step_bytes * e2e_steps * world / t_both / 1e9,
                 "what": "every rank copies exactly h2d_bytes_per_step up and d2h_bytes_per_step down per step (pinned, one cudaMemcpyAsync each, no "
                         "kernels), each direction alone; value = the step's algorithmic bytes / the slower direction's time (full duplex hides the "
                         "other), max over ranks; concurrent_pair_measured = both copies issued together on two streams"}
        del h_in, h_out, d_in, d_out

        # (b) pinned host buffers through ggq_slices
        host = []
        for w in work:
            pin_in, pin_out = g.PinnedBuffer(w["packed"].numel()), g.PinnedBuffer(w["n"] * 2)
            pin_in.array[:] = w["packed"].cpu().numpy()
            host.append((w, pin_in, pin_out))
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
        dt = timed(host_step, e2e_steps)
        dt_per_call = timed(host_step_per_call, e2e_steps)
        # every tensor: the host path and the device path agree byte for byte
        for w, pi, po in host:
            assert np.array_equal(po.view(np.uint16), w["out"].cpu().numpy().view(np.uint16)), "host/device path mismatch"

        # (c) pageable caller memory: what cast.rs hands in (file mmap / previous anonymous mmap) and out (fresh anonymous mmap)
        pg_in = [np.array(pi.array, copy=True) for _, pi, _ in host]
        pg_out = [np.zeros(w["n"], np.uint16) for w, _, _ in host]

        def pageable_step(cold):
            maps = []
            pj = (SliceJob * len(host))()
            for i, (w, _, _) in enumerate(host):
                if cold:
                    m = mmap.mmap(-1, w["n"] * 2, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
                    maps.append(m)
                    dst = np.frombuffer(m, np.uint16)
                else:
                    dst = pg_out[i]
                pj[i] = SliceJob(w["ty"], F16, 0, dst.ctypes.data, w["n"], pg_in[i].ctypes.data, w["nb"])
                del dst
            rc = L.ggq_slices(pj, len(host))
            assert rc == 0, L.ggq_last_error()
            for m in maps:
                m.close()
        pg_steps = max(1, min(e2e_steps, 5))
        dt_warm = timed(lambda: pageable_step(False), pg_steps)
        assert np.array_equal(pg_out[1], host[1][2].view(np.uint16)), "pageable path mismatch"
        dt_cold = timed(lambda: pageable_step(True), pg_steps)
        dt, dt_per_call, dt_warm, dt_cold = max_over_ranks([dt, dt_per_call, dt_warm, dt_cold])
        del pg_in, pg_out
        e2e_value = step_bytes * e2e_steps * world / dt / 1e9
        e2e = {"value": e2e_value, "unit": UNIT, "steps": e2e_steps, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
               "api": "ggq_slices: one synchronous call per step over the 8 tensors (host pointers, pinned)", "ms_per_step": dt / e2e_steps * 1e3,
               "per_call_value": step_bytes * e2e_steps * world / dt_per_call / 1e9,
               "per_call_api": "8 separate ggq_dequantize_slice calls per step (pipeline drains between tensors)",
               "pcie_floor": floor, "frac_of_pcie_floor": e2e_value / floor["value"],
               # where the host side is not full duplex (this pool's 4- and 8-GPU boxes: H2D and D2H share a limit), the pair of
               # raw copies issued together is what a pipeline can reach at best
               "frac_of_concurrent_copy_pair": e2e_value / floor["concurrent_pair_measured"],
               "pageable": {"value": step_bytes * pg_steps * world / dt_warm / 1e9, "cold_value": step_bytes * pg_steps * world / dt_cold / 1e9,
                            "unit": UNIT, "steps": pg_steps, "frac_of_pcie_floor": step_bytes * pg_steps * world / dt_warm / 1e9 / floor["value"],
                            "what": "same call, pageable caller buffers (numpy) bounced through the library's pinned staging; cold = output in a fresh "
                                    "anonymous mmap per tensor per step (cast.rs:158-161: first-touch faults inside the timing)"}}

        # (d) strong scaling through the library's own split: one process, all N GPUs, ONE copy of the workload
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=cpu_group)  # every rank is idle from here until rank 0 is done
            if rank == 0:
                for _, _, po in host:
                    po.array[:] = 0
                nd = L.ggq_set_shard_devices(world)
                try:
                    host_step()
                    t0 = time.perf_counter()
                    for _ in range(e2e_steps):
                        host_step()
                    dts = time.perf_counter() - t0
                finally:
                    L.ggq_set_shard_devices(1)
                for w, pi, po in host:
                    assert np.array_equal(po.view(np.uint16), w["out"].cpu().numpy().view(np.uint16)), "sharded bytes differ from the single-device bytes"
                from gguf_b200._lib import ShardPiece
                npieces = L.ggq_plan_shards(jobs, len(host), world, None, 0)
                pieces = (ShardPiece * npieces)()
                L.ggq_plan_shards(jobs, len(host), world, pieces, npieces)
                load = [0] * world
                for p in pieces:
                    w = host[p.job][0]
                    e, b = BLOCK[w["ty"]]
                    load[p.device] += (p.elem_end - p.elem_begin) // e * b + (p.elem_end - p.elem_begin) * 2
                e2e_strong = {"value": step_bytes * e2e_steps / dts / 1e9, "unit": UNIT, "n_devices": int(nd), "steps": e2e_steps, "ms_per_step": dts / e2e_steps * 1e3,
                              "scaling": "strong", "speedup_vs_this_ranks_weak_share": (step_bytes * e2e_steps / dts / 1e9) / (e2e_value / world),
                              "pieces": int(npieces), "pcie_bytes_per_device": load, "bytes_equal_single_device": True,
                              "api": "ONE process: ggq_set_shard_devices(N) then ggq_slices over one copy of the 8 tensors (pinned host buffers); "
                                     "split = ggq_plan_shards; the other ranks wait on a gloo barrier"}
            dist.barrier(group=cpu_group)
        for _, pi, po in host:
            pi.free(); po.free()

    # ---- informational (rank 0, N=1): the QUANTIZE direction end to end, where the arithmetic, not the host link, is the
    # reference's cost: f16 -> Q4_K / Q8_0 of one 4096x14336 tensor through ggq_quantize_slice (pinned host buffers, H2D +
    # kernel + D2H timed) against the CPU port on a bounded sample with all host threads ----
    e2e_quant = None
    if rank == 0 and world == 1 and not args.no_e2e and not args.no_cpu:
        from oracle import oracle as O
        n = SHAPES["ffn"][0] * SHAPES["ffn"][1]
        threads = os.cpu_count() or 1
        px = g.PinnedBuffer(n * 2)
        px.view(np.uint16)[:] = (np.random.default_rng(5).standard_normal(n, dtype=np.float32) * np.float32(0.02)).astype(np.float16).view(np.uint16)
        e2e_quant = {"unit": UNIT, "tensor": "4096x14336 f16", "cpu_threads": threads, "rows": []}
        for ty, name, sample in ((Q4K, "Q4_K", 1 << 22), (Q8_0, "Q8_0", n)):
            e, b = BLOCK[ty]
            pq = g.PinnedBuffer(n // e * b)
            nbytes = n * 2 + n // e * b
            for _ in range(2):
                assert L.ggq_quantize_slice(ty, F16, pq.ptr, n // e, px.ptr, n) == 0, L.ggq_last_error()
            t0 = time.perf_counter()
            for _ in range(3):
                L.ggq_quantize_slice(ty, F16, pq.ptr, n // e, px.ptr, n)
            gpu_gbs = nbytes * 3 / (time.perf_counter() - t0) / 1e9
            xs = np.array(px.view(np.uint16)[:sample], copy=True)
            O.quantize(ty, O.F16, xs[:1 << 20], threads=threads)
            t0 = time.perf_counter()
            want = O.quantize(ty, O.F16, xs, threads=threads)
            cpu_gbs = (sample * 2 + sample // e * b) / (time.perf_counter() - t0) / 1e9
            assert np.array_equal(pq.array[:want.size], want), "e2e quantize differs from the oracle"
            e2e_quant["rows"].append({"type": name, "gpu_e2e": gpu_gbs, "cpu_port": cpu_gbs, "cpu_sample_elems": sample, "ratio": gpu_gbs / cpu_gbs})
            pq.free()
        px.free()

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on the whole workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        port = CpuPort(threads)
        gbs, _ = port.run(3, 2, repeats=3)
        cold_gbs, _ = port.run(2, 1, cold=True, repeats=2)
        cpu = {"value": gbs, "unit": UNIT, "cores": threads, "kind": "port", "cold_value": cold_gbs,
               "sample": f"the whole workload (8 tensors, {STEP_BYTES / 1e6:.0f} MB algorithmic per pass), fastest of 3 blocks of 3 passes, oracle/ggq_oracle.c with {threads} pthreads, huge-page-advised buffers; "
                         "cold_value = outputs in fresh anonymous mmaps (cast.rs:158-161)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/f32->f16", "data": "synthetic", "config": CONFIG,
            "notes": {"l2": "each step touches %.0f MB (packed + f16 of 8 tensors) > 126 MB L2; no flush needed" % (step_bytes / 1e6),
                      "sharding": "by tensor, one replica of the workload per GPU, no collective", "cpu_binding_rank0": numa},
            "clocks": clocks, "e2e": e2e, "e2e_strong": e2e_strong, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "frac_of_peak_whole_step": value / world / peak, "per_kernel": per_kernel, "quant_per_kernel": quant_per_kernel,
            "quant_roofline": quant_roofline, "e2e_quantize": e2e_quant,
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
        if args.steps > 100:  # default --steps is sized for the 0.13 ms GPU step; the CPU step is ~20 ms x 8 tensors
            args.steps = 20
        reference_arm(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:])
    ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
