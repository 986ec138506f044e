"""Writes profiles/r02_quant_k_roofline.json from ncu captures of the shipped K-quant quantize kernels
(gpurun_out/r02_quant_k_<TYPE>.ncu-rep, made by `tools/gpu_round.sh ncu_kq`): warp instructions executed per launch ->
thread-instruction equivalents per element, the issue-rate roofline bench.py reports as `quant_roofline`, and the pipe
utilisation the DESIGN argues from.  usage: python tools/quant_roofline.py [n_elems]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096 * 14336
WANT = {"smsp__inst_executed.sum": "warp_instructions", "gpu__time_duration.sum": "ncu_us",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fp32_pipe_cycles_active_pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
        "launch__registers_per_thread": "registers", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct"}
out = {"n_elems": n, "sm_clock_mhz": 1965.0, "instr_per_elem": {}, "kernels": {},
       "note_fp32": "fp32_pipe_limit_us = ncu_us x sm__pipe_fma_cycles_active: what the launch would take with the FP32 pipe busy every cycle",
       "note": "instr_per_elem = smsp__inst_executed.sum * 32 / n_elems (thread-instruction equivalents); issue roofline = warp "
               "instructions / (148 SMs x 4 schedulers x 1.965 GHz)"}
try:
    PREV = json.load(open(os.path.join(ROOT, "profiles", "r02_quant_k_roofline.json")))
except (OSError, ValueError):
    PREV = {}
for ty in ("Q4K", "Q6K", "Q5K", "Q2K", "Q3K"):
    # r02b_*: captures made after the VIMNMX.RELU clamp (Q4K / Q5K / Q2K changed; tools/gpu_round.sh ncu_kq2)
    reps = [os.path.join(ROOT, "gpurun_out", f"{pre}_quant_k_{ty}.ncu-rep") for pre in ("r02b", "r02")]
    rep = next((r for r in reps if os.path.exists(r)), None)
    if rep is None:
        prev = PREV.get("kernels", {}).get(ty)   # keep the committed record of a kernel that was not re-captured
        if prev:
            out["instr_per_elem"][ty] = prev["instr_per_elem"]
            out["kernels"][ty] = prev
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, data = rows[0], rows[2]
    rec = {"kernel": data[head.index("Kernel Name")] if "Kernel Name" in head else ty}
    for i, h in enumerate(head):
        for m, name in WANT.items():
            if h == m or h.endswith("." + m):
                try:
                    rec[name] = float(data[i].replace(",", ""))
                except ValueError:
                    pass
    ipe = rec["warp_instructions"] * 32 / n
    rec["instr_per_elem"] = ipe
    rec["issue_limit_us"] = rec["warp_instructions"] / (148 * 4 * 1965e6) * 1e6
    # the time the FP32 pipe alone would need: ncu's busy fraction of it x the launch under ncu (a packed FADD2 / FMUL2 /
    # FFMA2 holds the pipe two cycles, so this counts lane operations, not instructions)
    if "fp32_pipe_cycles_active_pct" in rec:
        rec["fp32_pipe_limit_us"] = rec["ncu_us"] * rec["fp32_pipe_cycles_active_pct"] / 100.0
    out["instr_per_elem"][ty] = ipe
    out["kernels"][ty] = rec
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_quant_k_roofline.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
