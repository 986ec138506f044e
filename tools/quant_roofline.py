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
       "note": "instr_per_elem = smsp__inst_executed.sum * 32 / n_elems (thread-instruction equivalents); issue roofline = warp "
               "instructions / (148 SMs x 4 schedulers x 1.965 GHz)"}
for ty in ("Q4K", "Q6K", "Q5K", "Q2K", "Q3K"):
    rep = os.path.join(ROOT, "gpurun_out", f"r02_quant_k_{ty}.ncu-rep")
    if not os.path.exists(rep):
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
    out["instr_per_elem"][ty] = ipe
    out["kernels"][ty] = rec
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_quant_k_roofline.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
