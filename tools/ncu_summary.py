"""Per-launch text summary of an `ncu --set full` capture (reads the .ncu-rep with `ncu -i ... --page raw --csv`,
no GPU needed).  usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_elems] [--traffic-json out.json]

The metrics are the ones DESIGN.md argues from: DRAM bytes, issue-active, pipe utilisation, occupancy limits,
shared-memory bank conflicts and the top stall reasons (per issue-active cycle)."""
import csv
import io
import json
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]
STALL = re.compile(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio$")
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep = sys.argv[1]
    n_elems = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
    tj = sys.argv[sys.argv.index("--traffic-json") + 1] if "--traffic-json" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {}
    for i, h in enumerate(head):
        for m in METRICS:
            if h == m or h.endswith("." + m):
                col.setdefault(m, i)
    stalls = [(i, STALL.search(h).group(1)) for i, h in enumerate(head) if STALL.search(h)]
    kname = head.index("Kernel Name")
    traffic = {}
    for r in data:
        print(f"## launch {r[0]}: {r[kname]}")
        vals = {}
        for m in METRICS:
            if m in col and r[col[m]] != "":
                vals[m] = (r[col[m]], units[col[m]])
                print(f"{m:<82s} {r[col[m]]:>18s} {units[col[m]]}")
        if n_elems and "smsp__inst_executed.sum" in vals:
            print(f"{'thread-instructions per element':<82s} {32 * float(vals['smsp__inst_executed.sum'][0]) / n_elems:>18.2f}")
        top = sorted(((float(r[i]), n) for i, n in stalls if r[i] not in ("", "n/a")), reverse=True)[:8]
        for v, n in top:
            print(f"{'stall ' + n + '_per_issue_active.ratio':<82s} {v:>18.2f}")
        if "dram__bytes_read.sum" in vals and "dram__bytes_write.sum" in vals:
            b = sum(float(vals[k][0]) * TO_BYTES.get(vals[k][1], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            traffic.setdefault(r[kname], []).append(b)
        print()
    if tj:
        json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(tj, "w"), indent=1)


if __name__ == "__main__":
    main()
