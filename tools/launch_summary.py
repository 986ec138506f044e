"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel.
usage: python tools/launch_summary.py gpurun_out/r02_bench_launches.csv ["header comment"]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("ggq::", "")
    name = re.sub(r"\(unsigned int\)|\(int\)", "", name)
    a = agg.setdefault(name[:110], [0, 0.0])
    a[0] += 1
    a[1] += float(r[-1]) / 1e3
tot = sum(a[1] for a in agg.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("# kernel, launches, total us, share of all captured launches")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:<112s} {n:5d} {us:12.1f} {100 * us / tot:6.1f}%")
ours = {k: v for k, v in agg.items() if k.startswith(("dequant", "quant", "cast", "rearrange"))}
print(f"# {sum(v[0] for v in ours.values())} launches of this library's kernels, {sum(v[0] for v in agg.values()) - sum(v[0] for v in ours.values())} torch set-up kernels (randn / scale / cast of the synthetic inputs, clones of the rotating buffer sets)")
