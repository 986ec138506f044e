"""Why does the CPU port run at different speeds in the two bench arms?  Fresh process per condition."""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BODY = r'''
import os, sys, time
sys.path.insert(0, %r)
sys.argv = ["bench.py"]
%s
import importlib.util
spec = importlib.util.spec_from_file_location("bench", os.path.join(%r, "bench.py")); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
p = b.CpuPort(%s)
r1 = p.run(10, 3)
r2 = p.run(3, 1, cold=True)
print("%%-44s threads %%2d  warm %%6.1f GB/s  cold %%6.1f GB/s  affinity %%d" %% (%r, p.threads, r1[0], r2[0], len(os.sched_getaffinity(0))), flush=True)
'''
conds = [
    ("plain", "", "os.cpu_count()"),
    ("import torch", "import torch", "os.cpu_count()"),
    ("import torch + cuda init", "import torch; torch.cuda.init(); torch.zeros(1, device='cuda')", "os.cpu_count()"),
    ("import torch + cuda + 1 GB pinned", "import torch; torch.zeros(1, device='cuda'); _pin = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()", "os.cpu_count()"),
    ("plain, 2x threads", "", "2 * os.cpu_count()"),
    ("plain again", "", "os.cpu_count()"),
]
for f in ("/sys/kernel/mm/transparent_hugepage/enabled", "/sys/kernel/mm/transparent_hugepage/defrag"):
    try:
        print(f, open(f).read().strip())
    except OSError as e:
        print(f, e)
print("cpu_count", os.cpu_count(), flush=True)
for name, pre, thr in conds:
    subprocess.run([sys.executable, "-c", BODY % (ROOT, pre, ROOT, thr, name)], check=False)
