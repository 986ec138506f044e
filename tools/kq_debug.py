"""GPU-side diagnostic: per input category, which K-quant blocks differ from the oracle and where."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import gguf_b200 as g
from oracle import oracle as O
from data import edge_blocks, gaussian, to_fdt

rng = np.random.default_rng(3)
for ty in (10, 11, 12, 13, 14):
    n, b = O.block_info(ty)
    cats = {"gauss": gaussian(n * 300, ty), "student": (rng.standard_t(3, n * 100) * 0.02).astype(np.float32),
            "uniform": rng.random(n * 50, dtype=np.float32)}
    eb = edge_blocks(n).reshape(-1, n)
    for i, row in enumerate(eb):
        cats[f"edge{i}"] = row
    for name, x in cats.items():
        for fdt in (0, 1):
            src = to_fdt(x, fdt)
            got = g.quantize(ty, src, fdt).reshape(-1, b)
            want = O.quantize(ty, fdt, src).reshape(-1, b)
            bad = np.where((got != want).any(axis=1))[0]
            if len(bad):
                k = bad[0]
                offs = np.where(got[k] != want[k])[0]
                print(f"{g.TYPE_NAMES[ty]} fdt={fdt} {name}: {len(bad)}/{len(got)} blocks differ; block {k} offsets {offs[:12].tolist()} got {got[k][offs[:6]].tolist()} want {want[k][offs[:6]].tolist()}",
                      "finite" if np.isfinite(x.reshape(-1, n)[k]).all() else "NONFINITE", "absmax %.3g" % np.nanmax(np.abs(x.reshape(-1, n)[k])))
print("done")
