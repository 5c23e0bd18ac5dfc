"""Minimal driver for ncu: one bench workload on the classic form, a dozen ticks.  usage: prof_workload.py <workload> [ticks]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
from examples_b200 import Field2D

name = sys.argv[1]
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 14
if name == "shipped":
    gx, gy, n = 15, 15, 1000
elif name.startswith("n"):           # n<birds>: a square field of the shipped density
    n = int(name[1:])
    gx = gy = max(int(round((n / 0.0996) ** 0.5 / bench.D)), 6)
else:
    gx, gy, n, _ = bench.workload_geometry(name, 1, None)
W, H = bench.field_size(gx), bench.field_size(gy)
loc = bench.make_birds(n, W, H, name == "clustered")
f = Field2D(W, H, bench.D, True, capacity=int(n * 1.25) + 65536)
f.set_objects(np.arange(n, dtype=np.uint32), loc, np.zeros_like(loc))
f.lazy_update()
f.use_graph(False)
f.run(0, ticks)
f.synchronize()
print(name, "ms/tick", f.elapsed_ms() / ticks)
