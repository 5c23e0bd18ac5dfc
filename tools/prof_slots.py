"""Minimal driver for ncu: the 16 M-bird bench workload, a dozen ticks on one layout.  usage: prof_slots.py <layout 0|1> [workload]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
from examples_b200 import Field2D

lay = int(sys.argv[1])
name = sys.argv[2] if len(sys.argv) > 2 else "weak16m"
gx, gy, n, _ = bench.workload_geometry(name, 1, None)
W, H = bench.field_size(gx), bench.field_size(gy)
loc = bench.make_birds(n, W, H, False)
f = Field2D(W, H, bench.D, True, capacity=int(n * 1.25) + 65536)
f.set_layout(lay)
f.set_objects(np.arange(n, dtype=np.uint32), loc, np.zeros_like(loc))
f.lazy_update()
f.use_graph(False)
f.run(0, 14)
f.synchronize()
f.profile_enable(2)
f.run(14, 10)
f.synchronize()
print("layout", f.layout()[0], "rebin", [round(x, 4) for x in f.profile_rebin()], "k_step", f.profile_step_kernel()[0] / 10)
