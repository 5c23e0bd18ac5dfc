"""ncu driver: two in-process strips of `n` birds each on ONE GPU (peer copies instead of NCCL), a dozen ticks — the
per-kernel durations of a strip's tick.  usage: prof_strips.py [birds_per_strip] [ticks]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
from examples_b200.distributed import MgpuGroup

per = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 12
gx, gy, n, _ = bench.workload_geometry("weak16m", 2, per)
W, H = bench.field_size(gx), bench.field_size(gy)
loc = bench.make_birds(n, W, H, False)
ids = np.arange(n, dtype=np.uint32)
g = MgpuGroup(W, H, bench.D, capacity_per_rank=int(per * 1.25) + 65536, devices=[0, 0])
g.scatter_objects(ids, loc, np.zeros_like(loc)); g.commit(); g.synchronize()
g.run(0, ticks); g.synchronize()
import time
t0 = time.perf_counter(); g.run(ticks, ticks); g.synchronize()
print("strips", W, H, n, "ms/tick (both strips on one GPU)", (time.perf_counter() - t0) / ticks * 1e3)
