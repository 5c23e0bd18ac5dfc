"""Does running k in-process strips on ONE GPU (different streams) overlap the memory-bound rebin of one strip with
the issue-bound step of another?  (experiment)"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from examples_b200 import Field2D
from examples_b200.distributed import MgpuGroup
import bench
D = bench.D
gx, gy, n, _ = bench.workload_geometry("weak16m", 1, None)
W, H = bench.field_size(gx), bench.field_size(gy)
loc = bench.make_birds(n, W, H, False)
ids = np.arange(n, dtype=np.uint32); ld = np.zeros_like(loc)
def timeit(obj, ticks=40):
    obj.run(0, 10); obj.synchronize()
    t0 = time.perf_counter(); obj.run(10, ticks); obj.synchronize()
    return (time.perf_counter() - t0) / ticks * 1e3
f = Field2D(W, H, D, True, capacity=n); f.set_objects(ids, loc, ld); f.lazy_update()
print("single field: %.3f ms/tick" % timeit(f)); f.close()
for k in (2, 3, 4, 6, 8):
    g = MgpuGroup(W, H, D, capacity_per_rank=int(n / k * 1.15) + 100000, devices=[0] * k)
    g.scatter_objects(ids, loc, ld); g.commit(); g.synchronize()
    ms = timeit(g)
    print("%d strips on one GPU: %.3f ms/tick  (%.2fe10 updates/s)  owned %d" % (k, ms, n / ms / 1e7, sum(g.num_owned())))
    g.close()
