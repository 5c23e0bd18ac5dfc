"""Do two independent fields on two streams of ONE GPU overlap (step of one with rebin of the other)?  (experiment)"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from examples_b200 import Field2D
import bench
D = bench.D
def make(gx, gy, n):
    W, H = bench.field_size(gx), bench.field_size(gy)
    loc = bench.make_birds(n, W, H, False)
    f = Field2D(W, H, D, True, capacity=n)
    f.use_graph(False)
    f.set_objects(np.arange(n, dtype=np.uint32), loc, np.zeros_like(loc)); f.lazy_update(); f.run(0, 5); f.synchronize()
    return f
def timeit(fs, ticks=40, interleave=True):
    t0 = time.perf_counter()
    if interleave:
        for t in range(ticks):
            for f in fs: f.run(10 + t, 1)
    else:
        for f in fs: f.run(10, ticks)
    for f in fs: f.synchronize()
    return (time.perf_counter() - t0) / ticks * 1e3
one = make(672, 5376, 16_000_000)
print("one 16M field: %.3f ms/tick" % timeit([one], interleave=False)); one.close()
for k in (2, 4, 8):
    fs = [make(672 // k, 5376, 16_000_000 // k) for _ in range(k)]
    print("%d independent fields of %d birds, interleaved launches: %.3f ms per tick of all" % (k, 16_000_000 // k, timeit(fs)))
    print("%d independent fields, each enqueued whole: %.3f" % (k, timeit(fs, interleave=False)))
    for f in fs: f.close()
