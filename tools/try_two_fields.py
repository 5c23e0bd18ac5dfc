"""Premise test: do the memory-bound lazy_update of one field and the FP32-bound Bird::step of ANOTHER, independent
field overlap on one B200 when lazy_update runs on a high-priority stream (FLK_LAZY_PRIO=1, an experiment switch that is not in the tree any more)?  Two fields of 8 M birds
each against one field of 16 M.   python tools/try_two_fields.py"""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np
import bench
from examples_b200 import Field2D
D = bench.D
def field(n, gx, gy, seed):
    W, H = bench.field_size(gx), bench.field_size(gy)
    loc = bench.make_birds(n, W, H, False, seed=seed)
    f = Field2D(W, H, D, True, capacity=n)
    f.set_objects(np.arange(n, dtype=np.uint32), loc, np.zeros_like(loc)); f.lazy_update()
    return f
def drive(fields, ticks, t0):
    for t in range(t0, t0 + ticks):
        for f in fields:
            f.step(t)
        for f in fields:
            f.lazy_update()
    for f in fields:
        f.synchronize()
one = [field(16_000_000, 672, 5376, 1)]
two = [field(8_000_000, 336, 5376, 2), field(8_000_000, 336, 5376, 3)]
for name, fs in (("one field of 16 M", one), ("two fields of 8 M", two)):
    drive(fs, 10, 0)
    t = time.perf_counter(); drive(fs, 100, 10); dt = (time.perf_counter() - t) / 100 * 1e3
    print(f"{name}: {dt:.3f} ms per tick of 16 M birds (FLK_LAZY_PRIO={os.environ.get('FLK_LAZY_PRIO', '0')})")
