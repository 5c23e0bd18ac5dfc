"""Debug driver for the slotted layout: slots vs sorted runs on the same field, first mismatch located, step-kernel counters."""
import ctypes as C
import sys
import numpy as np
sys.path.insert(0, ".")
from examples_b200 import Field2D
from tests.helpers import D, from_arrays, to_arrays, uniform_birds


def counters(f):
    out = (C.c_uint32 * 16)()
    f._L.flk_debug_counters(f._h, out)
    return list(out)[:9]


# timing: k_step per tick, both layouts
import bench
for name in ("1m", "weak16m"):
    gx, gy, n, _ = bench.workload_geometry(name, 1, None)
    W, H = bench.field_size(gx), bench.field_size(gy)
    loc = bench.make_birds(n, W, H, False)
    ids = np.arange(n, dtype=np.uint32)
    for lay in (2, 1, 0):
        f = Field2D(W, H, D, True, capacity=int(n * 1.25) + 65536)
        f.set_layout(lay)
        f.set_objects(ids, loc, np.zeros_like(loc))
        f.lazy_update()
        f.run(0, 20)
        f.synchronize()
        f.profile_enable(2)
        f.run(20, 30)
        f.synchronize()
        ms, k = f.profile_step_kernel()
        reb = f.profile_rebin()
        tick = f.elapsed_ms() / 30
        f.profile_enable(0)
        f.run(50, 100)
        f.synchronize()
        print(f"{name} layout {f.layout()}: k_step {ms / k:.4f} ms, rebin {[round(x, 4) for x in reb]}, tick {tick:.4f} ms, "
              f"graph tick {f.elapsed_ms() / 100:.4f} ms", flush=True)
        f.close()
