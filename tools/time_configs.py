"""Tick time of the classic form on the BASELINE geometries (graph replays + per-kernel events).  usage: time_configs.py [names...]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import bench
from examples_b200 import Field2D

names = sys.argv[1:] or ["shipped", "1m", "weak16m"]
for name in names:
    if name == "shipped":
        gx = gy = 15; n = 1000
    elif name.startswith("n"):           # n<birds>: a square field of the shipped density
        n = int(name[1:])
        gx = gy = max(int(round((n / 0.0996) ** 0.5 / bench.D)), 6)
    else:
        gx, gy, n, _ = bench.workload_geometry(name, 1, None)
    W, H = bench.field_size(gx), bench.field_size(gy)
    loc = bench.make_birds(n, W, H, name == "clustered")
    f = Field2D(W, H, bench.D, True, capacity=int(n * 1.25) + 1024)
    f.set_objects(np.arange(n, dtype=np.uint32), loc, np.zeros_like(loc))
    f.lazy_update()
    f.run(0, 20)
    f.synchronize()
    f.profile_enable(2)
    k = 30 if n > 100000 else 200
    f.run(20, k)
    f.synchronize()
    ms, kn = f.profile_step_kernel()
    reb = f.profile_rebin()
    f.profile_enable(0)
    reps = 100 if n > 100000 else 2000
    f.run(20 + k, reps)
    f.synchronize()
    t = f.elapsed_ms() / reps
    ids, loc2, ld2 = f.snapshot_by_id()
    import hashlib
    dig = hashlib.sha256(loc2.tobytes() + ld2.tobytes()).hexdigest()[:12]
    print(f"{name}: n={n} k_step {ms / kn * 1e3:.1f} us, rebin {[round(x * 1e3, 1) for x in reb]} us, graph tick {t * 1e3:.1f} us, "
          f"{n / t / 1e6:.1f} M upd/s... {n / (t * 1e-3) / 1e9:.3f}e9 upd/s, digest {dig}", flush=True)
    f.close()
