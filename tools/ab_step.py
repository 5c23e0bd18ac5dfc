"""A/B of k_step build variants on the 16 M-bird workload: per variant (a libflk.so built with FLK_NVCC_EXTRA /
FLK_LIB_OUT) prints tick ms, k_step ms, rebin ms and a digest of the final state (variants must agree bit for bit).
    python tools/ab_step.py build_variants/libflk_a.so build_variants/libflk_b.so ...   (no argument: the in-tree lib)"""
import hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("AB_CHILD"):
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    from examples_b200 import Field2D
    ticks = int(os.environ.get("AB_TICKS", "100"))
    gx, gy, n, _ = bench.workload_geometry(os.environ.get("AB_WORKLOAD", "weak16m"), 1, None)
    W, H = bench.field_size(gx), bench.field_size(gy)
    loc = bench.make_birds(n, W, H, os.environ.get("AB_WORKLOAD") == "clustered")
    f = Field2D(W, H, bench.D, True, capacity=n)
    f.set_objects(np.arange(n, dtype=np.uint32), loc, np.zeros_like(loc)); f.lazy_update()
    f.run(0, 10); f.synchronize()
    f.profile_enable(2)
    f.run(10, ticks); f.synchronize()
    ms = f.elapsed_ms(); k_ms, k_n = f.profile_step_kernel(); rb = f.profile_rebin()
    f.profile_enable(False)
    f.run(10 + ticks, ticks); f.synchronize()
    ms_graph = f.elapsed_ms()
    i, l, d = f.snapshot_by_id()
    print(json.dumps({"lib": os.environ.get("FLK_LIB", "in-tree"), "tick_ms_profiled": ms / ticks, "tick_ms_graph": ms_graph / ticks,
                      "k_step_ms": k_ms / k_n, "scan_scatter_rank_ms": rb,
                      "updates_per_s": n * ticks / (ms_graph * 1e-3),
                      "digest": hashlib.sha256(i.tobytes() + l.tobytes() + d.tobytes()).hexdigest()[:16]}), flush=True)
else:
    libs = sys.argv[1:] or [None]
    for lib in libs:
        env = dict(os.environ, AB_CHILD="1")
        if lib:
            env["FLK_LIB"] = os.path.abspath(lib)
        subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False)
