#!/usr/bin/env python
"""BASELINE.json configs on ONE B200 + the CPU baselines SURVEY.md §8d asks for -> profiles/<tag>_configs.json.
   python tools/report_configs.py r01            (run under gpurun; multi-GPU configs 3/4 come from bench.py under torchrun)
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from examples_b200 import Field2D
from oracle import oracle as O

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
D = bench.D
out = {}


def gpu_run(n, W, H, loc, ld, ticks, warm):
    f = Field2D(W, H, D, True, capacity=n)
    f.set_objects(np.arange(n, dtype=np.uint32), loc, ld)
    f.lazy_update()
    res = {}
    if warm:
        f.run(0, warm); f.synchronize()
    f.run(warm, ticks); ms = f.elapsed_ms()
    res["ticks"] = ticks; res["warmup"] = warm; res["ms_per_tick"] = ms / ticks
    res["bird_updates_per_s"] = n * ticks / (ms * 1e-3)
    return f, res


def cpu_run(birds, W, H, ticks, threads):
    sim = O.Sim(W, H); sim.init(birds)
    sim.step(threads=threads)
    t0 = time.perf_counter()
    for _ in range(ticks):
        sim.step(threads=threads)
    dt = time.perf_counter() - t0
    return len(birds) * ticks / dt


cores = os.cpu_count()
# ---- config 1: shipped headless run (flockers/src/main.rs:40-47): 1000 birds, 100x100, 200 steps
birds = O.make_uniform(0xF10C, 100.0, 100.0, 1000)
loc = np.stack([birds["x"], birds["y"]], 1); ld = np.zeros_like(loc)
f, r = gpu_run(1000, 100.0, 100.0, loc, ld, 200, 0)
r["cpu_oracle_1_thread"] = cpu_run(birds, 100.0, 100.0, 200, 1)
r["cpu_oracle_all_cores"] = cpu_run(birds, 100.0, 100.0, 200, cores)
r["note"] = "200 ticks from the initial placement, no warm-up, CUDA graph replays; launch-latency bound at this size"
out["config1_shipped_1000_birds"] = r; f.close()

# ---- config 2: 1 M birds, 480x480 cells
W = bench.field_size(480)
birds = O.make_uniform(0xF10C, W, W, 1_000_000)
loc = np.stack([birds["x"], birds["y"]], 1); ld = np.zeros_like(loc)
f, r0 = gpu_run(1_000_000, W, W, loc, ld, 100, 0); f.close()
f, r = gpu_run(1_000_000, W, W, loc, ld, 100, 20); f.close()
r["unwarmed_ticks_0_100_updates_per_s"] = r0["bird_updates_per_s"]
r["cpu_oracle_1_thread"] = cpu_run(birds, W, W, 3, 1)
r["cpu_oracle_all_cores"] = cpu_run(birds, W, W, 20, cores)
out["config2_1M_birds"] = r

# ---- config 5: clustered flocks, 16 M birds (672 x 5376 cells)
gx, gy, n, _ = bench.workload_geometry("clustered", 1, None)
W, H = bench.field_size(gx), bench.field_size(gy)
loc = bench.make_birds(n, W, H, True); ld = np.zeros_like(loc)
f, r = gpu_run(n, W, H, loc, ld, 10, 5)
cs, ids = f.binning()
per = np.diff(cs.astype(np.int64))
occ = per[per > 0]
r["birds_per_cell_max"] = int(per.max()); r["birds_per_nonempty_cell_mean"] = float(occ.mean())
r["candidates_per_bird_mean"] = None
# mean 3x3-window population seen by a bird = sum over cells n_c * window(c) / N, from the cell grid
grid = per[: f.dw * f.dh].reshape(f.dw, f.dh)[: f.mx, : f.my].astype(np.float64)
win = sum(np.roll(np.roll(grid, dx, 0), dy, 1) for dx in (-1, 0, 1) for dy in (-1, 0, 1))
r["candidates_per_bird_mean"] = float((grid * win).sum() / grid.sum())
out["config5_clustered_16M"] = r; f.close()

out["host_cores"] = cores
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "profiles", f"{tag}_configs.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
