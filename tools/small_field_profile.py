"""Where a tick of a SMALL field goes (BASELINE configs 1 and 2): graph-replay tick, plain-launch tick, k_step and the three
lazy_update phases bracketed by events.   python tools/small_field_profile.py   (on a GPU box)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from examples_b200 import Field2D

out = {}
for name, n, cells in (("config1_1000", 1000, 15), ("8000", 8000, 43), ("config2_1M", 1_000_000, 480)):
    W = bench.field_size(cells)
    loc = bench.make_birds(n, W, W, False)
    f = Field2D(W, W, bench.D, True, capacity=n)
    f.set_objects(np.arange(n, dtype=np.uint32), loc, np.zeros_like(loc)); f.lazy_update()
    f.run(0, 50); f.synchronize()                      # graph captured and warm
    f.run(50, 200); f.synchronize(); graph_ms = f.elapsed_ms() / 200
    f.use_graph(False)
    f.run(250, 200); f.synchronize(); plain_ms = f.elapsed_ms() / 200
    f.profile_enable(2)
    f.run(450, 200); f.synchronize()
    k_ms, k_n = f.profile_step_kernel(); rb = f.profile_rebin()
    f.profile_enable(False)
    out[name] = {"birds": n, "tick_us_graph": graph_ms * 1e3, "tick_us_plain_launches": plain_ms * 1e3,
                 "k_step_us": k_ms / k_n * 1e3, "scan_scatter_rank_us": [x * 1e3 for x in rb]}
    f.close()
print(json.dumps(out, indent=1))
