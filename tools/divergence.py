#!/usr/bin/env python
"""Multi-step trajectory divergence, reported rather than hidden (north star).

  exact mode vs oracle         must be 0 at every tick (bit-identical)            -> checked, reported
  fast mode vs exact mode      per-tick max |dx| and the growth over 200 ticks    -> reported
  arbitrary bag order (CPU)    the reference's own indeterminacy: oracle with canonical bags vs oracle with
                               threaded, unsorted bags                               -> reported
Writes profiles/<tag>_divergence.json.   Run on a GPU box:  python tools/divergence.py r01
"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from tests.helpers import D, to_arrays, from_arrays

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
W = 100.0
n = 1000
birds = O.make_uniform(0xF10C, W, W, n)
ticks = 200
out = {"config": "BASELINE config 1: 1000 birds, 100x100 torus, 200 ticks, Philox draws seed 0xB12D", "field_width": W,
       "tolerance_per_step": 1e-5 * W}

def torus_err(a, b, dim):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    return np.minimum(d, dim - d)

from examples_b200 import Field2D
def gpu(mode):
    f = Field2D(W, W, D, True, capacity=n)
    f.set_math_mode(mode)
    f.set_objects(*to_arrays(birds)); f.lazy_update()
    return f
fe, ff = gpu(0), gpu(1)
sim = O.Sim(W, W); sim.init(birds)
simu = O.Sim(W, W, cfg=O.default_cfg(canonical_bags=0)); simu.init(birds)
rows = []
for t in range(ticks):
    fe.run(t, 1); ff.run(t, 1); sim.step(); simu.step(threads=4)
    ge = from_arrays(*fe.snapshot_by_id()); gf = from_arrays(*ff.snapshot_by_id()); o = sim.agents(); ou = simu.agents()
    exact_bits = bool(np.array_equal(ge.view(np.uint8), o.view(np.uint8)))
    rows.append({"tick": t + 1, "exact_vs_oracle_bit_identical": exact_bits,
                 "fast_vs_exact_max_abs": float(max(torus_err(gf["x"], ge["x"], W).max(), torus_err(gf["y"], ge["y"], W).max())),
                 "bag_order_max_abs": float(max(torus_err(ou["x"], o["x"], W).max(), torus_err(ou["y"], o["y"], W).max()))})
out["ticks"] = [r for r in rows if r["tick"] in (1, 2, 5, 10, 20, 50, 100, 150, 200)]
out["exact_bit_identical_all_ticks"] = all(r["exact_vs_oracle_bit_identical"] for r in rows)
out["fast_first_tick_max_abs"] = rows[0]["fast_vs_exact_max_abs"]
out["bag_order_first_tick_max_abs"] = rows[0]["bag_order_max_abs"]
out["note"] = ("single-step errors are ~1e-6 (a few ulp of the position); the model is chaotic, so any perturbation - fast "
               "division on the GPU or a different bag order inside the reference itself - grows to O(1) within ~100 ticks. "
               "Only the exact mode (default) follows the oracle bit for bit.")
# one-step error of fast mode started from the SAME state each tick (per-step tolerance check)
f1 = gpu(1)
per_step = []
for t in range(50):
    ids, loc, ld = fe.snapshot()
    f1.clear(); f1.set_objects(ids, loc, ld); f1.lazy_update()
    fe.run(ticks + t, 1); f1.run(ticks + t, 1)
    a = from_arrays(*fe.snapshot_by_id()); b = from_arrays(*f1.snapshot_by_id())
    per_step.append(float(max(torus_err(a["x"], b["x"], W).max(), torus_err(a["y"], b["y"], W).max())))
out["fast_one_step_error_from_identical_state"] = {"max_over_50_ticks": max(per_step), "within_tolerance": max(per_step) <= 1e-5 * W}
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "profiles", f"{tag}_divergence.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
