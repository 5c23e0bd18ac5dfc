"""Generates tests/golden/*.npz with the CPU oracle.

The reference cannot run here (no Rust toolchain, engine crate absent — SURVEY.md §8c), so these vectors
are outputs of oracle/flockers_oracle.c, itself cross-checked against oracle/brute.py.  They pin (a) the
oracle against regressions and (b) the CUDA path on the GPU box, where only these files travel.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.helpers import clustered_birds  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def run(birds, W, H, seed, ticks, cfg=None):
    sim = O.Sim(W, H, cfg=cfg)
    sim.init(birds)
    out = {}
    for t in range(1, max(ticks) + 1):
        sim.step(seed=seed)
        if t in ticks:
            out[f"state_{t}"] = sim.agents().view(np.uint8).copy()
    return out


def main():
    # BASELINE config 1: the shipped headless run (flockers/src/main.rs:40-47): 1000 birds, 100x100, 200 steps
    birds = O.make_uniform(0xF10C, 100.0, 100.0, 1000)
    ticks = [1, 2, 10, 50, 200]
    out = run(birds, 100.0, 100.0, 0xB12D, ticks)
    np.savez_compressed(os.path.join(HERE, "flockers_1000.npz"), init=birds.view(np.uint8), ticks=np.array(ticks),
                        seed=np.uint64(0xB12D), W=np.float32(100.0), H=np.float32(100.0), **out)
    # clustered (config 5 shape, small): dense cells, exact-distance query variant too
    cb = clustered_birds(4000, 200.0, 200.0, per_cluster=400, sigma=6.0)
    ticks = [1, 5, 20]
    out = run(cb, 200.0, 200.0, 0xB12D, ticks)
    out2 = run(cb, 200.0, 200.0, 0xB12D, ticks, cfg=O.default_cfg(relax=0))
    np.savez_compressed(os.path.join(HERE, "flockers_clustered_4000.npz"), init=cb.view(np.uint8), ticks=np.array(ticks),
                        seed=np.uint64(0xB12D), W=np.float32(200.0), H=np.float32(200.0), **out,
                        **{k.replace("state_", "exact_"): v for k, v in out2.items()})
    print("golden written")


if __name__ == "__main__":
    main()
