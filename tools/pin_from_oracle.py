"""The parity-pinning kit's output as the CPU restatement (oracle/) produces it: the same file rust/examples/dump_golden.rs
writes on the real engine.  usage: python tools/pin_from_oracle.py [pin_input.txt] [out.txt] [key=value ...]
key=value sets a switch of the restatement (window_plus1, wrap_with_slack, slack_to_zero: the [EXT] alternatives E3-E5)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tools import pin_format as P  # noqa: E402


def run(inp, **switches):
    """-> (queries, tresults, states) of the restatement; RANDOMNESS = 0 like the dumper (the reference's draws come from
    an unseeded thread RNG: with a zero weight the random term is +-0 whatever was drawn)"""
    W, H, d = float(inp["W"]), float(inp["H"]), float(inp["d"])
    plus1 = int(switches.get("window_plus1", 0))
    e4, e5 = int(switches.get("wrap_with_slack", 0)), int(switches.get("slack_to_zero", 0))
    fo = O.Field(W, H, d, bool(inp["toroidal"]))
    fo.set_semantics(e4, e5)
    fo.set_objects(inp["birds"])
    fo.lazy_update()
    queries = [fo.neighbors(float(x), float(y), float(r), relax=bool(rel), plus1=bool(plus1))["id"] for x, y, r, rel in inp["probes"]]
    tres = [O.toroidal_transform(float(a), float(dim)) if k == 0 else O.toroidal_distance(float(a), float(b), float(dim))
            for k, a, b, dim in inp["tprobes"]]
    cfg = O.default_cfg(randomness=0.0, window_plus1=plus1, wrap_with_slack=e4, slack_to_zero=e5)
    sim = O.Sim(W, H, d, bool(inp["toroidal"]), cfg=cfg)
    sim.init(inp["birds"])
    states = {}
    for t in range(1, inp["ticks"] + 1):
        sim.step()
        states[t] = sim.agents().copy()
    return queries, tres, states


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 and "=" not in sys.argv[1] else os.path.join(ROOT, "tests", "golden", "pin_input.txt")
    dst = sys.argv[2] if len(sys.argv) > 2 and "=" not in sys.argv[2] else os.path.join(ROOT, "tests", "golden", "oracle_pin.txt")
    sw = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
    q, t, s = run(P.read_input(src), **sw)
    P.write_output(dst, "oracle/flockers_oracle.c " + " ".join(f"{k}={v}" for k, v in sw.items()), q, t, s)
    print(dst)
