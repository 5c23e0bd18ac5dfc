"""Hypothesis-driven cross-check (SURVEY.md §8c item ix): on small random populations that deliberately include the
awkward coordinates (0, the seam, x == W, cell boundaries, coincident birds) the C oracle and the independent numpy
brute force agree bit for bit on neighbour sets and on two ticks of Bird::step."""
import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import brute as B
from oracle import oracle as O
from tests.helpers import D, bits_equal

W = 40.0   # 6 x 6 cells: every window touches the seam somewhere
f32 = np.float32
d32 = float(f32(D))
special = [0.0, W, float(np.nextafter(f32(W), f32(0))), d32, float(np.nextafter(f32(d32), f32(0))), 2 * d32, 1e-30,
           W / 2, 39.999, 1e-6, 6.6666665 * 5]
coord = st.one_of(st.sampled_from(special), st.floats(min_value=0.0, max_value=W, width=32, allow_nan=False))
heading = st.floats(min_value=-0.75, max_value=0.75, width=32, allow_nan=False)
bird = st.tuples(coord, coord, heading, heading)


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(st.lists(bird, min_size=1, max_size=24), st.booleans(), st.integers(min_value=0, max_value=2 ** 31))
def test_oracle_equals_brute_force_on_awkward_inputs(rows, relax, seed):
    n = len(rows)
    b = np.zeros(n, dtype=O.BIRD)
    b["id"] = np.arange(n)
    for i, (x, y, lx, ly) in enumerate(rows):
        b["x"][i], b["y"][i], b["ldx"][i], b["ldy"][i] = x, y, lx, ly
    sim = O.Sim(W, W, cfg=O.default_cfg(relax=int(relax)))
    sim.init(b)
    cur = b.copy()
    rng = np.random.default_rng(seed)
    for t in range(2):
        if t == 1:   # READ buffer now holds `cur`: compare the query itself
            for i in range(min(n, 6)):
                got = sorted(sim.field.neighbors(cur["x"][i], cur["y"][i], relax=relax)["id"].tolist())
                assert got == B.neighbor_ids(cur, cur["x"][i], cur["y"][i], W, W, D, relax=relax)
        r = rng.random((n, 2), dtype=np.float32)
        sim.step(injected=r)
        cur = B.tick(cur, W, W, D, r, relax=relax)
        a = sim.agents()
        nan = np.isnan(a["x"]) | np.isnan(cur["x"])
        assert np.array_equal(np.isnan(a["x"]), np.isnan(cur["x"]))
        assert bits_equal(a[~nan], cur[~nan])
