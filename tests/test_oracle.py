"""CPU tests of the oracle: known-answer cases, RNG KATs, and agreement with the independent brute force.

The reference ships no tests for this path — these are the build's own pins:
hand-computed cases (i)-(vii) of SURVEY.md §8c plus the Random123 Philox4x32-10 known-answer vectors.
"""
import math

import numpy as np
import pytest

from oracle import brute as B
from oracle import oracle as O
from tests.helpers import D, bits_equal, clustered_birds, uniform_birds

f32 = np.float32


def bird(i, x, y, ldx=0.0, ldy=0.0):
    b = np.zeros((), dtype=O.BIRD)
    b["id"], b["x"], b["y"], b["ldx"], b["ldy"] = i, x, y, ldx, ldy
    return b


def field_with(birds, W=100.0, H=100.0):
    f = O.Field(W, H)
    for b in birds:
        f.set_object_location(b)
    f.lazy_update()
    return f


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert [hex(v) for v in O.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(v) for v in O.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)] == \
        ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(v) for v in O.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_draw_is_24_bit_unit_float():
    for i in range(50):
        r1, r2 = O.draw(0xB12D, i, 3)
        for r in (r1, r2):
            assert 0.0 <= r < 1.0
            assert (r * 2 ** 24) == int(r * 2 ** 24)


def test_toroidal_transform_cases():
    assert O.toroidal_transform(5.0, 100.0) == 5.0
    assert O.toroidal_transform(100.0, 100.0) == 0.0
    assert O.toroidal_transform(101.5, 100.0) == 1.5
    assert O.toroidal_transform(-1.5, 100.0) == 98.5
    # E5: tiny negative input lands exactly on dim
    assert O.toroidal_transform(-1e-6, 100.0) == 100.0
    assert math.isnan(O.toroidal_transform(float("nan"), 100.0))


def test_toroidal_distance_cases():
    assert O.toroidal_distance(1.0, 3.0, 100.0) == -2.0
    assert O.toroidal_distance(0.5, 99.5, 100.0) == 1.0     # (iii) across the seam
    assert O.toroidal_distance(99.5, 0.5, 100.0) == -1.0
    assert O.toroidal_distance(0.0, 50.0, 100.0) == -50.0   # |d| == dim/2 stays on the fast path
    for a, b in np.random.default_rng(0).random((200, 2)) * 100:
        assert O.toroidal_distance(f32(a), f32(b), 100.0) == B.tdist(a, b, 100.0)


def test_field_geometry_shipped_config():
    f = O.Field(100.0, 100.0)
    assert (f.mx, f.my, f.dw, f.dh) == (15, 15, 16, 16)           # 15x15 (+1 slack)
    assert f.discretize(99.9, 0.0) == (14, 0)
    assert f.discretize(100.0, 100.0) == (15, 15)                 # slack cell (E5)
    for W, G in ((3200.0, 480), (12800.0, 1920), (35840.0, 5376)):
        assert O.Field(W, 100.0).mx == G


def test_two_birds_hand_computed():
    """(i): one step of bird 0 next to bird 1, checked against a float64 hand evaluation of bird.rs:52-140."""
    f = field_with([bird(0, 50, 50), bird(1, 53, 54, 0.3, -0.4)])
    out, drew = f.bird_step(bird(0, 50, 50), O.default_cfg(), 0.75, 0.75)
    assert drew
    dx, dy = -3.0, -4.0
    den = (dx * dx + dy * dy) ** 2 + 1.0
    avo = (400 * dx / den, 400 * dy / den)
    coh = (3.0 / 10, 4.0 / 10)
    con = (0.3, -0.4)
    rnd = (0.05 * 0.5 / math.sqrt(0.5), 0.05 * 0.5 / math.sqrt(0.5))
    vx = 0.8 * coh[0] + avo[0] + 0.7 * con[0] + 1.1 * rnd[0]
    vy = 0.8 * coh[1] + avo[1] + 0.7 * con[1] + 1.1 * rnd[1]
    n = math.hypot(vx, vy)
    assert out["ldx"] == pytest.approx(vx / n * 0.7, abs=1e-6)
    assert out["ldy"] == pytest.approx(vy / n * 0.7, abs=1e-6)
    assert out["x"] == pytest.approx(50 + vx / n * 0.7, abs=1e-5)
    assert out["y"] == pytest.approx(50 + vy / n * 0.7, abs=1e-5)


def test_lone_bird_only_randomness_and_momentum():
    """(ii): query returns only the caller -> count==0, draws still consumed (bird.rs:42,73-90,103)."""
    f = field_with([bird(7, 20, 20, 0.1, 0.2)])
    out, drew = f.bird_step(bird(7, 20, 20, 0.1, 0.2), O.default_cfg(), 0.25, 0.5)
    assert drew
    xr, yr = -0.5, 0.0
    vx = 1.1 * 0.05 * xr / 0.5 + 0.1
    vy = 0.2
    n = math.hypot(vx, vy)
    assert out["ldx"] == pytest.approx(vx / n * 0.7, abs=1e-6)
    assert out["ldy"] == pytest.approx(vy / n * 0.7, abs=1e-6)


def test_seam_and_corner_pair():
    """(iii): a pair straddling the corner sees each other through the toroidal window and distance."""
    a, b = bird(0, 0.5, 0.5), bird(1, 99.5, 99.5)
    f = field_with([a, b])
    assert sorted(f.neighbors(0.5, 0.5)["id"]) == [0, 1]
    assert sorted(f.neighbors(99.5, 99.5)["id"]) == [0, 1]
    cfg = O.default_cfg(randomness=0.0)
    out, _ = f.bird_step(a, cfg, 0.9, 0.9)
    # bird 1 is at toroidal offset (+1,+1): avoidance (400*1/5) dominates cohesion and pushes 0 away (+x,+y)
    assert out["ldx"] > 0 and out["ldy"] > 0
    assert out["ldx"] == pytest.approx(out["ldy"], abs=1e-7)


def test_consistency_divided_twice():
    """(iv): bird.rs:78-84 divides the consistency sum by count twice."""
    me = bird(0, 50, 50)
    f = field_with([me, bird(1, 50, 50, 0.4, 0.0), bird(2, 50, 50, 0.4, 0.0)])
    cfg = O.default_cfg(cohesion=0.0, avoidance=0.0, randomness=0.0, momentum=0.0, consistency=1.0, jump=1.0)
    out, _ = f.bird_step(me, cfg, 0.9, 0.9)
    # direction only (normalised), so check through a second bird with an orthogonal heading
    assert out["ldx"] == pytest.approx(1.0) and out["ldy"] == 0.0
    f2 = field_with([me, bird(1, 50, 50, 0.4, 0.0), bird(2, 50, 50, 0.0, 0.0)])
    cfg2 = O.default_cfg(cohesion=0.0, avoidance=0.0, randomness=0.0, consistency=1.0, momentum=1.0, jump=1.0)
    me2 = bird(0, 50, 50, 0.0, 0.1)
    out2, _ = f2.bird_step(me2, cfg2, 0.9, 0.9)
    # cons.x = (0.4/2)/2 = 0.1 ; momentum.y = 0.1  -> 45 degrees
    assert out2["ldx"] == pytest.approx(out2["ldy"], rel=1e-6)


def test_zero_displacement_path():
    """(v): dis == 0 keeps dx,dy at 0 (bird.rs:130)."""
    me = bird(3, 10, 10)
    f = field_with([me])
    out, _ = f.bird_step(me, O.default_cfg(randomness=0.0), 0.3, 0.3)
    assert (out["x"], out["y"], out["ldx"], out["ldy"]) == (10.0, 10.0, 0.0, 0.0)


def test_bird_landing_on_width_goes_to_slack_cell():
    """(vi)/E5: x == W puts the bird in the slack column, invisible to every query for one tick."""
    me = bird(0, 0.0, 50.0, -1.0, 0.0)
    other = bird(1, 1.0, 50.0)
    sim = O.Sim(100.0, 100.0, cfg=O.default_cfg(cohesion=0.0, avoidance=0.0, randomness=0.0, consistency=0.0,
                                                momentum=1.0, jump=1e-6))
    sim.init(np.array([me, other]))
    sim.step()
    a = sim.agents()
    assert a["x"][0] == 100.0
    assert sim.field.discretize(100.0, 50.0)[0] == sim.field.mx
    assert 0 not in sim.field.neighbors(1.0, 50.0)["id"]          # invisible to its neighbour
    assert 0 not in sim.field.neighbors(100.0, 50.0)["id"]        # and absent from its own query...
    assert 1 in sim.field.neighbors(100.0, 50.0)["id"]            # ...whose window is columns {mx-1, 0, 1}


def test_two_cells_away_is_absent_from_both_queries():
    """(vii)/E3: window half-width floor(10/6.67)=1, so a bird 8.0 away but 2 cells over is missed."""
    a, b = bird(0, 6.0, 50.0), bird(1, 14.0, 50.0)
    f = field_with([a, b])
    assert f.discretize(6.0, 50.0)[0] == 0 and f.discretize(14.0, 50.0)[0] == 2
    assert list(f.neighbors(6.0, 50.0, relax=True)["id"]) == [0]
    assert list(f.neighbors(6.0, 50.0, relax=False)["id"]) == [0]
    assert sorted(f.neighbors(6.0, 50.0, relax=True, plus1=True)["id"]) == [0, 1]   # MASON-style window (E3 alt)


def test_exact_query_filters_by_distance():
    f = field_with([bird(0, 50, 50), bird(1, 55, 50), bird(2, 59.5, 59.5)])
    assert sorted(f.neighbors(50, 50, relax=True)["id"]) == [0, 1, 2]
    assert sorted(f.neighbors(50, 50, relax=False)["id"]) == [0, 1]       # bird 2 is 13.4 away


@pytest.mark.parametrize("maker,n", [(uniform_birds, 250), (clustered_birds, 200)])
@pytest.mark.parametrize("relax", [True, False])
def test_oracle_matches_brute_force(maker, n, relax):
    """(ix): the C oracle and the numpy brute force agree bit for bit over several ticks."""
    W = H = 100.0
    birds = maker(n, W, H)
    sim = O.Sim(W, H, cfg=O.default_cfg(relax=int(relax)))
    sim.init(birds)
    cur = birds.copy()
    rng = np.random.default_rng(3)
    for _ in range(3):
        r = rng.random((n, 2), dtype=np.float32)
        # neighbour sets of the current READ state
        if sim.steps_done > 0:
            for i in (0, n // 2, n - 1):
                got = sorted(sim.field.neighbors(cur["x"][i], cur["y"][i], relax=relax)["id"])
                assert got == B.neighbor_ids(cur, cur["x"][i], cur["y"][i], W, H, D, relax=relax)
        sim.step(injected=r)
        cur = B.tick(cur, W, H, D, r, relax=relax)
        assert bits_equal(sim.agents(), cur)


def test_threads_and_bag_order_do_not_change_results():
    W = H = 200.0
    birds = uniform_birds(3000, W, H)
    s1, s2 = O.Sim(W, H), O.Sim(W, H)
    s1.init(birds)
    s2.init(birds)
    for _ in range(10):
        s1.step()
        s2.step(threads=4)
    assert bits_equal(s1.agents(), s2.agents())


def test_bag_order_sensitivity_is_within_tolerance():
    """The reference's bag order is unspecified; an arbitrary order (threaded pushes, no canonical re-sort)
    stays within 1e-5*W per tick of the canonical one."""
    W = H = 100.0
    birds = uniform_birds(1000, W, H)
    s1 = O.Sim(W, H)
    s1.init(birds)
    s2 = O.Sim(W, H, cfg=O.default_cfg(canonical_bags=0))
    s2.init(birds)
    for _ in range(2):
        s1.step()
        s2.step(threads=4)
    a, b = s1.agents(), s2.agents()
    assert np.abs(a["x"] - b["x"]).max() <= 2 * 1e-5 * W
    assert np.abs(a["y"] - b["y"]).max() <= 2 * 1e-5 * H


def test_golden_fixture_matches():
    """Regression pin of the oracle itself against tests/golden (see tests/golden/make_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "flockers_1000.npz"))
    birds = g["init"].view(O.BIRD).reshape(-1)
    assert bits_equal(birds, uniform_birds(1000, 100.0, 100.0))
    sim = O.Sim(100.0, 100.0)
    sim.init(birds)
    ticks = list(g["ticks"])
    for t in range(1, max(ticks) + 1):
        sim.step(seed=int(g["seed"]))
        if t in ticks:
            assert bits_equal(sim.agents(), g[f"state_{t}"].view(O.BIRD).reshape(-1)), t
