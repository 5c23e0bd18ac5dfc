"""GPU parity of the slot forms of the double buffer (include/flk.h: flk_set_layout): form 2, the slot-staged WRITE buffer
of the sorted-runs layout (default where it fits: lazy_update = scan + k_compact), and form 1, fixed-capacity cells as the
READ buffer too (examples_b200/csrc/flk_slots.cu, kept as a measured alternative).  The same results as the classic form
and as the oracle, bit for bit — including cells that overflow their slots (spill list), the switches between the forms,
CUDA-graph replays and uploads into a slotted WRITE buffer."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import D, bits_equal, clustered_birds, from_arrays, state_digest, to_arrays, uniform_birds

pytestmark = pytest.mark.gpu


def field(birds, W, H, layout=-1, cap=None, slots=None):
    from examples_b200 import Field2D
    env = {}
    if slots is not None:      # fewer slots per cell than the mean occupancy: only a forced form accepts that
        env = {"FLK_SLOT_C": str(slots), "FLK_LAYOUT": {1: "slots", 2: "staged"}[layout]}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        f = Field2D(W, H, D, True, capacity=cap or max(len(birds), 1))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if layout != -1 and slots is None:
        f.set_layout(layout)
    f.set_objects(*to_arrays(birds))
    f.lazy_update()
    return f


def state(f):
    return from_arrays(*f.snapshot_by_id())


@pytest.mark.parametrize("layout", [2, 1])
def test_uniform_field_uses_slots_and_matches_oracle(layout):
    W = 100.0
    birds = uniform_birds(1000, W, W)
    f = field(birds, W, W, layout=layout)
    lay, C, tot = f.layout()
    assert lay == layout and C == 16 and tot[3] == 0
    sim = O.Sim(W, W)
    sim.init(birds)
    for t in range(40):
        sim.step()
        f.run(t, 1)
        assert bits_equal(state(f), sim.agents()), t
    cs_g, ids_g = f.binning()
    dump, cs_o = sim.field.dump()
    assert np.array_equal(cs_g, cs_o) and np.array_equal(ids_g, dump["id"])


@pytest.mark.parametrize("n,W", [(20000, 450.0), (150000, 1230.0)])
def test_slots_equal_runs_and_oracle_mid_size(n, W):
    birds = uniform_birds(n, W, W)
    fs, fr, f2 = field(birds, W, W, layout=1), field(birds, W, W), field(birds, W, W, layout=2)
    assert fs.layout()[0] == 1 and fr.layout()[0] == 0 and f2.layout()[0] == 2      # automatic = the classic form
    sim = O.Sim(W, W)
    sim.init(birds)
    for _ in range(6):
        sim.step(threads=O.max_threads())
    for f in (fs, fr, f2):
        f.run(0, 6)
    a = sim.agents()
    assert bits_equal(state(fr), a)
    assert bits_equal(state(fs), a)
    assert bits_equal(state(f2), a)
    assert f2.layout()[0] == 2


@pytest.mark.parametrize("layout", [1, 2])
@pytest.mark.parametrize("slots", [4, 8])
def test_overflowing_cells_spill_and_stay_exact(slots, layout):
    """4 slots per cell at 4.4 birds per cell: about half the cells overflow into the spill list"""
    W = 200.0
    birds = uniform_birds(4000, W, W)
    f = field(birds, W, W, layout=layout, slots=slots)
    lay, C, tot = f.layout()
    assert lay == layout and C == slots
    if slots == 4 and layout == 1:
        assert tot[3] > 100 and tot[2] > 4
    sim = O.Sim(W, W)
    sim.init(birds)
    for t in range(8):
        sim.step()
        f.run(t, 1)
        assert bits_equal(state(f), sim.agents()), t
    lay, C, tot = f.layout()
    assert lay == layout
    if slots == 4:
        assert tot[3] > 100 and tot[2] > 4          # the ticks themselves overflowed cells
    from examples_b200 import Real2D
    for (x, y) in [(100.0, 100.0), (0.2, 199.9)]:
        assert np.array_equal(f.get_neighbors_within_relax_distance(Real2D(x, y), 10.0), sim.field.neighbors(x, y)["id"])


@pytest.mark.parametrize("layout", [1, 2])
def test_clustered_population_leaves_the_slot_forms(layout):
    """flocks far denser than the slots of a cell: a requested slot form is not entered (the spill list would hold
    most birds)"""
    W = 200.0
    birds = clustered_birds(5000, W, W)
    f = field(birds, W, W, layout=layout)
    assert f.layout()[0] == 0
    sim = O.Sim(W, W)
    sim.init(birds)
    for t in range(3):
        sim.step()
        f.run(t, 1)
    assert bits_equal(state(f), sim.agents())


@pytest.mark.parametrize("layout", [1, 2])
def test_graph_replays_from_both_parities(layout):
    W = 300.0
    birds = uniform_birds(9000, W, W)
    f = field(birds, W, W, layout=layout)
    assert f.layout()[0] == layout
    sim = O.Sim(W, W)
    sim.init(birds)
    t = 0
    for k in (9, 8, 11, 1, 16):       # odd counts flip the parity the next graph starts from
        f.run(t, k)
        for _ in range(k):
            sim.step()
        t += k
        assert bits_equal(state(f), sim.agents()), (t, k)


@pytest.mark.parametrize("layout", [1, 2])
def test_births_and_double_update_on_slots(layout):
    """set_object_location after the agents stepped and before update lands in the slotted WRITE buffer next to the
    successors; a second lazy_update empties the field"""
    W = 100.0
    birds = uniform_birds(600, W, W)
    f = field(birds, W, W, cap=1000, layout=layout)
    assert f.layout()[0] == layout
    sim = O.Sim(W, W)
    sim.init(birds)
    sim.field.lazy_update()
    next_id = 600
    rng = np.random.default_rng(8)
    for t in range(6):
        newb = np.zeros(40, dtype=O.BIRD)
        newb["id"] = np.arange(next_id, next_id + 40)
        newb["x"] = (rng.random(40) * W).astype(np.float32)
        newb["y"] = (rng.random(40) * W).astype(np.float32)
        next_id += 40
        assert len(sim.step_strip(t, 0xB12D, None, 0, sim.field.mx)) == 0
        sim.add_agents(newb)
        sim.field.lazy_update()
        f.step(t, 0xB12D)
        f.set_objects(*to_arrays(newb))
        f.lazy_update()
        assert f.layout()[0] == layout and f.num_objects() == next_id
        assert bits_equal(state(f), sim.agents()), t
    f.lazy_update()
    assert f.num_objects() == 0


def test_parameter_change_moves_the_field_to_sorted_runs():
    """a 5 x 5 window is not what the slotted step kernel is written for: flk_set_params rebuilds the sorted runs"""
    W = 160.0
    n = 2500
    birds = uniform_birds(n, W, W)
    f = field(birds, W, W, layout=1)
    assert f.layout()[0] == 1
    sim = O.Sim(W, W)
    sim.init(birds)
    rng = np.random.default_rng(3)
    for t in range(3):
        r = rng.random((n, 2), dtype=np.float32)
        sim.step(injected=r)
        f.step(t, injected=r)
        f.lazy_update()
    assert bits_equal(state(f), sim.agents())
    f.set_params(radius=15.0)
    assert f.layout()[0] != 1
    sim2 = O.Sim(W, W, cfg=O.default_cfg(radius=15.0))
    sim2.init(sim.agents())
    for t in range(3, 6):
        r = rng.random((n, 2), dtype=np.float32)
        sim2.step(injected=r)
        f.step(t, injected=r)
        f.lazy_update()
        assert bits_equal(state(f), sim2.agents()), t


@pytest.mark.parametrize("layout", [1, 2])
def test_dense_roundtrip_on_slots(layout):
    W = 250.0
    n = 6000
    birds = uniform_birds(n, W, W)
    f = field(birds, W, W, layout=layout)
    assert f.layout()[0] == layout
    rec = f.snapshot_dense(n)
    ids, loc, ld = f.snapshot_by_id()
    assert np.array_equal(rec[:, :2], loc) and np.array_equal(rec[:, 2:], ld)
    sim = O.Sim(W, W)
    sim.init(birds)
    for t in range(4):
        f.set_objects_dense(rec)                   # the e2e loop: upload everything, publish, tick, download
        f.lazy_update()
        f.run(t, 1)
        rec = f.snapshot_dense(n)
        sim.step()
        a = sim.agents()
        assert np.array_equal(rec[:, 0].view(np.uint32), a["x"].view(np.uint32)), t
        assert np.array_equal(rec[:, 3].view(np.uint32), a["ldy"].view(np.uint32)), t
    assert f.layout()[0] == layout


def test_one_million_birds_slots_equal_runs():
    W = 3200.0
    birds = uniform_birds(1_000_000, W, W)
    fs, fr, f2 = field(birds, W, W, layout=1), field(birds, W, W, layout=0), field(birds, W, W, layout=2)
    for f in (fs, fr, f2):
        f.run(0, 30)
    assert state_digest(state(fs)) == state_digest(state(fr)) == state_digest(state(f2))
    assert fs.layout()[0] == 1 and f2.layout()[0] == 2 and fs.layout()[2][0] == 1_000_000
