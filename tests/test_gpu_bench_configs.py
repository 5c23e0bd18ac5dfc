"""GPU parity on exactly the configurations bench.py measures (VERDICT r01 "what's weak" 1): the config-4 geometry
(16 M birds on 672 x 5376 cells), the clustered generator of config 5 at full size (dense cells: the window-column runs
exceed the shared-memory tile, so CTAs leave the packed tile loop), and a mid-size clustered field that forces both
regimes into one launch.  Whole populations, bit for bit against the threaded oracle, through the C ABI."""
import numpy as np
import pytest

import bench
from oracle import oracle as O
from tests.helpers import D, bits_equal, from_arrays, to_arrays

pytestmark = pytest.mark.gpu


def birds_from_loc(loc, headings_seed=None):
    n = len(loc)
    b = np.zeros(n, dtype=O.BIRD)
    b["id"] = np.arange(n, dtype=np.uint32)
    b["x"], b["y"] = loc[:, 0], loc[:, 1]
    if headings_seed is not None:
        h = np.random.default_rng(headings_seed).normal(0, 0.4, (n, 2)).astype(np.float32)
        b["ldx"], b["ldy"] = h[:, 0], h[:, 1]
    return b


def run_both(birds, W, H, ticks, first_step=0):
    from examples_b200 import Field2D
    f = Field2D(W, H, D, True, capacity=len(birds))
    f.set_objects(*to_arrays(birds))
    f.lazy_update()
    sim = O.Sim(W, H)
    sim.init(birds)
    for _ in range(ticks):
        sim.step(seed=bench.STEP_SEED, threads=O.max_threads())
    f.run(first_step, ticks, bench.STEP_SEED)
    g = from_arrays(*f.snapshot_by_id())
    return f, g, sim


def occupancy(f):
    cs, _ = f.binning()
    return np.diff(cs.astype(np.int64))


def test_config4_geometry_16m_birds_bit_exact():
    """bench.py's default workload at N=1: 16 M birds, 672 x 5376 cells, uniform random (BASELINE config 4)"""
    gx, gy, n, _ = bench.workload_geometry("weak16m", 1, None)
    W, H = bench.field_size(gx), bench.field_size(gy)
    birds = birds_from_loc(bench.make_birds(n, W, H, False))
    f, g, sim = run_both(birds, W, H, 2)
    assert (f.mx, f.my) == (gx, gy)
    assert bits_equal(g, sim.agents())
    dump, cs_o = sim.field.dump()
    cs_g, ids_g = f.binning()
    assert np.array_equal(cs_g, cs_o) and np.array_equal(ids_g, dump["id"])


def test_config5_clustered_16m_birds_bit_exact():
    """bench.py --workload clustered at N=1 (BASELINE config 5): gaussian flocks, sigma = 10, dense cells"""
    gx, gy, n, _ = bench.workload_geometry("clustered", 1, None)
    W, H = bench.field_size(gx), bench.field_size(gy)
    birds = birds_from_loc(bench.make_birds(n, W, H, True), headings_seed=11)
    f, g, sim = run_both(birds, W, H, 1)
    occ = occupancy(f)
    assert occ.max() > 64, occ.max()                       # the regime the test is about: cells far denser than the mean
    assert bits_equal(g, sim.agents())


@pytest.mark.parametrize("n,cells,per_cluster", [(400_000, 96, 2048), (1_000_000, 240, 4096)])
def test_clustered_mid_size_mixes_tile_and_overflow_paths(n, cells, per_cluster):
    """flocks dense enough that many window-column runs exceed the 256-record tile, next to sparse regions that stay on
    the tile path: both regimes inside one launch, several ticks"""
    from tests.helpers import clustered_birds
    W = bench.field_size(cells)
    birds = clustered_birds(n, W, W, seed=3, per_cluster=per_cluster, sigma=10.0)
    f, g, sim = run_both(birds, W, W, 3)
    # a window column run = 3 + rows-of-the-CTA cells; with > 90 birds per cell somewhere, 3 cells already overflow 256
    assert occupancy(f).max() > 90
    assert bits_equal(g, sim.agents())


@pytest.mark.parametrize("relax", [True, False])
@pytest.mark.parametrize("radius", [0.0, -1.0])
def test_non_positive_radius_sees_no_neighbours(radius, relax):
    """both queries return nothing for dist <= 0, so Bird::step keeps only the momentum term and draws nothing
    (bird.rs:42; ADVICE r01: k_step used to enumerate the bird's own cell)"""
    from examples_b200 import Field2D, Real2D
    from tests.helpers import clustered_birds
    W = 100.0
    birds = clustered_birds(800, W, W, seed=9)             # non-zero headings, so the birds move
    f = Field2D(W, W, D, True, capacity=len(birds))
    f.set_params(radius=radius, relax=relax)
    f.set_objects(*to_arrays(birds))
    f.lazy_update()
    sim = O.Sim(W, W, cfg=O.default_cfg(radius=radius, relax=int(relax)))
    sim.init(birds)
    for t in range(3):
        sim.step()
        f.run(t, 1)
        assert bits_equal(from_arrays(*f.snapshot_by_id()), sim.agents()), t
    q = f.get_neighbors_within_relax_distance if relax else f.get_neighbors_within_distance
    assert len(q(Real2D(50.0, 50.0), radius)) == 0


def test_device_errors_surface_in_synchronize():
    """ERR_CAPACITY / ERR_EXCHANGE raised inside a tick are reported by flk_synchronize, not only by
    flk_dist_num_owned"""
    from examples_b200 import FlkError
    from examples_b200.distributed import MgpuGroup, column_of, owner_of_column
    W = 100.0
    n = 30000
    b = np.zeros(n, dtype=O.BIRD)
    b["id"] = np.arange(n)
    rng = np.random.default_rng(1)
    b["x"] = (46.0 + rng.random(n)).astype(np.float32)       # columns 6 and 7: both sides of the 2-strip boundary
    b["y"] = (rng.random(n) * 99).astype(np.float32)
    g = MgpuGroup(W, W, D, capacity_per_rank=21000, devices=[0, 0])
    own = np.array([owner_of_column(g.mx, 2, int(c)) for c in column_of(b["x"], D)])
    for r in range(2):
        g.fields[r].set_objects(*to_arrays(b[own == r]))
    g.commit()
    with pytest.raises(FlkError) as e:
        for fr in g.fields:
            fr.synchronize()
    assert e.value.code == -3
