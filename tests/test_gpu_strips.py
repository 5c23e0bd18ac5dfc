"""GPU tests of the strip decomposition (flockers_mpi protocol) on ONE device: n strips driven in-process
(flk_mgpu_*) must reproduce the single-field run bit for bit — same Philox keys (id, step) and the same
canonical (cell, id) summation order make the result independent of which strip owns a bird."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import D, bits_equal, clustered_birds, from_arrays, to_arrays, uniform_birds

pytestmark = pytest.mark.gpu


def single(birds, W, H):
    from examples_b200 import Field2D
    f = Field2D(W, H, D, True, capacity=len(birds))
    f.set_objects(*to_arrays(birds))
    f.lazy_update()
    return f


def group(birds, W, H, n, cap=None):
    from examples_b200.distributed import MgpuGroup
    g = MgpuGroup(W, H, D, capacity_per_rank=cap or len(birds), devices=[0] * n)
    g.set_objects(*to_arrays(birds))
    g.commit()
    return g


@pytest.mark.parametrize("n", [2, 3, 4])
def test_strips_match_single_field_shipped_config(n):
    W = 100.0
    birds = uniform_birds(1000, W, W)
    f, g = single(birds, W, W), group(birds, W, W, n)
    assert sum(g.num_owned()) == 1000
    for t0 in range(0, 60, 10):
        f.run(t0, 10)
        g.run(t0, 10)
        assert bits_equal(from_arrays(*g.snapshot_by_id()), from_arrays(*f.snapshot_by_id())), (n, t0)
    assert sum(g.num_owned()) == 1000


def test_strips_match_oracle_with_injected_draws():
    W = 500.0
    n = 20000
    birds = uniform_birds(n, W, W)
    g = group(birds, W, W, 4)
    sim = O.Sim(W, W)
    sim.init(birds)
    rng = np.random.default_rng(9)
    for t in range(6):
        r = rng.random((n, 2), dtype=np.float32)
        sim.step(injected=r)
        g.step(t, injected=r)
        assert bits_equal(from_arrays(*g.snapshot_by_id()), sim.agents()), t


def test_strip_binning_is_the_global_binning_restricted_to_its_columns():
    from examples_b200.distributed import strip_of
    W = 300.0
    birds = uniform_birds(6000, W, W)
    g = group(birds, W, W, 3)
    g.run(0, 5)
    f = single(birds, W, W)
    f.run(0, 5)
    cs, ids = f.binning()
    dh = f.dh
    for r, sf in enumerate(g.fields):
        c0, c1 = strip_of(f.mx, 3, r)
        assert sf.strip() == (c0, c1)
        lcs, lids = sf.binning()
        # local column l (1..nown) is global column c0+l-1; ghosts: local 0 = c0-1, local nown+1 = c1 (wrapped)
        for l in range(0, c1 - c0 + 2):
            gc = (c0 + l - 1) % f.mx
            want = ids[cs[gc * dh]:cs[(gc + 1) * dh]]
            got = lids[lcs[l * dh]:lcs[(l + 1) * dh]]
            assert np.array_equal(got, want), (r, l)
            assert np.array_equal(np.diff(lcs[l * dh:(l + 1) * dh + 1]), np.diff(cs[gc * dh:(gc + 1) * dh + 1]))


def test_strips_clustered_heavy_migration():
    W = 300.0
    birds = clustered_birds(12000, W, W, per_cluster=600, sigma=8.0)
    f, g = single(birds, W, W), group(birds, W, W, 4)
    for t0 in range(0, 30, 10):
        f.run(t0, 10)
        g.run(t0, 10)
        assert bits_equal(from_arrays(*g.snapshot_by_id()), from_arrays(*f.snapshot_by_id())), t0


@pytest.mark.parametrize("n", [2, 5, 8])
def test_strips_long_unsynchronised_runs(n):
    """100 ticks enqueued back to back (no host synchronisation in between): exercises the event ordering between the
    main and the comm stream of every strip (send/receive buffer reuse across ticks)"""
    W = 400.0
    birds = uniform_birds(16000, W, W)
    f, g = single(birds, W, W), group(birds, W, W, n)
    for rep in range(3):
        f.run(100 * rep, 100)
        g.run(100 * rep, 100)
        assert bits_equal(from_arrays(*g.snapshot_by_id()), from_arrays(*f.snapshot_by_id())), (n, rep)


def test_strips_slack_column_and_seam():
    """a bird landing exactly on x == W (slack column, owned by the strip holding column 0) and birds crossing
    the seam between the last and the first strip"""
    W = 100.0
    b = np.zeros(8, dtype=O.BIRD)
    b["id"] = np.arange(8)
    b["x"] = [0.0, 0.5, 99.5, 99.99, 1.0, 46.6, 46.7, 53.3]
    b["y"] = [50.0, 50.5, 49.5, 50.0, 50.0, 70.0, 70.5, 70.2]
    b["ldx"] = [-1.0, -0.5, 0.7, 0.7, 0.0, 0.7, -0.7, -0.7]
    b["ldy"] = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.3, 0.0]
    for kw in (dict(cohesion=0.0, avoidance=0.0, randomness=0.0, consistency=0.0, momentum=1.0, jump=1e-6), dict()):
        f, g = single(b, W, W), group(b, W, W, 2, cap=64)
        f.set_params(**kw)
        g.set_params(**kw)
        for t in range(12):
            f.step(t)
            f.lazy_update()
            g.step(t)
            assert bits_equal(from_arrays(*g.snapshot_by_id()), from_arrays(*f.snapshot_by_id())), (kw, t)
            assert sum(g.num_owned()) == 8


def test_capacity_overflow_after_exchange_is_reported():
    """ghosts arriving over the halo can push a strip past its capacity: detected on the device, reported as
    FLK_E_CAPACITY at the next synchronising call instead of corrupting memory"""
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
        assert (own == r).sum() <= 21000
        g.fields[r].set_objects(*to_arrays(b[own == r]))
    g.commit()
    with pytest.raises(FlkError) as e:
        g.num_owned()
    assert e.value.code == -3


def test_population_balanced_strips_match_single_field():
    """Kdtree-style boundaries (flk_balance_columns): clustered flocks, uneven strip widths, same result bit for bit"""
    from examples_b200.distributed import MgpuGroup, balanced_columns, grid_columns
    W, H = 600.0, 300.0
    n = 12000
    birds = clustered_birds(n, W, H, per_cluster=1500, sigma=12.0)
    birds["x"] = (birds["x"] * np.float32(0.5)).astype(np.float32)       # everything in the left half of the field
    ids, loc, ld = to_arrays(birds)
    bounds = balanced_columns(loc[:, 0], W, 4)
    mx = grid_columns(W)
    assert bounds[0] == 0 and bounds[-1] == mx and np.all(np.diff(bounds) >= 2)
    assert bounds[3] <= mx // 2 + 2                                      # three of the four strips inside the left half
    f = single(birds, W, H)
    g = MgpuGroup(W, H, D, capacity_per_rank=n, devices=[0] * 4, col_bounds=bounds)
    assert [fld.strip() for fld in g.fields] == [(int(bounds[r]), int(bounds[r + 1])) for r in range(4)]
    g.scatter_objects(ids, loc, ld)
    g.commit()
    owned0 = g.num_owned()
    assert sum(owned0) == n and max(owned0) < 0.45 * n                   # equal widths would give strip 0 about half
    for t0 in range(0, 40, 10):
        f.run(t0, 10)
        g.run(t0, 10)
        assert bits_equal(from_arrays(*g.snapshot_by_id()), from_arrays(*f.snapshot_by_id())), t0
    for r, fld in enumerate(g.fields):                                   # get_block_by_location follows the boundaries
        c0, c1 = fld.strip()
        assert fld.owner((c0 + 0.5) * D, 1.0) == r and fld.owner((c1 - 0.5) * D, 1.0) == r
    assert g.fields[2].owner(W, 0.0) == 0                                # the slack column belongs to strip 0
    g.close()


def test_bad_strip_boundaries_are_rejected():
    from examples_b200 import _lib
    from examples_b200.distributed import MgpuGroup, grid_columns
    W = 200.0
    mx = grid_columns(W)
    for bad in ([0, 1, mx], [0, mx // 2, mx - 1], [1, mx // 2, mx], [0, mx, mx]):
        with pytest.raises(_lib.FlkError):
            MgpuGroup(W, W, D, capacity_per_rank=100, devices=[0, 0], col_bounds=bad)


def test_async_owned_snapshot_equals_the_synchronous_one():
    """flk_dist_snapshot_owned_async: exact count on return, records valid after synchronize (pinned buffers)"""
    import ctypes as C
    from examples_b200 import _lib
    L = _lib.load()
    W = 300.0
    n = 5000
    birds = uniform_birds(n, W, W)
    g = group(birds, W, W, 3)
    g.run(0, 7)
    for f in g.fields:
        ids, loc, ld = f.snapshot_owned()
        k = len(ids)
        ptrs = []
        for nbytes in (4 * n, 8 * n, 8 * n):
            p = C.c_void_p()
            _lib.check(L.flk_host_alloc(C.byref(p), nbytes))
            ptrs.append(p)
        try:
            got = f.snapshot_owned_async_raw(ptrs[0].value, ptrs[1].value, ptrs[2].value)
            assert got == k
            f.synchronize()
            a_id = np.ctypeslib.as_array(C.cast(ptrs[0], C.POINTER(C.c_uint32)), shape=(n,))[:k]
            a_loc = np.ctypeslib.as_array(C.cast(ptrs[1], C.POINTER(C.c_float)), shape=(n, 2))[:k]
            a_ld = np.ctypeslib.as_array(C.cast(ptrs[2], C.POINTER(C.c_float)), shape=(n, 2))[:k]
            assert np.array_equal(a_id, ids) and bits_equal(np.array(a_loc), loc) and bits_equal(np.array(a_ld), ld)
        finally:
            for p in ptrs:
                L.flk_host_free(p)
    g.close()


# ---- tiles: columns AND rows split (flk_mgpu_create_tiles), two-hop diagonal exchange ---------------------------------

def tiles(birds, W, H, n, rows, cap=None, **kw):
    from examples_b200.distributed import MgpuGroup
    g = MgpuGroup(W, H, D, capacity_per_rank=cap or len(birds), devices=[0] * n, tile_rows=rows, **kw)
    g.scatter_objects(*to_arrays(birds))
    g.commit()
    return g


@pytest.mark.parametrize("n,rows,W,H,count", [(4, 2, 200.0, 200.0, 4000), (6, 2, 300.0, 160.0, 5000),
                                               (6, 3, 200.0, 240.0, 5000), (9, 3, 260.0, 260.0, 7000)])
def test_tiles_match_single_field(n, rows, W, H, count):
    birds = uniform_birds(count, W, H)
    f, g = single(birds, W, H), tiles(birds, W, H, n, rows)
    assert sum(g.num_owned()) == count
    geo = [fld.tile() for fld in g.fields]
    assert len(set(geo)) == n                                            # n different tiles
    for t0 in range(0, 60, 10):
        f.run(t0, 10)
        g.run(t0, 10)
        assert bits_equal(from_arrays(*g.snapshot_by_id()), from_arrays(*f.snapshot_by_id())), (n, rows, t0)
    assert sum(g.num_owned()) == count
    for r, fld in enumerate(g.fields):                                   # get_block_by_location on the tile grid
        c0, c1, r0, r1 = fld.tile()
        assert fld.owner((c0 + 0.5) * D, (r0 + 0.5) * D) == r and fld.owner((c1 - 0.5) * D, (r1 - 0.5) * D) == r
    g.close()


def test_tiles_match_oracle_with_injected_draws_and_seam_birds():
    W = H = 160.0
    n = 3000
    birds = uniform_birds(n, W, H)
    # birds on the seam lines, in the corners and in the slack cells (x == W, y == H)
    birds["x"][:6] = [0.0, W, W, 0.0, 80.0, W]
    birds["y"][:6] = [0.0, H, 0.0, H, H, 80.0]
    g = tiles(birds, W, H, 4, 2)
    sim = O.Sim(W, H)
    sim.init(birds)
    rng = np.random.default_rng(5)
    for t in range(8):
        r = rng.random((n, 2), dtype=np.float32)
        sim.step(injected=r)
        g.step(t, injected=r)
        assert bits_equal(from_arrays(*g.snapshot_by_id()), sim.agents()), t
    g.close()


def test_tiles_clustered_with_uneven_boundaries():
    from examples_b200.distributed import balanced_columns, grid_columns
    W, H = 400.0, 300.0
    n = 9000
    birds = clustered_birds(n, W, H, per_cluster=900, sigma=10.0)
    ids, loc, ld = to_arrays(birds)
    cb = balanced_columns(loc[:, 0], W, 3)
    rb = balanced_columns(loc[:, 1], H, 2)
    f = single(birds, W, H)
    g = tiles(birds, W, H, 6, 2, col_bounds=cb, row_bounds=rb)
    assert g.fields[4].tile() == (int(cb[1]), int(cb[2]), int(rb[1]), int(rb[2]))
    for t0 in range(0, 30, 10):
        f.run(t0, 10)
        g.run(t0, 10)
        assert bits_equal(from_arrays(*g.snapshot_by_id()), from_arrays(*f.snapshot_by_id())), t0
    g.close()


@pytest.mark.parametrize("tile_rows", [1, 2])
def test_payload_travels_through_the_halo_and_migration_messages(tile_rows):
    """Field2D<O> with an object larger than Bird (template/src/model/crab.rs:11-19): its extra bytes follow it through
    Bird::step, through the strip / tile exchange (migrants AND ghosts, two hops on tiles) and through lazy_update"""
    from examples_b200 import Field2D
    from examples_b200.distributed import MgpuGroup
    from tests.helpers import clustered_birds
    W = 100.0
    n = 3000
    birds = clustered_birds(n, W, W, seed=21, per_cluster=150, sigma=6.0)        # headings: heavy migration
    ids, loc, ld = to_arrays(birds)
    pay = np.stack([ids * np.uint32(2654435761) + np.uint32(17), ~ids, ids ^ np.uint32(0xA5A5A5A5)], axis=1).astype(np.uint32)
    single = Field2D(W, W, D, True, capacity=n, payload_words=3)
    single.set_objects(ids, loc, ld, pay)
    single.lazy_update()
    g = MgpuGroup(W, W, D, capacity_per_rank=n, devices=[0, 0, 0, 0], tile_rows=tile_rows, payload_words=3)
    g.set_objects(ids, loc, ld, pay)
    g.commit()
    for t in range(25):
        single.step(t)
        single.lazy_update()
        g.step(t)
    s_id, s_loc, s_ld, s_pay = single.snapshot_with_payload()
    o = np.argsort(s_id, kind="stable")
    g_id, g_loc, g_ld, g_pay = g.snapshot_by_id_with_payload()
    assert np.array_equal(g_id, ids) and np.array_equal(s_id[o], ids)
    assert np.array_equal(g_pay, pay), "a payload was lost or attached to the wrong object on the way"
    assert np.array_equal(s_pay[o], pay)
    assert bits_equal(from_arrays(g_id, g_loc, g_ld), from_arrays(s_id[o], s_loc[o], s_ld[o]))
    moved = np.abs(g_loc - loc).max()
    assert moved > 5.0                                                            # birds did cross strip boundaries
