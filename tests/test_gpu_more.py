"""More GPU parity cases: odd geometries, other window widths, births between step and lazy_update, degenerate
populations.  Everything bit-exact against the oracle."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import D, bits_equal, from_arrays, to_arrays, uniform_birds

pytestmark = pytest.mark.gpu


def gpu(birds, W, H, cap=None, **params):
    from examples_b200 import Field2D
    f = Field2D(W, H, D, True, capacity=cap or max(len(birds), 1))
    if params:
        f.set_params(**params)
    f.set_objects(*to_arrays(birds))
    f.lazy_update()
    return f


def state(f):
    return from_arrays(*f.snapshot_by_id())


def make(n, W, H, seed=1):
    rng = np.random.default_rng(seed)
    b = np.zeros(n, dtype=O.BIRD)
    b["id"] = np.arange(n)
    b["x"] = (rng.random(n) * W).astype(np.float32)
    b["y"] = (rng.random(n) * H).astype(np.float32)
    b["x"] = np.minimum(b["x"], np.nextafter(np.float32(W), np.float32(0)))
    b["y"] = np.minimum(b["y"], np.nextafter(np.float32(H), np.float32(0)))
    return b


@pytest.mark.parametrize("W,H,n", [(20.0, 20.0, 60),        # 3 x 3 cells: every window is the whole field
                                    (100.0, 37.0, 700),      # non-square, H not a multiple of d (partial last row)
                                    (53.0, 211.0, 1500),     # partial last column
                                    (640.0, 46.7, 4000)])    # wide and thin: 96 x 8 cells
def test_odd_geometries(W, H, n):
    birds = make(n, W, H)
    f = gpu(birds, W, H)
    sim = O.Sim(W, H)
    sim.init(birds)
    fo = sim.field
    assert (f.dw, f.dh, f.mx, f.my) == (fo.dw, fo.dh, fo.mx, fo.my)
    for t in range(12):
        sim.step()
        f.run(t, 1)
        assert bits_equal(state(f), sim.agents()), (W, H, t)
    cs_g, ids_g = f.binning()
    dump, cs_o = sim.field.dump()
    assert np.array_equal(cs_g, cs_o) and np.array_equal(ids_g, dump["id"])


@pytest.mark.parametrize("radius,relax", [(15.0, True), (15.0, False), (5.0, True), (5.0, False), (20.5, True)])
def test_other_window_widths(radius, relax):
    """radius 15 -> 5x5 cells, radius 5 -> own cell only, radius 20.5 -> 7x7: the generic window walk"""
    W = 160.0
    birds = make(2500, W, W, seed=3)
    f = gpu(birds, W, W, radius=radius, relax=relax)
    sim = O.Sim(W, W, cfg=O.default_cfg(radius=radius, relax=int(relax)))
    sim.init(birds)
    for t in range(6):
        sim.step()
        f.run(t, 1)
        assert bits_equal(state(f), sim.agents()), (radius, relax, t)
    from examples_b200 import Real2D
    for (x, y) in [(80.0, 80.0), (0.3, 159.9), (159.9, 0.2)]:
        got = f.get_neighbors_within_relax_distance(Real2D(x, y), radius) if relax else \
            f.get_neighbors_within_distance(Real2D(x, y), radius)
        assert np.array_equal(got, sim.field.neighbors(x, y, radius, relax=relax)["id"])


def test_window_wider_than_grid_is_refused():
    from examples_b200 import Field2D, FlkError
    f = Field2D(30.0, 30.0, D, True, capacity=10)      # 5 x 5 cells
    with pytest.raises(FlkError) as e:
        f.set_params(radius=20.5)                       # 7 x 7 window
    assert e.value.code == -4


def test_other_weights_and_jump():
    W = 200.0
    birds = make(3000, W, W, seed=5)
    kw = dict(cohesion=0.3, avoidance=2.5, randomness=0.4, consistency=1.7, momentum=0.2, jump=3.1)
    f = gpu(birds, W, W, **kw)
    sim = O.Sim(W, W, cfg=O.default_cfg(**kw))
    sim.init(birds)
    for t in range(8):
        sim.step()
        f.run(t, 1)
        assert bits_equal(state(f), sim.agents()), t


def test_jump_larger_than_a_cell():
    """birds cross several cells per tick: nothing in the single-GPU path may assume adjacency"""
    W = 200.0
    birds = make(2000, W, W, seed=6)
    f = gpu(birds, W, W, jump=17.0)
    sim = O.Sim(W, W, cfg=O.default_cfg(jump=17.0))
    sim.init(birds)
    for t in range(6):
        sim.step()
        f.run(t, 1)
        assert bits_equal(state(f), sim.agents()), t


def test_births_between_step_and_lazy_update():
    """set_object_location after the agents stepped and before update: both land in the same WRITE buffer"""
    W = 100.0
    birds = uniform_birds(600, W, W)
    f = gpu(birds, W, W, cap=1000)
    sim = O.Sim(W, W)
    sim.init(birds)
    sim.field.lazy_update()
    next_id = 600
    rng = np.random.default_rng(8)
    for t in range(8):
        newb = np.zeros(40, dtype=O.BIRD)
        newb["id"] = np.arange(next_id, next_id + 40)
        newb["x"] = (rng.random(40) * W).astype(np.float32)
        newb["y"] = (rng.random(40) * W).astype(np.float32)
        next_id += 40
        # oracle: agents step (no update), births, update
        assert len(sim.step_strip(t, 0xB12D, None, 0, sim.field.mx)) == 0
        sim.add_agents(newb)
        sim.field.lazy_update()
        # GPU
        f.step(t, 0xB12D)
        f.set_objects(*to_arrays(newb))
        f.lazy_update()
        assert f.num_objects() == next_id
        assert bits_equal(state(f), sim.agents()), t


def test_everyone_in_one_cell_and_coincident_birds():
    """500 birds inside one cell (tile overflow -> L1 path, O(n^2) rank) including exactly coincident pairs"""
    W = 100.0
    rng = np.random.default_rng(9)
    b = np.zeros(500, dtype=O.BIRD)
    b["id"] = np.arange(500)
    b["x"] = (50.0 + rng.random(500) * 6.0).astype(np.float32)
    b["y"] = (50.0 + rng.random(500) * 6.0).astype(np.float32)
    b["x"][1::7] = b["x"][0::7][: len(b["x"][1::7])]      # coincident x and y (dx = dy = 0 for a pair)
    b["y"][1::7] = b["y"][0::7][: len(b["y"][1::7])]
    f = gpu(b, W, W)
    sim = O.Sim(W, W)
    sim.init(b)
    for t in range(5):
        sim.step()
        f.run(t, 1)
        assert bits_equal(state(f), sim.agents()), t


def test_strips_with_exact_query_and_other_parameters():
    from examples_b200.distributed import MgpuGroup
    W = 300.0
    birds = make(6000, W, W, seed=11)
    kw = dict(relax=False, cohesion=0.5, avoidance=1.5)
    f = gpu(birds, W, W, **kw)
    g = MgpuGroup(W, W, D, capacity_per_rank=6000, devices=[0, 0, 0])
    g.set_params(**kw)
    g.set_objects(*to_arrays(birds))
    g.commit()
    for t0 in (0, 7):
        f.run(t0, 7)
        g.run(t0, 7)
        assert bits_equal(from_arrays(*g.snapshot_by_id()), state(f)), t0


def test_strips_refuse_parameters_they_cannot_honour():
    from examples_b200 import FlkError
    from examples_b200.distributed import MgpuGroup
    g = MgpuGroup(300.0, 300.0, D, capacity_per_rank=100, devices=[0, 0])
    for kw in (dict(jump=7.0), dict(radius=15.0)):
        with pytest.raises(FlkError) as e:
            g.set_params(**kw)
        assert e.value.code == -4


# ---- dense (id-indexed) host interface and asynchronous snapshots ------------------------------------------------

def rec4_of(birds):
    return np.ascontiguousarray(np.stack([birds["x"], birds["y"], birds["ldx"], birds["ldy"]], axis=1))


def test_dense_upload_and_snapshot_match_the_id_list_forms():
    """flk_set_objects_dense / flk_snapshot_dense move Bird minus its id; results identical to the explicit-id forms
    and to the oracle over a few ticks"""
    from examples_b200 import Field2D
    W = H = 200.0
    birds = make(3000, W, H, seed=5)
    sim = O.Sim(W, H)
    sim.init(birds)
    f = Field2D(W, H, D, True, capacity=4000)
    f.set_objects_dense(rec4_of(birds))
    f.lazy_update()
    assert bits_equal(state(f), np.sort(birds, order="id"))
    for t in range(5):
        sim.step()
        f.run(t, 1)
        a = sim.agents()
        assert bits_equal(f.snapshot_dense(len(birds)), rec4_of(a)), t
    # feed the dense snapshot back in (host-held agents), continue: same trajectory
    r = f.snapshot_dense(len(birds))
    f.clear()
    f.set_objects_dense(r)
    f.lazy_update()
    sim.step()
    f.run(5, 1)
    assert bits_equal(f.snapshot_dense(len(birds)), rec4_of(sim.agents()))


def test_dense_snapshot_id_window_and_missing_ids():
    from examples_b200 import Field2D, _lib
    W = H = 100.0
    birds = make(500, W, H, seed=9)
    birds["id"] = np.arange(500) * 2 + 100          # ids 100, 102, ...: every other slot of the range is empty
    f = gpu(birds, W, H, cap=2000)
    r = f.snapshot_dense(1000, id0=100)
    assert bits_equal(r[0::2], rec4_of(birds))
    assert np.all(r[1::2].view(np.uint32) == 0xFFFFFFFF)
    with pytest.raises(_lib.FlkError) as e:         # ids outside the requested window are an error, not silence
        f.snapshot_dense(100, id0=100)
    assert e.value.code == _lib.E_INVALID
    with pytest.raises(_lib.FlkError) as e:
        f.snapshot_dense(5000)
    assert e.value.code == _lib.E_CAPACITY
    # dense upload with an id offset
    g = Field2D(W, H, D, True, capacity=600)
    g.set_objects_dense(rec4_of(birds), id0=7)
    g.lazy_update()
    ids, _, _ = g.snapshot_by_id()
    assert np.array_equal(ids, np.arange(500, dtype=np.uint32) + 7)


def test_async_snapshots_on_two_handles_overlap_safely():
    """two independent repetitions pipelined over two handles with pinned buffers: results equal the serial ones"""
    import ctypes as C
    from examples_b200 import Field2D, _lib
    L = _lib.load()
    W = H = 400.0
    n = 20000
    pops = [make(n, W, H, seed=s) for s in (21, 22)]
    want = []
    for p in pops:
        f = gpu(p, W, H)
        f.run(0, 3)
        want.append(f.snapshot_dense(n))
        f.close()
    fs = [Field2D(W, H, D, True, capacity=n) for _ in pops]
    bufs = []
    for _ in range(4):
        ptr = C.c_void_p()
        _lib.check(L.flk_host_alloc(C.byref(ptr), n * 16))
        bufs.append(ptr)
    try:
        def view(p):
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n, 4))
        for k, p in enumerate(pops):
            view(bufs[k])[:] = rec4_of(p)
        for t in range(3):
            for k, f in enumerate(fs):
                f.synchronize()                                        # the previous download of this handle is done
                f.set_objects_dense_raw(n, bufs[k].value)
                f.lazy_update()
                f.run(t, 1)
                f.snapshot_dense_async_raw(n, bufs[2 + k].value)
                bufs[k], bufs[2 + k] = bufs[2 + k], bufs[k]            # this tick's output is the next tick's input
        for f in fs:
            f.synchronize()
        for k in range(2):
            assert bits_equal(np.array(view(bufs[k])), want[k]), k
    finally:
        for f in fs:
            f.close()
        for p in bufs:
            L.flk_host_free(p)


def test_frames_download_beside_the_simulation():
    """flk_frame_begin / flk_frame_wait: the frame is the state at the time of the call even though later ticks are
    enqueued (and overwrite the READ buffer) before the download has finished"""
    import ctypes as C
    from examples_b200 import _lib
    L = _lib.load()
    W = H = 600.0
    n = 60000
    birds = make(n, W, H, seed=31)
    a, b = gpu(birds, W, H), gpu(birds, W, H)
    a.run(0, 10); b.run(0, 10)
    want_ids, want_loc, want_ld = b.snapshot()                 # (cell, id) order at tick 10
    p_id, p_rec = C.c_void_p(), C.c_void_p()
    _lib.check(L.flk_host_alloc(C.byref(p_id), 4 * n)); _lib.check(L.flk_host_alloc(C.byref(p_rec), 16 * n))
    try:
        for frame in range(2):                                  # the second begin queues behind the first
            got = a.frame_begin_raw(p_id.value, p_rec.value)
            assert got == n
            a.run(10 + 30 * frame, 30)                          # keeps the GPU busy and rewrites the READ buffer
            a.frame_wait()
            ids = np.ctypeslib.as_array(C.cast(p_id, C.POINTER(C.c_uint32)), shape=(n,)).copy()
            rec = np.ctypeslib.as_array(C.cast(p_rec, C.POINTER(C.c_float)), shape=(n, 4)).copy()
            if frame == 0:
                assert np.array_equal(ids, want_ids)
                assert bits_equal(rec[:, :2].copy(), want_loc) and bits_equal(rec[:, 2:].copy(), want_ld)
            else:
                b.run(10, 30)
                i2, l2, d2 = b.snapshot()
                assert np.array_equal(ids, i2) and bits_equal(rec[:, :2].copy(), l2) and bits_equal(rec[:, 2:].copy(), d2)
        a.synchronize()
    finally:
        L.flk_host_free(p_id); L.flk_host_free(p_rec)


@pytest.mark.parametrize("relax", [True, False])
def test_window_plus_one_switch_matches_the_restatement(relax):
    """[EXT] E3 of SURVEY.md §8c: if the real engine followed MASON (window = floor(dist/d) + 1 cells on either side),
    flk_set_engine_semantics(window_plus1=1) is the one switch to flip; the restatement carries the same switch"""
    from examples_b200 import Real2D
    W = 160.0
    birds = make(2500, W, W, seed=4)
    f = gpu(birds, W, W, relax=relax)
    f.set_engine_semantics(window_plus1=True)
    sim = O.Sim(W, W, cfg=O.default_cfg(relax=int(relax), window_plus1=1))
    sim.init(birds)
    for t in range(5):
        sim.step()
        f.run(t, 1)
        assert bits_equal(state(f), sim.agents()), (relax, t)
    for (x, y) in [(80.0, 80.0), (0.3, 159.9)]:
        got = f.get_neighbors_within_relax_distance(Real2D(x, y), 10.0) if relax else f.get_neighbors_within_distance(Real2D(x, y), 10.0)
        assert np.array_equal(got, sim.field.neighbors(x, y, 10.0, relax=relax, plus1=True)["id"])
    f.set_engine_semantics(window_plus1=False)
    assert len(f.get_neighbors_within_relax_distance(Real2D(80.0, 80.0), 10.0)) < len(sim.field.neighbors(80.0, 80.0, 10.0, plus1=True))
