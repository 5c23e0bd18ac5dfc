"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI, against the
CPU oracle on the same seeded inputs and against the committed golden fixtures.

Bar: bit-exact for binning, neighbour sets AND for positions/headings (the kernels evaluate the reference's
f32 operations one IEEE op at a time, in the oracle's canonical order).  The looser bound the north star
allows (max abs error <= 1e-5 * field width per step) is only used for the `fast` math mode.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import D, bits_equal, clustered_birds, from_arrays, to_arrays, uniform_birds

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def gpu_field(W, H, cap, **kw):
    from examples_b200 import Field2D
    return Field2D(W, H, D, True, capacity=cap, **kw)


def load_gpu(birds, W, H, cap=None):
    f = gpu_field(W, H, cap or max(len(birds), 1))
    f.set_objects(*to_arrays(birds))
    f.lazy_update()
    return f


def gpu_state(f):
    return from_arrays(*f.snapshot_by_id())


def oracle_field(birds, W, H):
    fo = O.Field(W, H)
    fo.set_objects(birds)
    fo.lazy_update()
    return fo


# ---- binning ---------------------------------------------------------------------------------------

@pytest.mark.parametrize("maker,n,W", [(uniform_birds, 1000, 100.0), (uniform_birds, 20000, 500.0),
                                       (clustered_birds, 5000, 200.0)])
def test_binning_bit_exact(maker, n, W):
    birds = maker(n, W, W)
    f = load_gpu(birds, W, W)
    fo = oracle_field(birds, W, W)
    assert (f.dw, f.dh, f.mx, f.my) == (fo.dw, fo.dh, fo.mx, fo.my)
    dump, cs_o = fo.dump()
    cs_g, ids_g = f.binning()
    assert np.array_equal(cs_g, cs_o)                      # cell_start identical
    assert np.array_equal(ids_g, dump["id"])               # (cell, ascending id) order identical
    ids, loc, ld = f.snapshot()
    assert bits_equal(from_arrays(ids, loc, ld), dump)


def test_binning_edge_positions():
    """x == W / y == H land in the slack row/col (E5); x just below W may round into it as well."""
    W = 100.0
    b = np.zeros(6, dtype=O.BIRD)
    b["id"] = np.arange(6)
    b["x"] = [0.0, 100.0, 99.99999, 50.0, 100.0, 6.6666665]
    b["y"] = [0.0, 50.0, 100.0, 99.99999, 100.0, 6.6666665]
    f = load_gpu(b, W, W)
    fo = oracle_field(b, W, W)
    cs_g, ids_g = f.binning()
    dump, cs_o = fo.dump()
    assert np.array_equal(cs_g, cs_o) and np.array_equal(ids_g, dump["id"])


def test_lazy_update_twice_empties_the_read_buffer():
    birds = uniform_birds(100, 100.0, 100.0)
    f = load_gpu(birds, 100.0, 100.0)
    assert f.num_objects() == 100
    f.lazy_update()
    assert f.num_objects() == 0
    assert len(f.get_neighbors_within_relax_distance(__import__("examples_b200").Real2D(50.0, 50.0), 10.0)) == 0


def test_empty_field_steps():
    f = gpu_field(100.0, 100.0, 16)
    f.lazy_update()
    f.step(0)
    f.lazy_update()
    assert f.num_objects() == 0


def test_capacity_is_enforced():
    from examples_b200 import FlkError
    f = gpu_field(100.0, 100.0, 10)
    with pytest.raises(FlkError) as e:
        f.set_objects(*to_arrays(uniform_birds(11, 100.0, 100.0)))
    assert e.value.code == -3


# ---- neighbour sets --------------------------------------------------------------------------------

@pytest.mark.parametrize("maker,n,W", [(uniform_birds, 1000, 100.0), (clustered_birds, 3000, 200.0)])
@pytest.mark.parametrize("relax", [True, False])
def test_neighbor_sets_bit_exact(maker, n, W, relax):
    birds = maker(n, W, W)
    f = load_gpu(birds, W, W)
    fo = oracle_field(birds, W, W)
    rng = np.random.default_rng(5)
    # every bird's own location + random probes incl. the seam and the far edge
    probes = np.concatenate([np.stack([birds["x"], birds["y"]], 1)[:400],
                             (rng.random((100, 2)) * W).astype(np.float32),
                             np.array([[0.0, 0.0], [W, W], [W, 3.0], [0.01, W - 0.01]], dtype=np.float32)])
    off, ids = f.query_batch(probes, 10.0, relax=relax)
    for q, (x, y) in enumerate(probes):
        want = fo.neighbors(x, y, 10.0, relax=relax)["id"]
        got = ids[int(off[q]):int(off[q + 1])]
        assert np.array_equal(got, want), (q, x, y)          # same ids in the same (query) order


def test_single_query_api_and_capacity_code():
    from examples_b200 import Real2D, _lib
    import ctypes as C
    birds = uniform_birds(1000, 100.0, 100.0)
    f = load_gpu(birds, 100.0, 100.0)
    fo = oracle_field(birds, 100.0, 100.0)
    for (x, y) in [(50.0, 50.0), (0.5, 99.5), (100.0, 100.0)]:
        assert np.array_equal(f.get_neighbors_within_relax_distance(Real2D(x, y), 10.0), fo.neighbors(x, y)["id"])
        assert np.array_equal(f.get_neighbors_within_distance(Real2D(x, y), 10.0), fo.neighbors(x, y, relax=False)["id"])
    assert len(f.get_neighbors_within_relax_distance(Real2D(50.0, 50.0), 0.0)) == 0     # dist <= 0 -> empty
    ids = np.zeros(4, dtype=np.uint32)
    n = C.c_uint32()
    rc = f._L.flk_get_neighbors_within_relax_distance(f._h, 50.0, 50.0, 10.0, C.c_void_p(ids.ctypes.data), 4, C.byref(n))
    want = fo.neighbors(50.0, 50.0)["id"]
    assert rc == _lib.E_CAPACITY and n.value == len(want) and np.array_equal(ids, want[:4])


def test_payload_travels_with_its_object():
    """Field2D<O> with an object that carries more than Bird (template's Crab, virusnetwork's NetNode): the extra
    words follow their id through set_object_location, Bird::step and lazy_update; the Bird fields are unaffected."""
    from examples_b200 import Field2D
    n, W = 5000, 300.0
    birds = uniform_birds(n, W, W)
    ids, loc, ld = to_arrays(birds)
    pay = np.stack([ids * 7 + 1, ids ^ 0xABCD, np.arange(n, dtype=np.uint32)[::-1]], axis=1).astype(np.uint32)
    f = Field2D(W, W, D, True, capacity=n, payload_words=3)
    f.set_objects(ids, loc, ld, payload=pay)
    f.lazy_update()
    plain = load_gpu(birds, W, W)
    for t0 in (0, 10):
        f.run(t0, 10)
        plain.run(t0, 10)
        gi, gl, gd, gp = f.snapshot_with_payload()
        assert np.array_equal(gp, pay[gi])                                   # payload still attached to its id
        o = np.argsort(gi, kind="stable")
        assert bits_equal(from_arrays(gi[o], gl[o], gd[o]), gpu_state(plain))   # and the flock is the same flock


def test_non_toroidal_field_queries_match_oracle():
    """Field2D::new(w, h, d, false) as in virusnetwork/src/main.rs:17: the query window is clamped, not wrapped"""
    from examples_b200 import Field2D
    W = 100.0
    birds = uniform_birds(800, W, W)
    f = Field2D(W, W, D, False, capacity=800)
    f.set_objects(*to_arrays(birds))
    f.lazy_update()
    fo = O.Field(W, W, toroidal=False)
    fo.set_objects(birds)
    fo.lazy_update()
    probes = np.array([[0.5, 0.5], [99.5, 99.5], [50.0, 50.0], [0.1, 50.0], [99.9, 3.0], [100.0, 100.0]], dtype=np.float32)
    for relax in (True, False):
        off, ids = f.query_batch(probes, 10.0, relax=relax)
        for q, (x, y) in enumerate(probes):
            want = fo.neighbors(x, y, 10.0, relax=relax)["id"]
            assert np.array_equal(ids[int(off[q]):int(off[q + 1])], want), (relax, q)


def test_get_and_get_location_by_id():
    """Field2D::get / get_location (visualization accessors): lookup by id in the READ buffer"""
    birds = uniform_birds(1000, 100.0, 100.0)
    f = load_gpu(birds, 100.0, 100.0)
    f.run(0, 3)
    ids, loc, ld = f.snapshot_by_id()
    for i in (0, 17, 999):
        b = f.get(i)
        assert b is not None and b.id == i
        assert (np.float32(b.loc.x), np.float32(b.loc.y)) == (loc[i, 0], loc[i, 1])
        assert (np.float32(b.last_d.x), np.float32(b.last_d.y)) == (ld[i, 0], ld[i, 1])
        assert f.get_location(i).x == b.loc.x
    assert f.get(5000) is None and f.get_location(123456) is None


# ---- Bird::step ------------------------------------------------------------------------------------

@pytest.mark.parametrize("maker,n,W", [(uniform_birds, 1000, 100.0), (clustered_birds, 4000, 200.0)])
@pytest.mark.parametrize("relax", [True, False])
def test_step_bit_exact_with_injected_draws(maker, n, W, relax):
    birds = maker(n, W, W)
    f = load_gpu(birds, W, W)
    f.set_params(relax=relax)
    sim = O.Sim(W, W, cfg=O.default_cfg(relax=int(relax)))
    sim.init(birds)
    rng = np.random.default_rng(11)
    for t in range(5):
        r = rng.random((n, 2), dtype=np.float32)
        sim.step(injected=r)
        f.step(t, injected=r)
        f.lazy_update()
        assert bits_equal(gpu_state(f), sim.agents()), f"tick {t}"


@pytest.mark.parametrize("relax", [True, False])
def test_step_on_a_non_toroidal_field_matches_oracle(relax):
    """Field2D::new(w, h, d, false): the query window is clamped at the field edge while Bird::step keeps calling
    toroidal_distance / toroidal_transform (bird.rs:54-55,137-138 do so whatever the field is) — birds on the rim."""
    from examples_b200 import Field2D
    W, n = 100.0, 1200
    birds = uniform_birds(n, W, W)
    birds["x"][:150] = np.float32(0.25)      # a column of birds on the left rim, some on the right and upper rims
    birds["x"][150:300] = np.float32(99.75)
    birds["y"][300:400] = np.float32(99.9)
    f = Field2D(W, W, D, False, capacity=n)
    f.set_params(relax=relax)
    f.set_objects(*to_arrays(birds))
    f.lazy_update()
    sim = O.Sim(W, W, toroidal=False, cfg=O.default_cfg(relax=int(relax)))
    sim.init(birds)
    rng = np.random.default_rng(5)
    for t in range(6):
        r = rng.random((n, 2), dtype=np.float32)
        sim.step(injected=r)
        f.step(t, injected=r)
        f.lazy_update()
        assert bits_equal(gpu_state(f), sim.agents()), f"tick {t}"


def test_trajectory_shipped_config_matches_golden():
    """BASELINE config 1 (1000 birds, 100x100, 200 ticks, Philox draws): every pinned tick bit-exact."""
    g = np.load(os.path.join(GOLDEN, "flockers_1000.npz"))
    birds = g["init"].view(O.BIRD).reshape(-1)
    f = load_gpu(birds, 100.0, 100.0)
    ticks = list(g["ticks"])
    t = 0
    for stop in ticks:
        f.run(t, stop - t, int(g["seed"]))
        t = stop
        assert bits_equal(gpu_state(f), g[f"state_{stop}"].view(O.BIRD).reshape(-1)), stop


def test_clustered_matches_golden_both_queries():
    g = np.load(os.path.join(GOLDEN, "flockers_clustered_4000.npz"))
    birds = g["init"].view(O.BIRD).reshape(-1)
    for prefix, relax in (("state_", True), ("exact_", False)):
        f = load_gpu(birds, 200.0, 200.0)
        f.set_params(relax=relax)
        t = 0
        for stop in list(g["ticks"]):
            f.run(t, stop - t, int(g["seed"]))
            t = stop
            assert bits_equal(gpu_state(f), g[f"{prefix}{stop}"].view(O.BIRD).reshape(-1)), (prefix, stop)


def test_graph_and_plain_launch_paths_agree():
    birds = uniform_birds(5000, 300.0, 300.0)
    a = load_gpu(birds, 300.0, 300.0)
    b = load_gpu(birds, 300.0, 300.0)
    b.use_graph(False)
    a.run(0, 25, 0xB12D)
    b.run(0, 25, 0xB12D)
    assert bits_equal(gpu_state(a), gpu_state(b))
    c = load_gpu(birds, 300.0, 300.0)
    for t in range(25):
        c.step(t, 0xB12D)
        c.lazy_update()
    assert bits_equal(gpu_state(a), gpu_state(c))


def test_edge_cases_match_oracle():
    """lone bird, seam/corner pair, dis==0, x==W slack cell (E5) and a NaN-producing draw."""
    W = 100.0
    b = np.zeros(8, dtype=O.BIRD)
    b["id"] = np.arange(8)
    b["x"] = [20.0, 0.5, 99.5, 0.0, 1.0, 70.0, 70.5, 40.0]
    b["y"] = [20.0, 0.5, 99.5, 50.0, 50.0, 70.0, 70.5, 80.0]
    b["ldx"] = [0.1, 0.0, 0.0, -1.0, 0.0, 0.2, -0.1, 0.0]
    b["ldy"] = [0.2, 0.0, 0.0, 0.0, 0.0, 0.0, 0.3, 0.0]
    for kw in (dict(), dict(cohesion=0.0, avoidance=0.0, randomness=0.0, consistency=0.0, momentum=1.0, jump=1e-6),
               dict(randomness=0.0)):
        f = load_gpu(b, W, W)
        f.set_params(**kw)
        sim = O.Sim(W, W, cfg=O.default_cfg(**kw))
        sim.init(b)
        rng = np.random.default_rng(2)
        for t in range(4):
            r = rng.random((8, 2), dtype=np.float32)
            sim.step(injected=r)
            f.step(t, injected=r)
            f.lazy_update()
            assert bits_equal(gpu_state(f), sim.agents()), (kw, t)
    # bird 7 draws r=(0.5,0.5): 0/0 -> NaN position, exactly like the reference arithmetic
    f = load_gpu(b, W, W)
    sim = O.Sim(W, W)
    sim.init(b)
    r = np.full((8, 2), 0.25, dtype=np.float32)
    r[7] = 0.5
    sim.step(injected=r)
    f.step(0, injected=r)
    f.lazy_update()
    g, o = gpu_state(f), sim.agents()
    assert np.isnan(g["x"][7]) and np.isnan(o["x"][7])       # NaN payload/sign bits are not part of the contract
    assert bits_equal(g[:7], o[:7])


def test_fast_math_mode_within_stated_tolerance():
    """fast mode (reciprocal division): max abs error <= 1e-5 * W per step against the exact oracle step."""
    W = 500.0
    birds = uniform_birds(20000, W, W)
    f = load_gpu(birds, W, W)
    f.set_math_mode(1)
    sim = O.Sim(W, W)
    sim.init(birds)
    sim.step()
    f.step(0)
    f.lazy_update()
    g, o = gpu_state(f), sim.agents()
    assert np.abs(g["x"] - o["x"]).max() <= 1e-5 * W and np.abs(g["y"] - o["y"]).max() <= 1e-5 * W
    assert np.abs(g["ldx"] - o["ldx"]).max() <= 1e-5 * W


def test_division_helper_is_ieee():
    """the hot loop's shared-reciprocal division equals __fdiv_rn bit for bit on 2^28 operand pairs
    (normal, zero, denormal numerators; denominators across the fast-path guard)."""
    import ctypes as C
    from examples_b200 import _lib
    L = _lib.load()
    bad = C.c_uint64(123)
    _lib.check(L.flk_selftest_div2(0, 1 << 28, 0xD1F, C.byref(bad)))
    assert bad.value == 0


# ---- full-size properties (BASELINE sizes; the oracle is only sampled) ---------------------------------

@pytest.mark.parametrize("n,W,ticks", [(1_000_000, 3200.0, 3), (16_000_000, 12800.0, 2)])
def test_full_size_properties(n, W, ticks):
    """BASELINE configs 2 and 3 (1 M birds on 480 x 480 cells, 16 M birds on 1920 x 1920 cells)"""
    birds = uniform_birds(n, W, W)
    f = load_gpu(birds, W, W)
    cs, ids = f.binning()
    assert cs[-1] == n and np.all(np.diff(cs.astype(np.int64)) >= 0)
    assert np.array_equal(np.sort(ids), np.arange(n, dtype=np.uint32))              # a permutation of the ids
    ids_s, loc, ld = f.snapshot()
    cx = np.floor(loc[:, 0] / np.float32(D)).astype(np.int64)
    cy = np.floor(loc[:, 1] / np.float32(D)).astype(np.int64)
    key = cx * f.dh + cy
    assert np.all(np.diff(key) >= 0)                                                # sorted by cell
    same = np.diff(key) == 0
    assert np.all(np.diff(ids_s.astype(np.int64))[same] > 0)                        # ascending id inside a cell
    assert np.array_equal(np.bincount(key, minlength=f.dw * f.dh), np.diff(cs.astype(np.int64)))
    # a few ticks on the GPU vs the threaded oracle, whole population, bit for bit
    sim = O.Sim(W, W)
    sim.init(birds)
    for _ in range(ticks):
        sim.step(threads=O.max_threads())
    f.run(0, ticks)
    g = gpu_state(f)
    assert bits_equal(g, sim.agents())
    # every displacement has length JUMP (bird.rs:129-133) and positions stay on the torus
    sp = np.hypot(g["ldx"].astype(np.float64), g["ldy"].astype(np.float64))
    assert np.all(np.abs(sp - 0.7) < 1e-6)
    assert g["x"].min() >= 0 and g["x"].max() <= W and g["y"].min() >= 0 and g["y"].max() <= W


@pytest.mark.parametrize("relax", [True, False])
def test_queries_return_the_objects_themselves(relax):
    """bird.rs:29-31,52-71: the reference's queries return Vec<Bird> and Bird::step reads elem.loc / elem.last_d.
    flk_get_neighbors_within_*_objects returns the same objects in the same order, bit for bit"""
    import ctypes as C
    from examples_b200 import Real2D, _lib
    W = 200.0
    birds = clustered_birds(3000, W, W)                   # non-zero headings
    f = load_gpu(birds, W, W)
    fo = oracle_field(birds, W, W)
    rng = np.random.default_rng(12)
    probes = np.concatenate([np.stack([birds["x"], birds["y"]], 1)[:150], (rng.random((50, 2)) * W).astype(np.float32),
                             np.array([[0.0, 0.0], [W, W], [0.01, W - 0.01]], dtype=np.float32)])
    off, objs = f.query_batch_objects(probes, 10.0, relax=relax)
    assert objs.dtype.itemsize == 20
    for q, (x, y) in enumerate(probes):
        want = fo.neighbors(x, y, 10.0, relax=relax)
        got = objs[int(off[q]):int(off[q + 1])]
        assert got.tobytes() == np.ascontiguousarray(want).tobytes(), (q, x, y)
        if q % 40 == 0:
            one = f.get_neighbors_objects(Real2D(float(x), float(y)), 10.0, relax=relax)
            assert one.tobytes() == got.tobytes()
    b = f.get_neighbors_birds(Real2D(100.0, 100.0), 10.0, relax=relax)
    want = fo.neighbors(100.0, 100.0, 10.0, relax=relax)
    assert [o.id for o in b] == list(want["id"]) and all(o.last_d.x == w["ldx"] for o, w in zip(b, want))
    # capacity behaviour: full size reported, the first `cap` objects valid
    out = np.zeros(3, dtype=f.BIRD_DTYPE)
    n = C.c_uint32()
    fn = f._L.flk_get_neighbors_within_relax_distance_objects if relax else f._L.flk_get_neighbors_within_distance_objects
    rc = fn(f._h, 100.0, 100.0, 10.0, C.c_void_p(out.ctypes.data), 3, C.byref(n))
    assert n.value == len(want)
    if len(want) > 3:
        assert rc == _lib.E_CAPACITY and out.tobytes() == np.ascontiguousarray(want[:3]).tobytes()
