"""Host-side mirror of the krABMaga surface the Flockers hot path is written against.

Names, argument meaning and call order follow the reference (krABMaga/examples):
  Field2D            krabmaga::engine::fields::field_2d::Field2D — call sites flockers/src/model/state.rs:24,49,55,
                     flockers/src/model/bird.rs:29-31,142-144
  Flocker (State)    flockers/src/model/state.rs:12-73
  Bird               flockers/src/model/bird.rs:13-24
  Schedule.step      krabmaga::engine::schedule::Schedule::step; hand-written
                     equivalent of the outer loop: forestfire_bayesian/src/main.rs:63-72
  FlockStepper       the repo's own idiom of ONE agent sweeping the whole field per tick
                     (schelling/src/model/updater.rs:17-80, forestfire/src/model/spread.rs:17-115)
  simulate           krabmaga::simulate_old! as used at flockers/src/main.rs:47

Every method forwards to libflk.so through ctypes (include/flk.h).  No numerics happen in Python.
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import check

# flockers/src/main.rs:26-33
COHESION = 0.8
AVOIDANCE = 1.0
RANDOMNESS = 1.1
CONSISTENCY = 0.7
MOMENTUM = 1.0
JUMP = 0.7
DISCRETIZATION = float(np.float32(10.0) / np.float32(1.5))
TOROIDAL = True


@dataclass
class Real2D:  # krabmaga::engine::location::Real2D
    x: float
    y: float


@dataclass
class Bird:  # flockers/src/model/bird.rs:13-18
    id: int
    loc: Real2D
    last_d: Real2D


def _vp(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Field2D:
    """GPU-resident Field2D<Bird>: READ buffer sorted by (cell, id), WRITE buffer in arrival order."""

    def __init__(self, width, height, discretization=DISCRETIZATION, toroidal=TOROIDAL, capacity=1 << 20, device=0,
                 payload_words=0):
        self._L = _lib.load()
        self._h = C.c_void_p()
        self.payload_words = payload_words   # Field2D<O>: extra 32-bit words every object carries (0 for Bird)
        check(self._L.flk_field_create_ex(width, height, discretization, int(toroidal), capacity, device,
                                          4 * payload_words, C.byref(self._h)))
        self.width, self.height, self.discretization, self.toroidal = width, height, discretization, toroidal
        self.capacity = capacity
        dw, dh, mx, my = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        check(self._L.flk_field_dims(self._h, C.byref(dw), C.byref(dh), C.byref(mx), C.byref(my)))
        self.dw, self.dh, self.mx, self.my = dw.value, dh.value, mx.value, my.value

    @classmethod
    def _adopt(cls, handle, width, height, discretization, capacity):
        self = cls.__new__(cls)
        self._L = _lib.load()
        self._h = handle
        self.payload_words = 0
        self.width, self.height, self.discretization, self.toroidal = width, height, discretization, True
        self.capacity = capacity
        dw, dh, mx, my = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        check(self._L.flk_field_dims(self._h, C.byref(dw), C.byref(dh), C.byref(mx), C.byref(my)))
        self.dw, self.dh, self.mx, self.my = dw.value, dh.value, mx.value, my.value
        return self

    def close(self):
        if getattr(self, "_h", None):
            self._L.flk_field_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration
    def set_params(self, cohesion=COHESION, avoidance=AVOIDANCE, randomness=RANDOMNESS, consistency=CONSISTENCY,
                   momentum=MOMENTUM, jump=JUMP, radius=10.0, relax=True):
        check(self._L.flk_set_params(self._h, cohesion, avoidance, randomness, consistency, momentum, jump, radius, int(relax)))

    def set_engine_semantics(self, window_plus1: bool = False):
        """[EXT] E3 of SURVEY.md §8c: query window of floor(dist/d) (+1 if window_plus1) cells on either side"""
        check(self._L.flk_set_engine_semantics(self._h, int(window_plus1)))

    def set_math_mode(self, mode: int):
        check(self._L.flk_set_math_mode(self._h, mode))

    def use_graph(self, on: bool):
        check(self._L.flk_use_graph(self._h, int(on)))

    def clear(self):
        check(self._L.flk_field_clear(self._h))

    def set_layout(self, layout: int):
        """-1 automatic, 0 sorted runs, 1 fixed-capacity cell slots (include/flk.h: flk_set_layout)"""
        check(self._L.flk_set_layout(self._h, layout))

    def layout(self):
        """(layout in use, slots per cell, [objects, largest column, largest cell, spilled objects])"""
        lay, c = C.c_int32(), C.c_int32()
        tot = (C.c_uint32 * 4)()
        check(self._L.flk_get_layout(self._h, C.byref(lay), C.byref(c), tot))
        return lay.value, c.value, list(tot)

    # ---- write side
    def set_object_location(self, bird: Bird, loc: Real2D):
        check(self._L.flk_set_object_location(self._h, bird.id, loc.x, loc.y, bird.last_d.x, bird.last_d.y))

    def set_objects(self, ids: np.ndarray, loc: np.ndarray, last_d: np.ndarray, payload: np.ndarray | None = None):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        loc = np.ascontiguousarray(loc, dtype=np.float32)
        last_d = np.ascontiguousarray(last_d, dtype=np.float32)
        n = len(ids)
        assert loc.size == 2 * n and last_d.size == 2 * n
        if payload is not None:
            payload = np.ascontiguousarray(payload, dtype=np.uint32)
            assert payload.size == n * getattr(self, "payload_words", 0)
            check(self._L.flk_set_objects_ex(self._h, n, _vp(ids), _vp(loc), _vp(last_d), _vp(payload)))
        else:
            check(self._L.flk_set_objects(self._h, n, _vp(ids), _vp(loc), _vp(last_d)))

    def set_objects_raw(self, n, id_ptr, loc_ptr, ld_ptr):
        """Pointer form (pinned torch tensors in bench.py)."""
        check(self._L.flk_set_objects(self._h, n, C.c_void_p(id_ptr), C.c_void_p(loc_ptr), C.c_void_p(ld_ptr)))

    def lazy_update(self):
        check(self._L.flk_lazy_update(self._h))

    def commit(self):
        """publish the WRITE buffer (single GPU: lazy_update; dist handles add the first halo exchange)"""
        self.lazy_update()

    def num_owned(self) -> int:
        return self.num_objects()

    def snapshot_owned_raw(self, id_ptr, loc_ptr, ld_ptr) -> int:
        return self.snapshot_raw(id_ptr, loc_ptr, ld_ptr)

    # ---- hot path
    def step(self, step: int, seed: int = 0xB12D, injected: np.ndarray | None = None):
        if injected is not None:
            injected = np.ascontiguousarray(injected, dtype=np.float32)
            check(self._L.flk_step(self._h, step, seed, _vp(injected), injected.size // 2))
        else:
            check(self._L.flk_step(self._h, step, seed, None, 0))

    def run(self, first_step: int, n_steps: int, seed: int = 0xB12D):
        check(self._L.flk_run(self._h, first_step, n_steps, seed))

    # ---- read side
    def _query(self, fn, loc: Real2D, dist: float) -> np.ndarray:
        cap = 256
        while True:
            ids = np.zeros(cap, dtype=np.uint32)
            n = C.c_uint32()
            rc = fn(self._h, loc.x, loc.y, dist, _vp(ids), cap, C.byref(n))
            if rc == _lib.E_CAPACITY and n.value > cap:
                cap = n.value
                continue
            check(rc)
            return ids[: n.value].copy()

    def get_neighbors_within_relax_distance(self, loc: Real2D, dist: float) -> np.ndarray:
        """ids of the birds the reference would return (flockers/src/model/bird.rs:29-31), in query order."""
        return self._query(self._L.flk_get_neighbors_within_relax_distance, loc, dist)

    def get_neighbors_within_distance(self, loc: Real2D, dist: float) -> np.ndarray:
        return self._query(self._L.flk_get_neighbors_within_distance, loc, dist)

    BIRD_DTYPE = np.dtype([("id", "<u4"), ("x", "<f4"), ("y", "<f4"), ("ldx", "<f4"), ("ldy", "<f4")])   # flk_bird

    def get_neighbors_objects(self, loc: Real2D, dist: float, relax: bool = True) -> np.ndarray:
        """what the reference's queries return: the objects themselves (`Vec<Bird>`, bird.rs:29-31), as a structured array
        of flk_bird {id, x, y, ldx, ldy} in query order"""
        fn = self._L.flk_get_neighbors_within_relax_distance_objects if relax else self._L.flk_get_neighbors_within_distance_objects
        cap = 256
        while True:
            out = np.zeros(cap, dtype=self.BIRD_DTYPE)
            n = C.c_uint32()
            rc = fn(self._h, loc.x, loc.y, dist, _vp(out), cap, C.byref(n))
            if rc == _lib.E_CAPACITY and n.value > cap:
                cap = n.value
                continue
            check(rc)
            return out[: n.value].copy()

    def get_neighbors_birds(self, loc: Real2D, dist: float, relax: bool = True) -> list:
        """the same as a list of Bird (the shape `for elem in vec` of bird.rs:52 consumes)"""
        return [Bird(int(o["id"]), Real2D(float(o["x"]), float(o["y"])), Real2D(float(o["ldx"]), float(o["ldy"])))
                for o in self.get_neighbors_objects(loc, dist, relax)]

    def query_batch_objects(self, loc: np.ndarray, dist: float, relax: bool = True):
        loc = np.ascontiguousarray(loc, dtype=np.float32)
        nq = loc.size // 2
        offsets = np.zeros(nq + 1, dtype=np.uint64)
        cap = max(64 * nq, 1024)
        while True:
            out = np.zeros(cap, dtype=self.BIRD_DTYPE)
            rc = self._L.flk_query_batch_objects(self._h, nq, _vp(loc), dist, int(relax), _vp(offsets), _vp(out), cap)
            if rc == _lib.E_CAPACITY and int(offsets[nq]) > cap:
                cap = int(offsets[nq])
                continue
            check(rc)
            return offsets, out[: int(offsets[nq])]

    def query_batch(self, loc: np.ndarray, dist: float, relax: bool = True):
        loc = np.ascontiguousarray(loc, dtype=np.float32)
        nq = loc.size // 2
        offsets = np.zeros(nq + 1, dtype=np.uint64)
        cap = max(64 * nq, 1024)
        while True:
            ids = np.zeros(cap, dtype=np.uint32)
            rc = self._L.flk_query_batch(self._h, nq, _vp(loc), dist, int(relax), _vp(offsets), _vp(ids), cap)
            if rc == _lib.E_CAPACITY and int(offsets[nq]) > cap:
                cap = int(offsets[nq])
                continue
            check(rc)
            return offsets, ids[: int(offsets[nq])]

    def get(self, bird_id: int) -> Bird | None:
        """Field2D::get(&object): the stored copy of the object with this id, or None"""
        out = (C.c_float * 4)()
        found = C.c_int32()
        check(self._L.flk_get_object(self._h, bird_id, out, C.byref(found)))
        if not found.value:
            return None
        return Bird(bird_id, Real2D(out[0], out[1]), Real2D(out[2], out[3]))

    def get_location(self, bird_id: int) -> Real2D | None:
        """Field2D::get_location(object) (flockers/src/visualization/bird_vis.rs:23)"""
        b = self.get(bird_id)
        return None if b is None else b.loc

    def num_objects(self) -> int:
        n = C.c_uint64()
        check(self._L.flk_num_objects(self._h, C.byref(n)))
        return n.value

    def snapshot(self):
        """(ids, loc[n,2], last_d[n,2]) of the READ buffer in (cell, id) order."""
        n = self.capacity
        ids = np.zeros(n, dtype=np.uint32)
        loc = np.zeros((n, 2), dtype=np.float32)
        ld = np.zeros((n, 2), dtype=np.float32)
        cnt = C.c_uint64()
        check(self._L.flk_snapshot(self._h, _vp(ids), _vp(loc), _vp(ld), C.byref(cnt)))
        k = cnt.value
        return ids[:k], loc[:k], ld[:k]

    def snapshot_with_payload(self):
        """(ids, loc, last_d, payload[n, payload_words]) of the READ buffer in (cell, id) order"""
        n = self.capacity
        ids = np.zeros(n, dtype=np.uint32)
        loc = np.zeros((n, 2), dtype=np.float32)
        ld = np.zeros((n, 2), dtype=np.float32)
        pay = np.zeros((n, max(self.payload_words, 1)), dtype=np.uint32)
        cnt = C.c_uint64()
        check(self._L.flk_snapshot_ex(self._h, _vp(ids), _vp(loc), _vp(ld), _vp(pay), C.byref(cnt)))
        k = cnt.value
        return ids[:k], loc[:k], ld[:k], pay[:k, : self.payload_words]

    def snapshot_raw(self, id_ptr, loc_ptr, ld_ptr) -> int:
        cnt = C.c_uint64()
        check(self._L.flk_snapshot(self._h, C.c_void_p(id_ptr), C.c_void_p(loc_ptr), C.c_void_p(ld_ptr), C.byref(cnt)))
        return cnt.value

    # ---- dense (id-indexed) host interface: 16 bytes per object each way
    def set_objects_dense(self, rec4: np.ndarray, id0: int = 0):
        """set_object_location for ids id0..id0+n-1; rec4[k] = (loc.x, loc.y, last_d.x, last_d.y) of object id0+k"""
        rec4 = np.ascontiguousarray(rec4, dtype=np.float32)
        assert rec4.size % 4 == 0
        check(self._L.flk_set_objects_dense(self._h, rec4.size // 4, id0, _vp(rec4)))

    def snapshot_dense(self, n: int, id0: int = 0) -> np.ndarray:
        """rec4[n,4] indexed by id - id0 (NaN rows for ids the field does not hold)"""
        out = np.zeros((n, 4), dtype=np.float32)
        check(self._L.flk_snapshot_dense(self._h, id0, n, _vp(out)))
        return out

    def set_objects_dense_raw(self, n: int, ptr: int, id0: int = 0):
        check(self._L.flk_set_objects_dense(self._h, n, id0, C.c_void_p(ptr)))

    def snapshot_dense_async_raw(self, n: int, ptr: int, id0: int = 0):
        """enqueue only; call synchronize() before reading the (pinned) buffer"""
        check(self._L.flk_snapshot_dense_async(self._h, id0, n, C.c_void_p(ptr)))

    # ---- frames for a front-end: download beside the simulation
    def frame_begin_raw(self, id_ptr: int, rec4_ptr: int) -> int:
        """enqueue a frame (pinned host buffers); ticks enqueued afterwards run while it downloads; -> object count"""
        n = C.c_uint64()
        check(self._L.flk_frame_begin(self._h, C.c_void_p(id_ptr) if id_ptr else None,
                                      C.c_void_p(rec4_ptr) if rec4_ptr else None, C.byref(n)))
        return n.value

    def frame_wait(self):
        check(self._L.flk_frame_wait(self._h))

    def snapshot_by_id(self):
        ids, loc, ld = self.snapshot()
        o = np.argsort(ids, kind="stable")
        return ids[o], loc[o], ld[o]

    def binning(self):
        cs = np.zeros(self.dw * self.dh + 1, dtype=np.uint32)
        ids = np.zeros(max(self.num_objects(), 1), dtype=np.uint32)
        check(self._L.flk_get_binning(self._h, _vp(cs), _vp(ids)))
        return cs, ids[: self.num_objects()]

    # ---- timing
    def synchronize(self):
        check(self._L.flk_synchronize(self._h))

    def elapsed_ms(self) -> float:
        ms = C.c_float()
        check(self._L.flk_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        n = C.c_uint64()
        check(self._L.flk_launch_count(self._h, C.byref(n)))
        return n.value

    def profile_enable(self, on):
        """1/True: events around each k_step launch; 2: also around the lazy_update phases; 0/False: off"""
        check(self._L.flk_profile_enable(self._h, int(on)))

    def profile_rebin(self):
        """(scan_ms, scatter_ms, rank_ms) per lazy_update, averaged over the profiled ones"""
        ms = (C.c_double * 3)()
        n = C.c_uint64()
        check(self._L.flk_profile_rebin(self._h, ms, C.byref(n)))
        k = max(n.value, 1)
        return ms[0] / k, ms[1] / k, ms[2] / k

    def profile_step_kernel(self):
        ms, n = C.c_double(), C.c_uint64()
        check(self._L.flk_profile_step_kernel(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


class Agent:  # krabmaga::engine::agent::Agent
    def step(self, state):
        raise NotImplementedError

    def is_stopped(self, state) -> bool:
        return False


class FlockStepper(Agent):
    """One scheduled agent whose step() is `for every Bird: Bird::step` on the GPU."""

    def __init__(self, seed: int = 0xB12D):
        self.seed = seed
        self.injected = None  # optional per-tick draws, indexed by bird id (parity tests)

    def step(self, state):
        state.field1.step(state.step, self.seed, self.injected)


class Schedule:
    """krabmaga::engine::schedule::Schedule restricted to repeating agents (Appendix B)."""

    def __init__(self):
        self.step_count = 0
        self.time = 0.0
        self.events: list[Agent] = []

    def schedule_repeating(self, agent: Agent, time: float = 0.0, ordering: int = 0):
        self.events.append(agent)

    def step(self, state):
        if self.step_count == 0:
            state.update(0)
        state.before_step(self)
        for agent in list(self.events):
            agent.step(state)
        state.after_step(self)
        self.step_count += 1
        self.time += 1.0
        state.step = self.step_count
        state.update(self.step_count)


class Flocker:
    """flockers/src/model/state.rs:12-73 with field1 on the GPU."""

    def __init__(self, dim, initial_flockers: int, device: int = 0, init_seed: int = 0xF10C, step_seed: int = 0xB12D,
                 capacity: int | None = None):
        self.step = 0
        self.dim = dim
        self.initial_flockers = initial_flockers
        self.init_seed, self.step_seed = init_seed, step_seed
        self.field1 = Field2D(dim[0], dim[1], DISCRETIZATION, TOROIDAL, capacity or max(initial_flockers, 1), device)
        self.stepper = FlockStepper(step_seed)

    def reset(self):  # state.rs:32-35
        self.step = 0
        self.field1.clear()

    def init(self, schedule: Schedule, birds=None):
        """state.rs:37-52.  `birds` = (ids, loc, last_d) arrays; default: uniform placement drawn on the host
        with numpy's Philox (the reference uses the unseeded thread RNG, so any placement is 'reference')."""
        if birds is None:
            rng = np.random.Generator(np.random.Philox(self.init_seed))
            r = rng.random((self.initial_flockers, 2), dtype=np.float32)
            loc = r * np.array(self.dim, dtype=np.float32)
            ids = np.arange(self.initial_flockers, dtype=np.uint32)
            ld = np.zeros_like(loc)
            birds = (ids, loc, ld)
        self.field1.set_objects(*birds)
        schedule.schedule_repeating(self.stepper, 0.0, 0)

    def update(self, step: int):  # state.rs:54-56
        self.field1.lazy_update()

    def before_step(self, schedule):
        pass

    def after_step(self, schedule):
        pass


def simulate(state: Flocker, n_steps: int, reps: int = 1, verbose: bool = False):
    """simulate_old!(state, step, reps, Info::Normal) (flockers/src/main.rs:47): returns [(seconds, steps/s)]."""
    out = []
    for _ in range(reps):
        schedule = Schedule()
        state.reset()
        state.init(schedule)
        state.field1.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_steps):
            schedule.step(state)
        state.field1.synchronize()
        dt = time.perf_counter() - t0
        out.append((dt, n_steps / dt if dt > 0 else float("inf")))
        if verbose:
            print(f"Time elapsed: {dt:.6f}s  steps/s: {out[-1][1]:.2f}")
    return out
