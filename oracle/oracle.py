"""ctypes loader for the CPU oracle (oracle/flockers_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package (examples_b200/).
PARITY UNPINNED: see the header of flockers_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

BIRD = np.dtype([("id", "<u4"), ("x", "<f4"), ("y", "<f4"), ("ldx", "<f4"), ("ldy", "<f4")])
assert BIRD.itemsize == 20


class Cfg(C.Structure):
    _fields_ = [
        ("cohesion", C.c_float), ("avoidance", C.c_float), ("randomness", C.c_float),
        ("consistency", C.c_float), ("momentum", C.c_float), ("jump", C.c_float),
        ("radius", C.c_float), ("relax", C.c_int), ("window_plus1", C.c_int), ("canonical_bags", C.c_int),
        ("wrap_with_slack", C.c_int), ("slack_to_zero", C.c_int),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "flockers_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        vp, f32, i32, i64, u64, u32 = C.c_void_p, C.c_float, C.c_int, C.c_int64, C.c_uint64, C.c_uint32
        sig = {
            "orc_cfg_default": (None, [C.POINTER(Cfg)]),
            "orc_toroidal_transform": (f32, [f32, f32]),
            "orc_toroidal_distance": (f32, [f32, f32, f32]),
            "orc_philox4x32_10": (None, [vp, vp, vp]),
            "orc_draw": (None, [u64, u32, u64, C.POINTER(f32), C.POINTER(f32)]),
            "orc_field_new": (vp, [f32, f32, f32, i32, i32]),
            "orc_field_free": (None, [vp]),
            "orc_field_set_semantics": (None, [vp, i32, i32]),
            "orc_field_dims": (None, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
            "orc_discretize": (None, [vp, f32, f32, C.POINTER(i32), C.POINTER(i32)]),
            "orc_set_object_location": (None, [vp, vp, f32, f32]),
            "orc_insert_read": (None, [vp, vp]),
            "orc_lazy_update": (None, [vp]),
            "orc_get_neighbors": (i64, [vp, f32, f32, f32, i32, i32, vp, i64]),
            "orc_field_dump": (i64, [vp, vp, i64, vp]),
            "orc_bird_step": (i32, [vp, vp, C.POINTER(Cfg), f32, f32]),
            "orc_sim_new": (vp, [f32, f32, f32, i32, C.POINTER(Cfg)]),
            "orc_sim_free": (None, [vp]),
            "orc_sim_field": (vp, [vp]),
            "orc_sim_num_agents": (i64, [vp]),
            "orc_sim_steps_done": (u64, [vp]),
            "orc_sim_init": (None, [vp, vp, i64]),
            "orc_make_uniform": (None, [u64, f32, f32, i64, vp]),
            "orc_sim_step": (None, [vp, u64, vp, i32]),
            "orc_sim_agents": (None, [vp, vp]),
            "orc_sim_replace_agents": (None, [vp, vp, i64]),
            "orc_sim_step_strip": (i64, [vp, u64, u64, vp, i32, i32, vp, i64]),
            "orc_sim_add_agents": (None, [vp, vp, i64]),
            "orc_field_columns": (i64, [vp, i32, i32, vp, i64]),
            "orc_field_cells": (i64, [vp, i32, i32, i32, i32, vp, i64]),
            "orc_sim_step_tile": (i64, [vp, u64, u64, vp, i32, i32, i32, i32, vp, i64]),
            "orc_max_threads": (i32, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def default_cfg(**kw) -> Cfg:
    c = Cfg()
    lib().orc_cfg_default(C.byref(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


DISCRETIZATION = float(np.float32(10.0) / np.float32(1.5))  # flockers/src/main.rs:32


def toroidal_transform(v, dim):
    return lib().orc_toroidal_transform(v, dim)


def toroidal_distance(a, b, dim):
    return lib().orc_toroidal_distance(a, b, dim)


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_ptr(c), _ptr(k), _ptr(o))
    return o


def draw(seed, bird_id, step):
    a, b = C.c_float(), C.c_float()
    lib().orc_draw(seed, bird_id, step, C.byref(a), C.byref(b))
    return a.value, b.value


def make_uniform(seed, w, h, n) -> np.ndarray:
    out = np.zeros(n, dtype=BIRD)
    lib().orc_make_uniform(seed, w, h, n, _ptr(out))
    return out


class Field:
    """Field2D<Bird> restatement."""

    def __init__(self, w, h, d=DISCRETIZATION, toroidal=True, canonical=True, _borrow=None):
        self._own = _borrow is None
        self.h = lib().orc_field_new(w, h, d, int(toroidal), int(canonical)) if _borrow is None else _borrow
        dw, dh, mx, my = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        lib().orc_field_dims(self.h, C.byref(dw), C.byref(dh), C.byref(mx), C.byref(my))
        self.dw, self.dh, self.mx, self.my = dw.value, dh.value, mx.value, my.value
        self.width, self.height, self.disc = w, h, d

    def __del__(self):
        if getattr(self, "_own", False) and self.h:
            lib().orc_field_free(self.h)
            self.h = None

    def set_semantics(self, wrap_with_slack=False, slack_to_zero=False):
        """the [EXT] alternatives E4 / E5 of SURVEY.md §8c (defaults = recalled upstream behaviour)"""
        lib().orc_field_set_semantics(self.h, int(wrap_with_slack), int(slack_to_zero))

    def discretize(self, x, y):
        cx, cy = C.c_int(), C.c_int()
        lib().orc_discretize(self.h, x, y, C.byref(cx), C.byref(cy))
        return cx.value, cy.value

    def set_object_location(self, bird: np.ndarray, x=None, y=None):
        b = np.array(bird, dtype=BIRD).reshape(())
        lib().orc_set_object_location(self.h, _ptr(b), float(b["x"]) if x is None else x,
                                      float(b["y"]) if y is None else y)

    def set_objects(self, birds: np.ndarray):
        birds = np.ascontiguousarray(birds, dtype=BIRD)
        for i in range(len(birds)):
            lib().orc_set_object_location(self.h, C.c_void_p(birds.ctypes.data + 20 * i),
                                          float(birds["x"][i]), float(birds["y"][i]))

    def insert_read(self, birds: np.ndarray):
        birds = np.ascontiguousarray(birds, dtype=BIRD)
        for i in range(len(birds)):
            lib().orc_insert_read(self.h, C.c_void_p(birds.ctypes.data + 20 * i))

    def lazy_update(self):
        lib().orc_lazy_update(self.h)

    def neighbors(self, x, y, dist=10.0, relax=True, plus1=False, cap=1 << 16) -> np.ndarray:
        out = np.zeros(cap, dtype=BIRD)
        n = lib().orc_get_neighbors(self.h, x, y, dist, int(relax), int(plus1), _ptr(out), cap)
        assert n <= cap
        return out[:n].copy()

    def dump(self):
        """READ buffer in (cell, bag) order + cell_start[dw*dh+1]."""
        n = lib().orc_field_dump(self.h, None, 0, None)
        out = np.zeros(max(n, 1), dtype=BIRD)
        cs = np.zeros(self.dw * self.dh + 1, dtype=np.uint32)
        lib().orc_field_dump(self.h, _ptr(out), n, _ptr(cs))
        return out[:n], cs

    def columns(self, c0, c1) -> np.ndarray:
        n = lib().orc_field_columns(self.h, c0, c1, None, 0)
        out = np.zeros(max(n, 1), dtype=BIRD)
        lib().orc_field_columns(self.h, c0, c1, _ptr(out), n)
        return out[:n]

    def cells(self, c0, c1, r0, r1) -> np.ndarray:
        """READ birds of the cell rectangle [c0,c1) x [r0,r1)"""
        n = lib().orc_field_cells(self.h, c0, c1, r0, r1, None, 0)
        out = np.zeros(max(n, 1), dtype=BIRD)
        lib().orc_field_cells(self.h, c0, c1, r0, r1, _ptr(out), n)
        return out[:n]

    def bird_step(self, bird: np.ndarray, cfg: Cfg, r1, r2):
        b = np.array(bird, dtype=BIRD).reshape(())
        drew = lib().orc_bird_step(self.h, _ptr(b), C.byref(cfg), r1, r2)
        return b, bool(drew)


class Sim:
    """Flocker state + Schedule restatement (flockers/src/model/state.rs)."""

    def __init__(self, w, h, d=DISCRETIZATION, toroidal=True, cfg: Cfg | None = None):
        self.cfg = cfg or default_cfg()
        self.h = lib().orc_sim_new(w, h, d, int(toroidal), C.byref(self.cfg))
        self.field = Field(w, h, d, toroidal, _borrow=lib().orc_sim_field(self.h))
        self.w, self.hh = w, h

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_sim_free(self.h)
            self.h = None

    def init(self, birds: np.ndarray):
        birds = np.ascontiguousarray(birds, dtype=BIRD)
        assert np.all(np.diff(birds["id"].astype(np.int64)) > 0), "ascending ids"
        lib().orc_sim_init(self.h, _ptr(birds), len(birds))

    def step(self, seed=0xB12D, injected: np.ndarray | None = None, threads=1):
        inj = None
        if injected is not None:
            injected = np.ascontiguousarray(injected, dtype=np.float32)
            inj = _ptr(injected)
        lib().orc_sim_step(self.h, seed, inj, threads)

    @property
    def steps_done(self):
        return lib().orc_sim_steps_done(self.h)

    def agents(self) -> np.ndarray:
        n = lib().orc_sim_num_agents(self.h)
        out = np.zeros(max(n, 1), dtype=BIRD)
        lib().orc_sim_agents(self.h, _ptr(out))
        return out[:n]

    # rank-decomposed protocol pieces (flockers_mpi)
    def step_strip(self, step, seed, injected, col0, col1):
        n = lib().orc_sim_num_agents(self.h)
        out = np.zeros(max(n, 1), dtype=BIRD)
        inj = None
        if injected is not None:
            injected = np.ascontiguousarray(injected, dtype=np.float32)
            inj = _ptr(injected)
        nl = lib().orc_sim_step_strip(self.h, step, seed, inj, col0, col1, _ptr(out), n)
        return out[:nl].copy()

    def step_tile(self, step, seed, injected, col0, col1, row0, row1):
        n = lib().orc_sim_num_agents(self.h)
        out = np.zeros(max(n, 1), dtype=BIRD)
        inj = None
        if injected is not None:
            injected = np.ascontiguousarray(injected, dtype=np.float32)
            inj = _ptr(injected)
        nl = lib().orc_sim_step_tile(self.h, step, seed, inj, col0, col1, row0, row1, _ptr(out), n)
        return out[:nl].copy()

    def add_agents(self, birds):
        birds = np.ascontiguousarray(birds, dtype=BIRD)
        lib().orc_sim_add_agents(self.h, _ptr(birds), len(birds))


def max_threads() -> int:
    return lib().orc_max_threads()
