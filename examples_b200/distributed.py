"""Multi-GPU host layer: x-strip decomposition of the torus, one strip per GPU (flockers_mpi).

  DistField2D   one rank per process (torchrun); NCCL send/recv inside libflk.so; torch.distributed is used
                only to hand the NCCL unique id from rank 0 to the other ranks.
  MgpuGroup     all strips driven by one host thread (flk_mgpu_*), peer copies instead of NCCL.

Reference protocol: flockers_mpi/src/model/state.rs:51-175 and bird.rs:185-195 (owner by location, ghosts before
the step, migration after it).  `strip_of` / `owner_of_column` restate Kdtree::get_block_by_location
(state.rs:75, bird.rs:190) for a 1-D periodic strip partition.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from .field2d import DISCRETIZATION, Field2D, _vp


def strip_of(mx: int, world: int, rank: int) -> tuple[int, int]:
    """global cell columns [col0, col1) owned by `rank`; the slack column mx belongs to the owner of column 0."""
    return rank * mx // world, (rank + 1) * mx // world


def owner_of_column(mx: int, world: int, col: int) -> int:
    if col >= mx or col < 0:
        return 0
    for r in range(world):
        c0, c1 = strip_of(mx, world, r)
        if c0 <= col < c1:
            return r
    raise ValueError(col)


def balanced_columns(x, width, world: int, discretization=DISCRETIZATION) -> np.ndarray:
    """Kdtree-style partition along x (flockers_mpi/src/model/state.rs:33): world+1 column boundaries such that every
    strip holds about the same number of the sampled birds (flk_balance_columns)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros(world + 1, dtype=np.int32)
    check(_lib.load().flk_balance_columns(_vp(x), x.size, 1, width, discretization, world, _vp(out)))
    return out


def column_of(x, discretization=DISCRETIZATION) -> np.ndarray:
    """floor(x/d) in f32, like Field2D's discretisation."""
    q = np.floor(np.asarray(x, dtype=np.float32) / np.float32(discretization))
    return np.nan_to_num(q, nan=0.0).astype(np.int64)


def grid_columns(width, discretization=DISCRETIZATION) -> int:
    return int(np.ceil(np.float32(width) / np.float32(discretization)))


class _StripField(Field2D):
    """shared accessors of a strip handle"""

    def strip(self) -> tuple[int, int]:
        a, b = C.c_int32(), C.c_int32()
        check(self._L.flk_dist_strip(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def tile(self) -> tuple[int, int, int, int]:
        """global cell range (col0, col1, row0, row1) this rank owns (strips: every row)"""
        a, b, c, d = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        check(self._L.flk_dist_tile(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return a.value, b.value, c.value, d.value

    def owner(self, x: float, y: float) -> int:
        r = C.c_int32()
        check(self._L.flk_dist_owner(self._h, x, y, C.byref(r)))
        return r.value

    def num_owned(self) -> int:
        n = C.c_uint64()
        check(self._L.flk_dist_num_owned(self._h, C.byref(n)))
        return n.value

    def num_objects(self) -> int:
        # READ buffer incl. ghosts; refreshed from the device
        ids = None
        n = C.c_uint64()
        check(self._L.flk_snapshot(self._h, ids, None, None, C.byref(n)))
        return n.value

    def snapshot_owned_raw(self, id_ptr, loc_ptr, ld_ptr) -> int:
        n = C.c_uint64()
        check(self._L.flk_dist_snapshot_owned(self._h, C.c_void_p(id_ptr), C.c_void_p(loc_ptr), C.c_void_p(ld_ptr), C.byref(n)))
        return n.value

    def snapshot_owned_async_raw(self, id_ptr, loc_ptr, ld_ptr) -> int:
        """count now, records after synchronize() (pinned buffers)"""
        n = C.c_uint64()
        check(self._L.flk_dist_snapshot_owned_async(self._h, C.c_void_p(id_ptr), C.c_void_p(loc_ptr), C.c_void_p(ld_ptr),
                                                    C.byref(n)))
        return n.value

    def snapshot_owned(self):
        cap = self.capacity
        ids = np.zeros(cap, dtype=np.uint32)
        loc = np.zeros((cap, 2), dtype=np.float32)
        ld = np.zeros((cap, 2), dtype=np.float32)
        n = self.snapshot_owned_raw(ids.ctypes.data, loc.ctypes.data, ld.ctypes.data)
        return ids[:n], loc[:n], ld[:n]

    def binning(self):
        """LOCAL binning: (columns owned + 3) x dh cells, dh = local rows (tiles: rows owned + 3)"""
        ncell = (self.strip()[1] - self.strip()[0] + 3) * self.dh
        cs = np.zeros(ncell + 1, dtype=np.uint32)
        n = self.num_objects()
        ids = np.zeros(max(n, 1), dtype=np.uint32)
        check(self._L.flk_get_binning(self._h, _vp(cs), _vp(ids)))
        return cs, ids[:n]


class DistField2D(_StripField):
    """One strip of the torus on this process's GPU; collective calls must be made by every rank."""

    def __init__(self, width, height, discretization=DISCRETIZATION, capacity=1 << 20, device=0, rank=0, world=1,
                 col_bounds=None, tile_rows=1, row_bounds=None):
        import torch
        import torch.distributed as dist
        self._L = _lib.load()
        self._h = C.c_void_p()
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            check(self._L.flk_dist_unique_id(buf))
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            uid = uid.cuda(device)
        dist.broadcast(uid, src=0)
        raw = (C.c_uint8 * 128)(*uid.cpu().tolist())
        cb = None if col_bounds is None else np.ascontiguousarray(col_bounds, dtype=np.int32)
        rb = None if row_bounds is None else np.ascontiguousarray(row_bounds, dtype=np.int32)
        check(self._L.flk_dist_create_tiles(width, height, discretization, capacity, device, rank, world, raw, tile_rows,
                                            _vp(cb), _vp(rb), C.byref(self._h)))
        self.width, self.height, self.discretization, self.toroidal = width, height, discretization, True
        self.capacity, self.rank, self.world = capacity, rank, world
        dw, dh, mx, my = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        check(self._L.flk_field_dims(self._h, C.byref(dw), C.byref(dh), C.byref(mx), C.byref(my)))
        self.dw, self.dh, self.mx, self.my = dw.value, dh.value, mx.value, my.value

    def commit(self):
        check(self._L.flk_dist_commit(self._h))

    def lazy_update(self):
        raise _lib.FlkError(_lib.E_STATE, "a dist handle folds lazy_update into commit()/step()")

    def step(self, step: int, seed: int = 0xB12D, injected: np.ndarray | None = None):
        if injected is not None:
            injected = np.ascontiguousarray(injected, dtype=np.float32)
            check(self._L.flk_dist_step(self._h, step, seed, _vp(injected), injected.size // 2))
        else:
            check(self._L.flk_dist_step(self._h, step, seed, None, 0))


class MgpuGroup:
    """n strips in one process.  `devices[r]` is the CUDA ordinal of strip r (ordinals may repeat: that is how
    the strip logic is exercised on a single GPU)."""

    def __init__(self, width, height, discretization=DISCRETIZATION, capacity_per_rank=1 << 20, devices=(0, 0),
                 col_bounds=None, tile_rows=1, row_bounds=None, payload_words=0):
        self._L = _lib.load()
        n = len(devices)
        self.n = n
        self.payload_words = payload_words    # Field2D<O>: extra 32-bit words every object carries through the exchange
        self._arr = (C.c_void_p * n)()
        devs = (C.c_int * n)(*devices)
        cb = None if col_bounds is None else np.ascontiguousarray(col_bounds, dtype=np.int32)
        rb = None if row_bounds is None else np.ascontiguousarray(row_bounds, dtype=np.int32)
        check(self._L.flk_mgpu_create_ex(width, height, discretization, capacity_per_rank, n, devs, tile_rows, _vp(cb),
                                         _vp(rb), 4 * payload_words, self._arr))
        self.fields = []
        for r in range(n):
            f = _StripField._adopt(C.c_void_p(self._arr[r]), width, height, discretization, capacity_per_rank)
            f.rank, f.world = r, n
            f.payload_words = payload_words
            self.fields.append(f)
        self.mx, self.my, self.dh = self.fields[0].mx, self.fields[0].my, self.fields[0].dh

    def close(self):
        for f in self.fields:
            f.close()
        self.fields = []

    def set_params(self, **kw):
        for f in self.fields:
            f.set_params(**kw)

    def set_math_mode(self, mode):
        for f in self.fields:
            f.set_math_mode(mode)

    def set_objects(self, ids, loc, last_d, payload=None):
        """init scatter (flockers_mpi state.rs:51-103): every strip is offered every bird and keeps the ones it owns"""
        for f in self.fields:
            f.set_objects(ids, loc, last_d, payload)

    def snapshot_by_id_with_payload(self):
        """(ids, loc, last_d, payload[n, payload_words]) of the owned objects of every strip / tile, sorted by id"""
        parts = []
        for f in self.fields:
            cap = f.capacity
            ids = np.zeros(cap, dtype=np.uint32)
            loc = np.zeros((cap, 2), dtype=np.float32)
            ld = np.zeros((cap, 2), dtype=np.float32)
            pay = np.zeros((cap, max(self.payload_words, 1)), dtype=np.uint32)
            n = C.c_uint64()
            check(self._L.flk_dist_snapshot_owned_ex(f._h, _vp(ids), _vp(loc), _vp(ld), _vp(pay), C.byref(n)))
            k = n.value
            parts.append((ids[:k], loc[:k], ld[:k], pay[:k, : self.payload_words]))
        ids, loc, ld, pay = (np.concatenate([p[i] for p in parts]) for i in range(4))
        o = np.argsort(ids, kind="stable")
        return ids[o], loc[o], ld[o], pay[o]

    def scatter_objects(self, ids, loc, last_d):
        """init scatter done on the host like flockers_mpi's rank 0 (state.rs:65-92): every bird goes to the strip
        that owns its cell column (needs only capacity_per_rank >= the strip's own population)"""
        ids = np.asarray(ids)
        loc = np.asarray(loc, dtype=np.float32).reshape(-1, 2)
        last_d = np.asarray(last_d, dtype=np.float32).reshape(-1, 2)
        col = column_of(loc[:, 0], self.fields[0].discretization)
        row = column_of(loc[:, 1], self.fields[0].discretization)
        for f in self.fields:
            c0, c1, r0, r1 = f.tile()
            mc = (col >= c0) & (col < c1)
            if c0 == 0:
                mc |= col >= self.mx          # the slack column belongs to the tiles holding column 0
            mr = (row >= r0) & (row < r1)
            if r0 == 0:
                mr |= row >= self.my
            m = mc & mr
            f.set_objects(ids[m], loc[m], last_d[m])

    def commit(self):
        check(self._L.flk_mgpu_commit(self._arr, self.n))

    def step(self, step: int, seed: int = 0xB12D, injected: np.ndarray | None = None):
        if injected is not None:
            injected = np.ascontiguousarray(injected, dtype=np.float32)
            check(self._L.flk_mgpu_step(self._arr, self.n, step, seed, _vp(injected), injected.size // 2))
        else:
            check(self._L.flk_mgpu_step(self._arr, self.n, step, seed, None, 0))

    def run(self, first_step: int, n_steps: int, seed: int = 0xB12D):
        check(self._L.flk_mgpu_run(self._arr, self.n, first_step, n_steps, seed))

    def synchronize(self):
        for f in self.fields:
            f.synchronize()

    def num_owned(self):
        return [f.num_owned() for f in self.fields]

    def snapshot_by_id(self):
        parts = [f.snapshot_owned() for f in self.fields]
        ids = np.concatenate([p[0] for p in parts])
        loc = np.concatenate([p[1] for p in parts])
        ld = np.concatenate([p[2] for p in parts])
        o = np.argsort(ids, kind="stable")
        return ids[o], loc[o], ld[o]


class DistFlocker:
    """flockers_mpi/src/model/state.rs:21-190 with field1 on this rank's GPU strip.

    `before_step` (ghosts), every owned `Bird::step`, `after_step` (migration) and `update` (lazy_update) are ONE
    collective call, `field1.step`; the State hooks are kept so that a Schedule written against the reference
    still drives it."""

    def __init__(self, dim, initial_flockers: int, device: int, rank: int, world: int, capacity: int | None = None,
                 init_seed: int = 0xF10C, step_seed: int = 0xB12D):
        self.step = 0
        self.dim = dim
        self.initial_flockers = initial_flockers
        self.rank, self.world = rank, world
        self.init_seed, self.step_seed = init_seed, step_seed
        cap = capacity or max(2 * initial_flockers // world + 1024, 1024)
        self.field1 = DistField2D(dim[0], dim[1], DISCRETIZATION, capacity=cap, device=device, rank=rank, world=world)

    def init(self, schedule=None):
        """state.rs:51-103: rank 0's placement; here every rank draws the same placement and keeps what it owns
        (flk_set_objects on a strip handle drops birds of other strips), which needs no MPI scatter."""
        rng = np.random.Generator(np.random.Philox(self.init_seed))
        r = rng.random((self.initial_flockers, 2), dtype=np.float32)
        loc = r * np.array(self.dim, dtype=np.float32)
        ids = np.arange(self.initial_flockers, dtype=np.uint32)
        col = column_of(loc[:, 0], DISCRETIZATION)
        c0, c1 = self.field1.strip()
        mine = (col >= c0) & (col < c1)
        if c0 == 0:
            mine |= col >= self.field1.mx
        self.field1.set_objects(ids[mine], loc[mine], np.zeros_like(loc[mine]))
        self.field1.commit()

    def before_step(self, schedule=None):   # state.rs:113-133, folded into field1.step
        pass

    def after_step(self, schedule=None):    # state.rs:139-175, folded into field1.step
        pass

    def update(self, step: int):            # state.rs:105-107, folded into field1.step
        self.step = step

    def tick(self):
        self.before_step()
        self.field1.step(self.step, self.step_seed)
        self.after_step()
        self.update(self.step + 1)
