"""Soak on the benchmark geometry: 16 M birds, hundreds of ticks, the single field against four in-process strips that
share the GPU (their kernels overlap on different streams), bit for bit.  Short runs never showed it, this one did: the
single-launch scan once published a tile status after the last CTA had cleared the state words (missing release fence
before the done counter), which corrupted the next tick's cell_start once in a few hundred ticks — only with other
kernels running beside the scan.  Also exercises the rare events of long runs (x == W slack cells, migration bursts)."""
import hashlib

import numpy as np
import pytest

import bench

pytestmark = pytest.mark.gpu


def digest(t):
    i, l, d = t
    return hashlib.sha256(i.tobytes() + l.tobytes() + d.tobytes()).hexdigest()[:16]


@pytest.mark.parametrize("tile_rows", [1, 2])
def test_strips_sharing_one_gpu_match_the_single_field_over_hundreds_of_ticks(tile_rows):
    from examples_b200 import Field2D
    from examples_b200.distributed import MgpuGroup
    ticks = 400 if tile_rows == 1 else 200
    gx, gy, n, _ = bench.workload_geometry("weak16m", 1, None)
    W, H = bench.field_size(gx), bench.field_size(gy)
    loc = bench.make_birds(n, W, H, False)
    ids = np.arange(n, dtype=np.uint32)
    ld = np.zeros_like(loc)
    f = Field2D(W, H, bench.D, True, capacity=n)
    f.set_objects(ids, loc, ld)
    f.lazy_update()
    g = MgpuGroup(W, H, bench.D, capacity_per_rank=int(n / 4 * 1.2) + 100000, devices=[0] * 4, tile_rows=tile_rows)
    g.scatter_objects(ids, loc, ld)
    g.commit()
    done = 0
    while done < ticks:
        f.run(done, 100)
        g.run(done, 100)
        done += 100
        a, b = f.snapshot_by_id(), g.snapshot_by_id()
        assert sum(g.num_owned()) == n
        assert digest(a) == digest(b), f"strips diverged from the single field by tick {done}"
    g.close()
    f.close()
