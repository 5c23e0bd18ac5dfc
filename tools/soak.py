"""Soak: 16 M birds, hundreds of ticks, single field vs 4 in-process strips on one GPU, compared bit for bit (rare
events: x == W slack cells, seam crossings, migration bursts).  python tools/soak.py [ticks] [tile_rows]"""
import sys, time, hashlib
sys.path.insert(0, '.')
import numpy as np
import bench
from examples_b200 import Field2D
from examples_b200.distributed import MgpuGroup
D = bench.D
ticks = int(sys.argv[1]) if len(sys.argv) > 1 else 300
gx, gy, n, _ = bench.workload_geometry("weak16m", 1, None)
W, H = bench.field_size(gx), bench.field_size(gy)
loc = bench.make_birds(n, W, H, False)
ids = np.arange(n, dtype=np.uint32); ld = np.zeros_like(loc)
f = Field2D(W, H, D, True, capacity=n); f.set_objects(ids, loc, ld); f.lazy_update()
tile_rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1   # 2: 2 x 2 tiles instead of 4 strips
g = MgpuGroup(W, H, D, capacity_per_rank=int(n / 4 * 1.2) + 100000, devices=[0] * 4, tile_rows=tile_rows)
g.scatter_objects(ids, loc, ld); g.commit()
def digest(t):
    i, l, d = t
    return hashlib.sha256(i.tobytes() + l.tobytes() + d.tobytes()).hexdigest()[:16]
done = 0
while done < ticks:
    k = min(100, ticks - done)
    f.run(done, k); g.run(done, k); done += k
    a, b = f.snapshot_by_id(), g.snapshot_by_id()
    atW = int((a[1][:, 0] == np.float32(W)).sum()); atH = int((a[1][:, 1] == np.float32(H)).sum())
    print(f"tick {done}: single {digest(a)} strips {digest(b)} equal={digest(a)==digest(b)} owned={sum(g.num_owned())} x==W:{atW} y==H:{atH}", flush=True)
    assert digest(a) == digest(b)
print("SOAK_OK")
