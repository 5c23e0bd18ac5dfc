"""How fast does this box move 320 MB between pinned host memory and the GPU, and how long do the pieces of the
e2e step take?  (diagnostic; run under gpurun)"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from examples_b200 import Field2D
n = 16_000_000
h = torch.empty(n * 5, dtype=torch.float32).pin_memory()
d = torch.empty(n * 5, dtype=torch.float32, device='cuda')
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {n*20/dt/1e9:.1f} GB/s ({dt*1e3:.2f} ms for 320 MB)")
D = float(np.float32(10.0) / np.float32(1.5))
W = float(np.float32(np.float32(672) * np.float32(D))); H = float(np.float32(np.float32(5376) * np.float32(D)))
f = Field2D(W, H, D, True, capacity=n + 1000)
rng = np.random.default_rng(0)
ids = torch.arange(n, dtype=torch.int32).pin_memory()
loc = torch.from_numpy((rng.random((n, 2), dtype=np.float32) * np.array([W, H], dtype=np.float32)).astype(np.float32)).pin_memory()
ld = torch.zeros((n, 2), dtype=torch.float32).pin_memory()
def timed(label, fn, reps=3):
    fn(); f.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    f.synchronize(); print(f"{label}: {(time.perf_counter()-t0)/reps*1e3:.2f} ms")
def up(): f.set_objects_raw(n, ids.data_ptr(), loc.data_ptr(), ld.data_ptr()); f.lazy_update()
timed("set_objects + lazy_update", up)
timed("run 1 tick", lambda: f.run(0, 1))
timed("snapshot", lambda: f.snapshot_raw(ids.data_ptr(), loc.data_ptr(), ld.data_ptr()))
