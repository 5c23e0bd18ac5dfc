"""Debug driver for the slotted layout: slots vs sorted runs on the same field, first mismatch located, step-kernel counters."""
import ctypes as C
import sys
import numpy as np
sys.path.insert(0, ".")
from examples_b200 import Field2D
from tests.helpers import D, from_arrays, to_arrays, uniform_birds


def counters(f):
    out = (C.c_uint32 * 16)()
    f._L.flk_debug_counters(f._h, out)
    return list(out)[:9]


for n, W, ticks in [(1000, 100.0, 3), (20000, 450.0, 2), (1_000_000, 3200.0, 2)]:
    birds = uniform_birds(n, W, W)
    fs = Field2D(W, W, D, True, capacity=n)
    fs.set_layout(1)
    fr = Field2D(W, W, D, True, capacity=n)
    fr.set_layout(0)
    for f in (fs, fr):
        f.set_objects(*to_arrays(birds))
        f.lazy_update()
    print("layout", fs.layout(), fr.layout()[0], flush=True)
    a0, b0 = from_arrays(*fs.snapshot_by_id()), from_arrays(*fr.snapshot_by_id())
    print("  after publish equal:", np.array_equal(a0.view(np.uint8), b0.view(np.uint8)))
    cs_s, ids_s = fs.binning()
    cs_r, ids_r = fr.binning()
    print("  binning equal:", np.array_equal(cs_s, cs_r), np.array_equal(ids_s, ids_r))
    for t in range(ticks):
        counters(fs)
        fs.step(t); fs.lazy_update()
        fr.step(t); fr.lazy_update()
        fs.synchronize()
        c = counters(fs)
        a, b = from_arrays(*fs.snapshot_by_id()), from_arrays(*fr.snapshot_by_id())
        bad = np.nonzero((a["x"].view(np.uint32) != b["x"].view(np.uint32)) | (a["y"].view(np.uint32) != b["y"].view(np.uint32)) |
                         (a["ldx"].view(np.uint32) != b["ldx"].view(np.uint32)) | (a["id"] != b["id"]))[0] if len(a) == len(b) else None
        print(f"  n={n} tick {t}: ctas={c[0]} tile={c[1]} general={c[2]} flag={c[3]} notcovered={c[4]} rowcap={c[5]} genbirds={c[7]}"
              f" len {len(a)} {len(b)} mismatches {None if bad is None else len(bad)}", flush=True)
        if bad is not None and len(bad):
            prev = b0 if t == 0 else prevb
            for i in bad[:5]:
                px, py = prev["x"][i], prev["y"][i]
                print("    id", i, "prev cell", int(np.floor(px / np.float32(D))), int(np.floor(py / np.float32(D))),
                      "slots", a[i], "runs", b[i])
        prevb = b
    fs.close(); fr.close()


# timing: k_step per tick, both layouts
import bench
for name in ("1m", "weak16m"):
    gx, gy, n, _ = bench.workload_geometry(name, 1, None)
    W, H = bench.field_size(gx), bench.field_size(gy)
    loc = bench.make_birds(n, W, H, False)
    ids = np.arange(n, dtype=np.uint32)
    for lay in (1, 0):
        f = Field2D(W, H, D, True, capacity=int(n * 1.25) + 65536)
        f.set_layout(lay)
        f.set_objects(ids, loc, np.zeros_like(loc))
        f.lazy_update()
        f.run(0, 20)
        f.synchronize()
        f.profile_enable(2)
        f.run(20, 30)
        f.synchronize()
        ms, k = f.profile_step_kernel()
        reb = f.profile_rebin()
        tick = f.elapsed_ms() / 30
        f.profile_enable(0)
        f.run(50, 100)
        f.synchronize()
        print(f"{name} layout {f.layout()}: k_step {ms / k:.4f} ms, rebin {[round(x, 4) for x in reb]}, tick {tick:.4f} ms, "
              f"graph tick {f.elapsed_ms() / 100:.4f} ms", flush=True)
        f.close()
