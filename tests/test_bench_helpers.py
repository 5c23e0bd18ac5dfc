"""CPU tests of bench.py's workload construction (no GPU, no timing)."""
import json
import subprocess
import sys
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_field_sizes_give_the_requested_grid():
    for cells in (15, 480, 672, 1344, 1920, 2688, 5376):
        w = bench.field_size(cells)
        assert int(np.ceil(np.float32(w) / np.float32(bench.D))) == cells


def test_workload_geometry_matches_baseline_configs():
    gx, gy, n, _ = bench.workload_geometry("weak16m", 8, None)
    assert (gx, gy, n) == (5376, 5376, 128_000_000)          # BASELINE config 4: 128M on 8 GPUs
    gx, gy, n, _ = bench.workload_geometry("weak16m", 1, None)
    assert (gx, gy, n) == (672, 5376, 16_000_000)
    assert abs(n / (bench.field_size(gx) * bench.field_size(gy)) - 0.1) < 0.001      # density ~0.1 birds/unit^2
    assert bench.workload_geometry("strong16m", 4, None)[:3] == (1920, 1920, 16_000_000)
    assert bench.workload_geometry("1m", 1, None)[:3] == (480, 480, 1_000_000)


def test_generated_birds_stay_inside_their_strip():
    W = bench.field_size(64)
    for x0, x1 in ((0.0, W / 2), (W / 2, W)):
        loc = bench.make_birds(20000, W, W, False, x0, x1)
        assert loc[:, 0].min() >= x0 and loc[:, 0].max() < x1 and loc[:, 1].max() < W
        locc = bench.make_birds(20000, W, W, True, x0, x1)
        assert locc[:, 0].min() >= x0 and locc[:, 0].max() < x1


def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-birds", "20000"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "bird-updates/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_roofline_inputs_come_from_the_committed_capture():
    """roofline.traffic and roofline.issue_slots use the per-launch DRAM bytes / warp-instructions of the committed ncu
    capture, only for the launch size it was taken on"""
    assert bench.ncu_traffic_per_launch(16_000_000) > 16_000_000 * bench.ALGO_BYTES_PER_UPDATE
    assert bench.ncu_traffic_per_launch(1_000_000) is None
    wi = bench.ncu_traffic_per_launch(16_000_000, "warp_instructions_per_launch")
    assert 16_000_000 / 32 * 500 < wi < 16_000_000 / 32 * 5000          # ~1440 instructions per warp of 32 birds
    r = bench.issue_roofline(wi, 0.85, 148, 1965.0)
    assert abs(r["peak"] - 148 * 4 * 1.965) < 1e-6 and 0.5 < r["frac"] < 1.0
    assert bench.issue_roofline(None, 0.85, 148, 1965.0) is None and bench.issue_roofline(wi, 0.85, 148, None) is None
