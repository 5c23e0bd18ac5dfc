"""The C++ host mirror (examples_b200/host) runs the Flockers model end to end through the C ABI."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_flockers_demo_runs_shipped_config():
    from examples_b200 import build
    exe = build.build_host_demo()
    res = subprocess.run([exe, "1000", "100", "200"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert "birds: 1000" in res.stdout and "steps/s" in res.stdout


def test_python_mirror_simulate_matches_flk_run():
    """Schedule.step-driven simulate (update(0), agents, update(t+1)) equals the fused flk_run loop."""
    import numpy as np
    from examples_b200 import Field2D, Flocker, Schedule, simulate
    st = Flocker((100.0, 100.0), 1000)
    simulate(st, 30, 1)
    a = st.field1.snapshot_by_id()
    # same placement through the raw field
    rng = np.random.Generator(np.random.Philox(0xF10C))
    loc = rng.random((1000, 2), dtype=np.float32) * np.array([100.0, 100.0], dtype=np.float32)
    f = Field2D(100.0, 100.0, capacity=1000)
    f.set_objects(np.arange(1000, dtype=np.uint32), loc, np.zeros_like(loc))
    f.lazy_update()
    f.run(0, 30, 0xB12D)
    b = f.snapshot_by_id()
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
