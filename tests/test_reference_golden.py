"""Parity-pinning kit, checking side (INTEGRATION.md "Pinning parity to the real engine").

`tests/golden/reference_pin.txt` is what rust/examples/dump_golden.rs writes when the UNMODIFIED reference model runs on the
real krabmaga engine over `tests/golden/pin_input.txt`.  This image has no Rust toolchain, so the file does not exist here
and the reference-facing tests skip; the machinery itself is exercised against the CPU restatement's own dump, so that the
day the file arrives one command (`python -m pytest tests/test_reference_golden.py`) pins parity or names the [EXT]
assumption (E3 window width, E4 wrap modulus, E5 slack cell) that has to change."""
import itertools
import os

import numpy as np
import pytest

from tools import pin_format as P
from tools import pin_from_oracle as PO

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
INPUT = os.path.join(GOLDEN, "pin_input.txt")
REFERENCE = os.path.join(GOLDEN, "reference_pin.txt")


def compare(ref, queries, tresults, states):
    """-> list of human-readable differences between a pin output (dict from P.read_output) and computed results"""
    diffs = []
    for i, ids in enumerate(queries):
        want = ref["queries"].get(i)
        if want is None:
            diffs.append(f"query {i} missing")
        elif list(map(int, ids)) != want:
            kind = "order" if sorted(map(int, ids)) == sorted(want) else "set"
            diffs.append(f"query {i}: neighbour {kind} differs ({len(ids)} vs {len(want)} ids)")
    for i, v in enumerate(tresults):
        got = int(np.float32(v).view(np.uint32))
        if ref["tresults"].get(i) != got and not (np.isnan(np.float32(v)) and np.isnan(P.f32("%08x" % ref["tresults"].get(i, 0)))):
            diffs.append(f"toroidal probe {i}: {got:08x} vs {ref['tresults'].get(i, 0):08x}")
    for t, b in sorted(states.items()):
        want = ref["states"].get(t)
        if want is None:
            diffs.append(f"state {t} missing")
            continue
        bad = int((b.view(np.uint8).reshape(len(b), -1) != want.view(np.uint8).reshape(len(want), -1)).any(axis=1).sum()) \
            if len(b) == len(want) else -1
        if bad:
            err = float(max(np.abs(b["x"] - want["x"]).max(), np.abs(b["y"] - want["y"]).max())) if bad > 0 else float("nan")
            diffs.append(f"state after tick {t}: {bad} birds differ (max position error {err:.3g})")
    return diffs


def test_pin_input_is_the_committed_one(tmp_path):
    """the committed input is what its generator writes (so the Rust side and this side read the same birds)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_pin_input", os.path.join(GOLDEN, "make_pin_input.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    old = open(INPUT).read()
    m.HERE = str(tmp_path)
    m.main()
    assert open(tmp_path / "pin_input.txt").read() == old
    inp = P.read_input(INPUT)
    assert len(inp["birds"]) == 400 and inp["ticks"] == 12 and inp["birds"]["x"][0] == np.float32(100.0)


def test_kit_round_trip_and_switch_detection(tmp_path):
    """the comparison machinery: a dump of the restatement compares clean against itself, and a dump made with one of
    the [EXT] alternatives is told apart from the default by exactly the probes written for it"""
    inp = P.read_input(INPUT)
    q, t, s = PO.run(inp)
    path = tmp_path / "pin.txt"
    P.write_output(path, "oracle", q, t, s)
    ref = P.read_output(path)
    assert compare(ref, q, t, s) == []
    for sw in ({"window_plus1": 1}, {"wrap_with_slack": 1}, {"slack_to_zero": 1}):
        q2, t2, s2 = PO.run(inp, **sw)
        d = compare(ref, q2, t2, s2)
        assert d, f"the pin input cannot tell {sw} from the default"


@pytest.mark.skipif(not os.path.exists(REFERENCE), reason="no dump of the real engine (needs cargo: rust/examples/dump_golden.rs)")
def test_restatement_matches_the_real_engine():
    """PINS THE ORACLE: which setting of the [EXT] switches reproduces the real engine's dump?  The default must."""
    inp = P.read_input(INPUT)
    ref = P.read_output(REFERENCE)
    report = {}
    for e3, e4, e5 in itertools.product((0, 1), repeat=3):
        q, t, s = PO.run(inp, window_plus1=e3, wrap_with_slack=e4, slack_to_zero=e5)
        report[(e3, e4, e5)] = compare(ref, q, t, s)
    matching = [k for k, v in report.items() if not v]
    assert (0, 0, 0) in matching, ("the restatement's defaults do not reproduce the real engine; settings (E3, E4, E5) that do: "
                                   f"{matching}; first differences with the defaults: {report[(0, 0, 0)][:8]}")


@pytest.mark.gpu
def test_cuda_path_matches_the_pin_dump():
    """the CUDA path against the real engine's dump when there is one, else against the restatement's: queries in
    order, every state bit for bit"""
    from examples_b200 import Field2D, Real2D
    inp = P.read_input(INPUT)
    if os.path.exists(REFERENCE):
        ref = P.read_output(REFERENCE)
    else:
        q, t, s = PO.run(inp)
        ref = {"queries": {i: list(map(int, x)) for i, x in enumerate(q)}, "tresults": {}, "states": s}
    b = inp["birds"]
    W, H, d = float(inp["W"]), float(inp["H"]), float(inp["d"])
    f = Field2D(W, H, d, True, capacity=len(b))
    f.set_params(randomness=0.0)
    f.set_objects(np.ascontiguousarray(b["id"]), np.stack([b["x"], b["y"]], 1), np.stack([b["ldx"], b["ldy"]], 1))
    f.lazy_update()
    for i, (x, y, r, rel) in enumerate(inp["probes"]):
        k = int(np.floor(np.float32(r) / np.float32(d)))
        if 2 * k + 1 > f.mx:
            continue
        got = f.get_neighbors_within_relax_distance(Real2D(float(x), float(y)), float(r)) if rel else \
            f.get_neighbors_within_distance(Real2D(float(x), float(y)), float(r))
        assert list(map(int, got)) == ref["queries"][i], i
    for t in range(1, inp["ticks"] + 1):
        f.step(t - 1)
        f.lazy_update()
        ids, loc, ld = f.snapshot_by_id()
        want = ref["states"][t]
        assert np.array_equal(ids, want["id"])
        assert np.array_equal(loc[:, 0].view(np.uint32), want["x"].view(np.uint32)), t
        assert np.array_equal(loc[:, 1].view(np.uint32), want["y"].view(np.uint32)), t
        assert np.array_equal(ld[:, 0].view(np.uint32), want["ldx"].view(np.uint32)), t
        assert np.array_equal(ld[:, 1].view(np.uint32), want["ldy"].view(np.uint32)), t
