"""CPU tests of the drop-in boundary: libflk.so loads, exports every symbol include/flk.h declares, refuses to
compute without a device, and the product package never touches the oracle."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "flk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(flk_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from examples_b200 import build
    build.build()
    from examples_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(lib):
    L = C.CDLL(lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/flk.h but not exported by libflk.so"


def test_binding_covers_every_declared_symbol(lib):
    assert sorted(lib.SIGNATURES) == declared_symbols()
    lib.load()


def test_no_torch_types_in_signatures():
    src = open(os.path.join(ROOT, "include", "flk.h")).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)   # declarations only
    assert "torch" not in code.lower() and "at::" not in code and "std::" not in code and "Tensor" not in code


def test_version_and_error_strings(lib):
    L = lib.load()
    assert b"sm_100a" in L.flk_version()
    assert L.flk_field_destroy(None) == 0
    assert L.flk_lazy_update(None) == lib.E_INVALID
    assert b"null handle" in L.flk_last_error()


def test_invalid_arguments_are_rejected_before_touching_the_device(lib):
    L = lib.load()
    h = C.c_void_p()
    assert L.flk_field_create(100.0, 100.0, 6.6666665, 1, 0, 0, C.byref(h)) == lib.E_INVALID   # capacity 0
    assert L.flk_field_create(100.0, 100.0, 6.6666665, 1, 10, 0, None) == lib.E_INVALID


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    L = lib.load()
    h = C.c_void_p()
    rc = L.flk_field_create(100.0, 100.0, 6.6666665, 1, 1000, 0, C.byref(h))
    assert rc == lib.E_CUDA and not h.value
    assert b"no CPU fallback" in L.flk_last_error()
    from examples_b200 import Field2D, FlkError
    with pytest.raises(FlkError):
        Field2D(100.0, 100.0, capacity=10)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "examples_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dp, fn)).read()
                assert "oracle" not in text.lower(), f"{fn} mentions the oracle"


def test_balance_columns_is_host_side_and_balanced(lib):
    """flk_balance_columns (the Kdtree-style partition along x) needs no device"""
    import ctypes as C
    import numpy as np
    from examples_b200.distributed import balanced_columns, column_of, grid_columns
    W = 1000.0
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.normal(200.0, 30.0, 30000), rng.uniform(0, W, 10000)]).astype(np.float32)
    x = np.mod(x, np.float32(W)).astype(np.float32)
    mx = grid_columns(W)
    for world in (1, 2, 3, 8):
        b = balanced_columns(x, W, world)
        assert b[0] == 0 and b[-1] == mx and np.all(np.diff(b) >= 2)
        col = column_of(x)
        counts = np.array([np.sum((col >= b[r]) & (col < b[r + 1])) for r in range(world)])
        assert counts.sum() == x.size
        if world > 1:
            assert counts.max() < 2.2 * x.size / world, (world, counts)     # equal widths would put ~80 % in one strip
    # a uniform sample gives (nearly) equal widths; too many strips for the grid is an error
    u = rng.uniform(0, W, 50000).astype(np.float32)
    assert np.abs(np.diff(balanced_columns(u, W, 5)) - mx / 5).max() <= 3
    out = (C.c_int32 * 200)()
    assert lib.load().flk_balance_columns(u.ctypes.data_as(C.c_void_p), u.size, 1, W, 6.6666665, 100, out) < 0


def test_rust_shim_declares_every_export():
    """rust/src/ffi.rs (source only: no cargo in this image) binds exactly the symbols include/flk.h declares"""
    ffi = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    have = sorted(set(re.findall(r"pub fn (flk_[a-z0-9_]+)", ffi)))
    assert have == declared_symbols()
    assert "pub struct flk_bird" in ffi and "#[repr(C)]" in ffi
