"""examples_b200 — B200-native Flockers hot path behind krABMaga's Field2D/State/Agent surface.

The compute lives in libflk.so (CUDA, sm_100a; C ABI in include/flk.h).  This package is the thin
host-side mirror of the reference interface used by tests and bench.py; it never falls back to a CPU path.
"""
from ._lib import FlkError, load  # noqa: F401
from .field2d import (  # noqa: F401
    AVOIDANCE, COHESION, CONSISTENCY, DISCRETIZATION, JUMP, MOMENTUM, RANDOMNESS, TOROIDAL,
    Bird, Field2D, Flocker, FlockStepper, Real2D, Schedule, simulate,
)
