"""Shared test utilities: workload generators and structured-array conversions."""
import numpy as np

from oracle import oracle as O

D = O.DISCRETIZATION


def uniform_birds(n, W, H, seed=0xF10C):
    return O.make_uniform(seed, W, H, n)


def clustered_birds(n, W, H, seed=7, per_cluster=256, sigma=10.0):
    """BASELINE config 5 shape: gaussian blobs (sigma ~1.5 cells), wrapped onto the torus."""
    rng = np.random.default_rng(seed)
    k = max(n // per_cluster, 1)
    centres = rng.random((k, 2)) * [W, H]
    which = rng.integers(0, k, n)
    p = centres[which] + rng.normal(0.0, sigma, (n, 2))
    p = np.mod(p, [W, H]).astype(np.float32)
    p[:, 0] = np.where(p[:, 0] >= np.float32(W), 0, p[:, 0])
    p[:, 1] = np.where(p[:, 1] >= np.float32(H), 0, p[:, 1])
    b = np.zeros(n, dtype=O.BIRD)
    b["id"] = np.arange(n, dtype=np.uint32)
    b["x"], b["y"] = p[:, 0], p[:, 1]
    h = rng.normal(0, 0.4, (n, 2)).astype(np.float32)
    b["ldx"], b["ldy"] = h[:, 0], h[:, 1]
    return b


def to_arrays(b):
    ids = np.ascontiguousarray(b["id"])
    loc = np.ascontiguousarray(np.stack([b["x"], b["y"]], axis=1))
    ld = np.ascontiguousarray(np.stack([b["ldx"], b["ldy"]], axis=1))
    return ids, loc, ld


def from_arrays(ids, loc, ld):
    b = np.zeros(len(ids), dtype=O.BIRD)
    b["id"] = ids
    b["x"], b["y"] = loc[:, 0], loc[:, 1]
    b["ldx"], b["ldy"] = ld[:, 0], ld[:, 1]
    return b


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def state_digest(b):
    """order-independent 64-bit digest of a population (sorted by id first)."""
    import hashlib
    s = np.sort(b, order="id")
    return hashlib.sha256(np.ascontiguousarray(s).view(np.uint8).tobytes()).hexdigest()[:16]
