"""Writes tests/golden/pin_input.txt, the input of the parity-pinning kit (INTEGRATION.md "Pinning parity to the real engine").

The same file is read by
  rust/examples/dump_golden.rs   the UNMODIFIED reference model on the real krabmaga engine (needs cargo; not in this image)
  tools/pin_from_oracle.py       the CPU restatement under oracle/ (writes the same output format)
and tests/test_reference_golden.py compares what they wrote, and the CUDA path, field by field.

Text format, every float as the hex of its IEEE-754 bits:
  field <W> <H> <d> <toroidal>
  birds <n>            then n lines   <id> <x> <y> <last_dx> <last_dy>
  ticks <T>
  probes <P>           then P lines   <x> <y> <dist> <relax 0|1>          (queries on the published initial state)
  tprobes <Q>          then Q lines   <kind 0 transform | 1 distance> <a> <b> <dim>
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def bits(v):
    return "%08x" % int(np.float32(v).view(np.uint32))


def main():
    rng = np.random.default_rng(0xF10C)
    W = H = np.float32(100.0)
    d = np.float32(10.0) / np.float32(1.5)
    n, T = 400, 12
    loc = (rng.random((n, 2), dtype=np.float32) * W).astype(np.float32)
    # edge cases of SURVEY.md §8c: x == W (E5), the seam, a cell boundary, the far corner, coincident birds
    loc[0] = [100.0, 50.0]
    loc[1] = [50.0, 100.0]
    loc[2] = [0.0, 0.0]
    loc[3] = [np.nextafter(np.float32(100.0), np.float32(0)), np.nextafter(np.float32(100.0), np.float32(0))]
    loc[4] = [d, d]
    loc[5] = [d * np.float32(14), d * np.float32(14)]
    loc[6] = loc[7] = [33.25, 71.5]
    ld = (rng.normal(0, 0.4, (n, 2))).astype(np.float32)
    probes = [(float(x), float(y), 10.0, r) for (x, y) in loc[:24] for r in (1, 0)]
    probes += [(50.0, 50.0, 10.0, 1), (50.0, 50.0, 15.0, 1), (50.0, 50.0, 5.0, 0), (0.5, 99.5, 10.0, 1), (99.9, 0.1, 10.0, 0),
               (100.0, 100.0, 10.0, 1), (50.0, 50.0, 0.0, 1), (50.0, 50.0, 20.5, 0)]
    tp = []
    for v in (-1e-6, -0.0, 0.0, 0.25, 99.99999, 100.0, 100.00001, 250.5, -250.5, 1e-30, -1e-30):
        tp.append((0, v, 0.0, 100.0))
    for a, b in ((1.0, 99.0), (99.0, 1.0), (50.0, 0.0), (0.0, 50.0), (25.0, 75.0), (75.0, 25.0), (10.0, 20.0), (100.0, 0.0),
                 (0.0, 100.0), (50.000004, 0.0), (49.999996, 0.0)):
        tp.append((1, a, b, 100.0))
    with open(os.path.join(HERE, "pin_input.txt"), "w") as f:
        f.write("# flk pin input v1 (tests/golden/make_pin_input.py)\n")
        f.write(f"field {bits(W)} {bits(H)} {bits(d)} 1\n")
        f.write(f"birds {n}\n")
        for i in range(n):
            f.write(f"{i} {bits(loc[i, 0])} {bits(loc[i, 1])} {bits(ld[i, 0])} {bits(ld[i, 1])}\n")
        f.write(f"ticks {T}\n")
        f.write(f"probes {len(probes)}\n")
        for x, y, r, rel in probes:
            f.write(f"{bits(x)} {bits(y)} {bits(r)} {rel}\n")
        f.write(f"tprobes {len(tp)}\n")
        for k, a, b, dim in tp:
            f.write(f"{k} {bits(a)} {bits(b)} {bits(dim)}\n")


if __name__ == "__main__":
    main()
