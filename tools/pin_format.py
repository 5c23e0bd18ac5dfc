"""Reader / writer of the parity-pinning kit's text files (tests/golden/make_pin_input.py describes the input format).

Output format (written by rust/examples/dump_golden.rs on the real engine and by tools/pin_from_oracle.py on the CPU
restatement; every float as the hex of its bits):
  # flk pin output v1 <producer>
  query <index> <count> <id> ...            ids in the order the query returned them
  tresult <index> <bits>
  state <tick> <n>       then n lines   <id> <x> <y> <last_dx> <last_dy>     (agents in ascending id, after `tick` ticks)
"""
import numpy as np

BIRD = np.dtype([("id", "<u4"), ("x", "<f4"), ("y", "<f4"), ("ldx", "<f4"), ("ldy", "<f4")])


def f32(tok):
    return np.array([int(tok, 16)], dtype=np.uint32).view(np.float32)[0]


def bits(v):
    return "%08x" % int(np.float32(v).view(np.uint32))


def read_input(path):
    lines = [l.split() for l in open(path) if l.strip() and not l.startswith("#")]
    it = iter(lines)
    out = {}
    for tok in it:
        if tok[0] == "field":
            out["W"], out["H"], out["d"], out["toroidal"] = f32(tok[1]), f32(tok[2]), f32(tok[3]), int(tok[4])
        elif tok[0] == "birds":
            b = np.zeros(int(tok[1]), dtype=BIRD)
            for i in range(len(b)):
                t = next(it)
                b[i] = (int(t[0]), f32(t[1]), f32(t[2]), f32(t[3]), f32(t[4]))
            out["birds"] = b
        elif tok[0] == "ticks":
            out["ticks"] = int(tok[1])
        elif tok[0] == "probes":
            out["probes"] = [(lambda t: (f32(t[0]), f32(t[1]), f32(t[2]), int(t[3])))(next(it)) for _ in range(int(tok[1]))]
        elif tok[0] == "tprobes":
            out["tprobes"] = [(lambda t: (int(t[0]), f32(t[1]), f32(t[2]), f32(t[3])))(next(it)) for _ in range(int(tok[1]))]
    return out


def read_output(path):
    out = {"queries": {}, "tresults": {}, "states": {}, "producer": ""}
    lines = [l.rstrip("\n") for l in open(path)]
    k = 0
    while k < len(lines):
        l = lines[k]
        k += 1
        if l.startswith("# flk pin output"):
            out["producer"] = l.split("v1", 1)[-1].strip()
            continue
        tok = l.split()
        if not tok or tok[0].startswith("#"):
            continue
        if tok[0] == "query":
            out["queries"][int(tok[1])] = [int(x) for x in tok[3:3 + int(tok[2])]]
        elif tok[0] == "tresult":
            out["tresults"][int(tok[1])] = int(tok[2], 16)
        elif tok[0] == "state":
            n = int(tok[2])
            b = np.zeros(n, dtype=BIRD)
            for i in range(n):
                t = lines[k + i].split()
                b[i] = (int(t[0]), f32(t[1]), f32(t[2]), f32(t[3]), f32(t[4]))
            k += n
            out["states"][int(tok[1])] = b
    return out


def write_output(path, producer, queries, tresults, states):
    with open(path, "w") as f:
        f.write(f"# flk pin output v1 {producer}\n")
        for i, ids in enumerate(queries):
            f.write(f"query {i} {len(ids)} " + " ".join(str(int(x)) for x in ids) + "\n")
        for i, v in enumerate(tresults):
            f.write(f"tresult {i} {bits(v)}\n")
        for t in sorted(states):
            b = states[t]
            f.write(f"state {t} {len(b)}\n")
            for r in b:
                f.write(f"{int(r['id'])} {bits(r['x'])} {bits(r['y'])} {bits(r['ldx'])} {bits(r['ldy'])}\n")
