"""Independent brute-force restatement (numpy float32 scalars, O(N^2), no bags, no C).

TEST INFRASTRUCTURE ONLY.  A second, structurally different restatement of
flockers/src/model/bird.rs:27-145 + SURVEY.md Appendix B used to cross-check
oracle/flockers_oracle.c (two independent restatements must agree bit for bit).
Instead of bags it derives every neighbour set from first principles: bird j is a
candidate of bird i iff cell(j) lies in the toroidally wrapped window of cell(i).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def tt(v, dim):
    """toroidal_transform (Appendix B)."""
    v, dim = f32(v), f32(dim)
    if v >= f32(0) and v < dim:
        return v
    v = f32(np.fmod(v, dim))
    if v < f32(0):
        v = f32(v + dim)
    return v


def tdist(a, b, dim):
    """toroidal_distance (Appendix B)."""
    a, b, dim = f32(a), f32(b), f32(dim)
    if abs(f32(a - b)) <= f32(dim / f32(2)):
        return f32(a - b)
    d = f32(tt(a, dim) - tt(b, dim))
    if f32(d * f32(2)) > dim:
        return f32(d - dim)
    if f32(d * f32(2)) < -dim:
        return f32(d + dim)
    return d


def cell(x, d):
    q = np.floor(f32(x) / f32(d))
    if np.isnan(q):
        return 0
    return int(q)


def neighbor_ids(birds, i_x, i_y, W, H, d, dist=10.0, relax=True, plus1=False):
    """ids (set semantics) returned by the query at (i_x, i_y) over the given READ population."""
    mx = int(np.ceil(f32(W) / f32(d)))
    my = int(np.ceil(f32(H) / f32(d)))
    k = int(np.floor(f32(dist) / f32(d))) + (1 if plus1 else 0)
    cx, cy = cell(i_x, d), cell(i_y, d)
    wx = {(cx + o) % mx for o in range(-k, k + 1)}
    wy = {(cy + o) % my for o in range(-k, k + 1)}
    out = []
    for b in birds:
        if cell(b["x"], d) in wx and cell(b["y"], d) in wy:
            if not relax:
                ddx = tdist(i_x, b["x"], W)
                ddy = tdist(i_y, b["y"], H)
                if np.sqrt(f32(f32(ddx * ddx) + f32(ddy * ddy))) > f32(dist):
                    continue
            out.append(int(b["id"]))
    return sorted(out)


def tick(birds, W, H, d, r, weights=(0.8, 1.0, 1.1, 0.7, 1.0, 0.7), dist=10.0, relax=True, plus1=False):
    """One synchronous tick over `birds` (structured array, any order); candidates are summed in
    window order (x-outer, y-inner) then ascending id — the canonical order of the C oracle.
    `r[id] = (r1, r2)`.  Returns the new population sorted by id."""
    COH, AVO, RND, CON, MOM, JUMP = (f32(w) for w in weights)
    W, H, d = f32(W), f32(H), f32(d)
    mx = int(np.ceil(W / d))
    my = int(np.ceil(H / d))
    k = int(np.floor(f32(dist) / d)) + (1 if plus1 else 0)
    cells = [(cell(b["x"], d), cell(b["y"], d)) for b in birds]
    out = np.array(birds, copy=True)
    for i, me in enumerate(birds):
        cx, cy = cells[i]
        cand = []
        for ox in range(-k, k + 1):
            for oy in range(-k, k + 1):
                key = ((cx + ox) % mx, (cy + oy) % my)
                grp = [j for j in range(len(birds)) if cells[j] == key]
                grp.sort(key=lambda j: int(birds["id"][j]))
                cand += grp
        if not relax:
            keep = []
            for j in cand:
                ddx = tdist(me["x"], birds["x"][j], W)
                ddy = tdist(me["y"], birds["y"][j], H)
                if np.sqrt(f32(f32(ddx * ddx) + f32(ddy * ddy))) <= f32(dist):
                    keep.append(j)
            cand = keep
        avo = [f32(0), f32(0)]
        coh = [f32(0), f32(0)]
        rnd = [f32(0), f32(0)]
        con = [f32(0), f32(0)]
        if cand:
            xa = ya = xc = yc = xs = ys = f32(0)
            count = 0
            for j in cand:
                if int(birds["id"][j]) == int(me["id"]):
                    continue
                dx = tdist(me["x"], birds["x"][j], W)
                dy = tdist(me["y"], birds["y"][j], H)
                count += 1
                sq = f32(f32(dx * dx) + f32(dy * dy))
                den = f32(f32(sq * sq) + f32(1))
                xa = f32(xa + f32(dx / den))
                ya = f32(ya + f32(dy / den))
                xc = f32(xc + dx)
                yc = f32(yc + dy)
                xs = f32(xs + birds["ldx"][j])
                ys = f32(ys + birds["ldy"][j])
            if count > 0:
                c = f32(count)
                xa, ya, xc, yc, xs, ys = (f32(v / c) for v in (xa, ya, xc, yc, xs, ys))
                con = [f32(xs / c), f32(ys / c)]
            else:
                con = [xs, ys]
            avo = [f32(f32(400) * xa), f32(f32(400) * ya)]
            coh = [f32(f32(-xc) / f32(10)), f32(f32(-yc) / f32(10))]
            r1, r2 = f32(r[int(me["id"])][0]), f32(r[int(me["id"])][1])
            xr = f32(f32(r1 * f32(2)) - f32(1))
            yr = f32(f32(r2 * f32(2)) - f32(1))
            with np.errstate(invalid="ignore", divide="ignore"):
                s = np.sqrt(f32(f32(xr * xr) + f32(yr * yr)))
                rnd = [f32(f32(f32(0.05) * xr) / s), f32(f32(f32(0.05) * yr) / s)]
        mom = [f32(me["ldx"]), f32(me["ldy"])]
        dd = []
        for a in (0, 1):
            v = f32(COH * coh[a])
            v = f32(v + f32(AVO * avo[a]))
            v = f32(v + f32(CON * con[a]))
            v = f32(v + f32(RND * rnd[a]))
            v = f32(v + f32(MOM * mom[a]))
            dd.append(v)
        dis = np.sqrt(f32(f32(dd[0] * dd[0]) + f32(dd[1] * dd[1])))
        if dis > f32(0):
            dd = [f32(f32(dd[0] / dis) * JUMP), f32(f32(dd[1] / dis) * JUMP)]
        out["ldx"][i], out["ldy"][i] = dd[0], dd[1]
        out["x"][i] = tt(f32(me["x"] + dd[0]), W)
        out["y"][i] = tt(f32(me["y"] + dd[1]), H)
    return out[np.argsort(out["id"], kind="stable")]
