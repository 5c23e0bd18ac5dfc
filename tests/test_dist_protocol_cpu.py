"""CPU test of the N>1 path (world_size 2 and 3 over gloo): the halo + migration protocol of flockers_mpi
(flockers_mpi/src/model/state.rs:105-175, bird.rs:185-195), driven through the host-side partition helpers of
examples_b200.distributed with the CPU oracle as the per-rank compute, must reproduce the single-process
oracle bit for bit.  This is the protocol the CUDA strips implement (tests/test_gpu_strips.py checks those)."""
import os
import socket

import numpy as np
import pytest

from tests.helpers import D, bits_equal, uniform_birds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _exchange(dist, torch, O, rank, world, out_left, out_right):
    """send one bird array to each ring neighbour, receive theirs (two messages per rank per call)"""
    left, right = (rank - 1) % world, (rank + 1) % world

    def to_t(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).copy())

    res = {}
    # sizes first, then payloads; order recvs so that world==2 pairs sendL with the peer's recvR
    sl, sr = to_t(out_left), to_t(out_right)
    nl, nr = torch.tensor([sl.numel()]), torch.tensor([sr.numel()])
    rl, rr = torch.zeros(1, dtype=torch.int64), torch.zeros(1, dtype=torch.int64)
    ops = [dist.P2POp(dist.isend, nl, left), dist.P2POp(dist.isend, nr, right),
           dist.P2POp(dist.irecv, rr, right), dist.P2POp(dist.irecv, rl, left)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    bl, br = torch.zeros(int(rl), dtype=torch.uint8), torch.zeros(int(rr), dtype=torch.uint8)
    ops = [dist.P2POp(dist.isend, sl, left), dist.P2POp(dist.isend, sr, right),
           dist.P2POp(dist.irecv, br, right), dist.P2POp(dist.irecv, bl, left)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    res["from_left"] = bl.numpy().view(O.BIRD).copy()
    res["from_right"] = br.numpy().view(O.BIRD).copy()
    return res


def _worker(rank, world, port, n, W, ticks, q):
    import torch
    import torch.distributed as dist
    from examples_b200.distributed import column_of, grid_columns, owner_of_column, strip_of
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    birds = uniform_birds(n, W, W)
    mx = grid_columns(W, D)
    c0, c1 = strip_of(mx, world, rank)
    # init scatter (state.rs:51-103): every rank keeps the birds whose column it owns
    cols = column_of(birds["x"], D)
    mine = np.array([owner_of_column(mx, world, int(c)) == rank for c in cols])
    sim = O.Sim(W, W)
    sim.init(birds[mine])
    sim.field.lazy_update()                       # Schedule::step: update(0) before tick 0
    seed = 0xB12D
    for t in range(ticks):
        # before_step (state.rs:113-133): my boundary columns become the neighbours' ghosts
        halo = _exchange(dist, torch, O, rank, world, sim.field.columns(c0, c0 + 1), sim.field.columns(c1 - 1, c1))
        sim.field.insert_read(halo["from_left"])
        sim.field.insert_read(halo["from_right"])
        # every owned Bird::step; leavers reported by owner (bird.rs:190-195)
        leavers = sim.step_strip(t, seed, None, c0, c1)
        lcol = column_of(leavers["x"], D)
        dest = np.array([owner_of_column(mx, world, int(c)) for c in lcol], dtype=np.int64)
        left, right = (rank - 1) % world, (rank + 1) % world
        to_left = leavers[dest == left] if world > 2 else leavers[(dest == left) & (lcol >= c1 if rank == 0 else lcol < c0)]
        to_right = leavers[dest == right] if world > 2 else leavers[(dest == right) & ~((lcol >= c1) if rank == 0 else (lcol < c0))]
        assert len(to_left) + len(to_right) == len(leavers)        # JUMP < d: leavers only reach adjacent strips
        # after_step (state.rs:139-175): migrate
        got = _exchange(dist, torch, O, rank, world, to_left, to_right)
        sim.add_agents(got["from_left"])
        sim.add_agents(got["from_right"])
        sim.field.lazy_update()                   # update (state.rs:105-107)
    mine_now = sim.agents()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine_now.view(np.uint8).tobytes())
    if rank == 0:
        allb = np.concatenate([np.frombuffer(g, dtype=O.BIRD) for g in gathered])
        q.put(np.sort(allb, order="id").view(np.uint8).tobytes())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_and_migration_protocol_matches_single_process_oracle(world):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    O.build()
    n, W, ticks = 1500, 100.0, 25
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, W, ticks, q)) for r in range(world)]
    for p in procs:
        p.start()
    data = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = np.frombuffer(data, dtype=O.BIRD)
    sim = O.Sim(W, W)
    sim.init(uniform_birds(n, W, W))
    for _ in range(ticks):
        sim.step()
    assert len(got) == n
    assert bits_equal(got, sim.agents())


def test_partition_helpers():
    from examples_b200.distributed import grid_columns, owner_of_column, strip_of
    assert grid_columns(100.0, D) == 15 and grid_columns(35840.0, D) == 5376
    for mx, world in ((15, 2), (15, 4), (1920, 8), (5376, 8), (17, 3)):
        edges = [strip_of(mx, world, r) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == mx
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        assert all(c1 - c0 >= 2 for c0, c1 in edges)
        for c in range(mx):
            r = owner_of_column(mx, world, c)
            assert edges[r][0] <= c < edges[r][1]
        assert owner_of_column(mx, world, mx) == 0        # slack column -> owner of column 0


# ---- tiles: model check of the two-hop exchange (the routing rules of emit_successor / k_unpack in flk_kernels.cu) ----

def _local(g0, gown, m, has_slack, gc):
    """global cell coordinate -> local index of a tile axis (0 ghost, 1..own owned, own+1 ghost, own+2 slack) or None;
    the same arithmetic as local_col / local_row"""
    if gc >= m:
        return gown + 2 if has_slack else None
    rel = gc - g0
    if rel < -1:
        rel += m
    if rel > gown:
        rel -= m
    if rel < -1 or rel > gown:
        return None
    return rel + 1


def _sides(l, own):
    """which of my two neighbours along one axis get a bird that landed at local index l (messages 'lo' / 'hi')"""
    out = set()
    if l in (0, 1):
        out.add("lo")
    if l in (own, own + 1):
        out.add("hi")
    return out


@pytest.mark.parametrize("Px,Py,mx,my", [(2, 2, 9, 8), (3, 2, 13, 7), (2, 3, 8, 12), (4, 3, 17, 13)])
def test_tile_two_hop_routing_delivers_every_ghost_exactly_once(Px, Py, mx, my):
    """For every tile T and every cell a successor of one of T's birds can land in (T's owned cells +-1, wrapped), the
    set of tiles that end up holding the bird after 'left/right exchange, forward border rows, down/up exchange'
    must be exactly the tiles whose local grid (owned + ghost ring) contains that cell, each reached once."""
    cb = [px * mx // Px for px in range(Px)] + [mx]
    rb = [py * my // Py for py in range(Py)] + [my]

    def geom(px, py):
        return cb[px], cb[px + 1] - cb[px], rb[py], rb[py + 1] - rb[py]

    def holds(px, py, gx, gy):
        c0, nown, r0, rown = geom(px, py)
        lc, lr = _local(c0, nown, mx, c0 == 0, gx), _local(r0, rown, my, r0 == 0, gy)
        return lc is not None and lr is not None and lc <= nown + 1 and lr <= rown + 1   # slack cells are not in play here

    for px in range(Px):
        for py in range(Py):
            c0, nown, r0, rown = geom(px, py)
            for ox in range(-1, nown + 1):                 # landing cells: owned region +-1 ring, global (wrapped)
                for oy in range(-1, rown + 1):
                    gx, gy = (c0 + ox) % mx, (r0 + oy) % my
                    expect = {(qx, qy) for qx in range(Px) for qy in range(Py) if holds(qx, qy, gx, gy)}
                    got = []                                # (tile) deliveries, duplicates kept
                    lc, lr = _local(c0, nown, mx, c0 == 0, gx), _local(r0, rown, my, r0 == 0, gy)
                    assert lc is not None and lr is not None
                    got.append((px, py))                    # emit_successor keeps it (key valid)
                    hop1 = []
                    for sx in _sides(lc, nown):             # left/right messages
                        qx = (px - 1) % Px if sx == "lo" else (px + 1) % Px
                        hop1.append((qx, py))
                    for sy in _sides(lr, rown):             # down/up messages with my own successors
                        qy = (py - 1) % Py if sy == "lo" else (py + 1) % Py
                        got.append((px, qy))
                    for (qx, qy) in hop1:                   # k_unpack phase 0 at the receiver: keep + forward by ITS rows
                        qc0, qn, qr0, qrn = geom(qx, qy)
                        l2c, l2r = _local(qc0, qn, mx, qc0 == 0, gx), _local(qr0, qrn, my, qr0 == 0, gy)
                        assert l2c is not None and l2r is not None
                        got.append((qx, qy))
                        for sy in _sides(l2r, qrn):
                            got.append((qx, (qy - 1) % Py if sy == "lo" else (qy + 1) % Py))
                    # with two tiles along an axis both neighbours are the same tile and it sees the cell on both sides
                    # only if the axis is so short that ghost rings overlap; the grids used here avoid that
                    assert sorted(set(got)) == sorted(got), (px, py, gx, gy, got)
                    assert set(got) == expect, (px, py, gx, gy, sorted(got), sorted(expect))


# ---- tiles over gloo: the two-hop halo + migration protocol with the CPU oracle as the per-rank compute --------------

def _exchange_pair(dist, torch, O, lo_rank, hi_rank, out_lo, out_hi):
    """one message to the neighbour on the low side and one to the high side of an axis; returns (from_lo, from_hi).
    With two tiles along the axis both neighbours are the same rank: the peer's first message (its low-side one) is
    what arrives on my high side, exactly the ordering rule of the NCCL exchange."""
    def to_t(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).copy())

    sl, sh = to_t(out_lo), to_t(out_hi)
    nl, nh = torch.tensor([sl.numel()]), torch.tensor([sh.numel()])
    rl, rh = torch.zeros(1, dtype=torch.int64), torch.zeros(1, dtype=torch.int64)
    for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, nl, lo_rank), dist.P2POp(dist.isend, nh, hi_rank),
                                     dist.P2POp(dist.irecv, rh, hi_rank), dist.P2POp(dist.irecv, rl, lo_rank)]):
        w.wait()
    bl, bh = torch.zeros(int(rl), dtype=torch.uint8), torch.zeros(int(rh), dtype=torch.uint8)
    for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, sl, lo_rank), dist.P2POp(dist.isend, sh, hi_rank),
                                     dist.P2POp(dist.irecv, bh, hi_rank), dist.P2POp(dist.irecv, bl, lo_rank)]):
        w.wait()
    return bl.numpy().view(O.BIRD).copy(), bh.numpy().view(O.BIRD).copy()


def _tile_worker(rank, Px, Py, port, n, W, H, ticks, q):
    import torch
    import torch.distributed as dist
    from examples_b200.distributed import column_of, grid_columns
    from oracle import oracle as O
    world = Px * Py
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    px, py = rank % Px, rank // Px
    mx, my = grid_columns(W, D), grid_columns(H, D)
    c0, c1 = px * mx // Px, (px + 1) * mx // Px
    r0, r1 = py * my // Py, (py + 1) * my // Py
    left, right = py * Px + (px - 1) % Px, py * Px + (px + 1) % Px
    down, up = ((py - 1) % Py) * Px + px, ((py + 1) % Py) * Px + px
    birds = uniform_birds(n, W, H)
    cx, cy = column_of(birds["x"], D), column_of(birds["y"], D)
    mine = (((cx >= c0) & (cx < c1)) | ((cx >= mx) & (c0 == 0))) & (((cy >= r0) & (cy < r1)) | ((cy >= my) & (r0 == 0)))
    sim = O.Sim(W, H)
    sim.init(birds[mine])
    sim.field.lazy_update()
    F = sim.field
    cat = lambda parts: np.concatenate(parts) if parts else np.zeros(0, dtype=O.BIRD)
    seed = 0xB12D
    for t in range(ticks):
        # before_step, hop 1: my first / last owned column (owned rows) become the left / right neighbour's ghosts
        fl, fr = _exchange_pair(dist, torch, O, left, right, F.cells(c0, c0 + 1, r0, r1), F.cells(c1 - 1, c1, r0, r1))
        F.insert_read(fl); F.insert_read(fr)
        # hop 2: my first / last owned row INCLUDING the ghost columns just received (the diagonal neighbours' cells)
        gl, gr = (c0 - 1) % mx, c1 % mx
        row = lambda r: cat([F.cells(gl, gl + 1, r, r + 1), F.cells(c0, c1, r, r + 1), F.cells(gr, gr + 1, r, r + 1)])
        fd, fu = _exchange_pair(dist, torch, O, down, up, row(r0), row(r1 - 1))
        F.insert_read(fd); F.insert_read(fu)
        # every owned Bird::step; a bird that leaves the block is reported, not kept
        leavers = sim.step_tile(t, seed, None, c0, c1, r0, r1)
        lx, ly = column_of(leavers["x"], D), column_of(leavers["y"], D)
        assert np.all(lx < mx) and np.all(ly < my)                 # the slack cells are reached from their own block only
        go_l, go_r = lx == (c0 - 1) % mx, lx == c1 % mx
        assert np.all(go_l | go_r | ((lx >= c0) & (lx < c1)))       # JUMP < d: one cell at most
        # after_step, hop 1: along x (whatever the row), hop 2: along y, forwarding what hop 1 brought in
        al, ar = _exchange_pair(dist, torch, O, left, right, leavers[go_l], leavers[go_r])
        pool = cat([leavers[~go_l & ~go_r], al, ar])
        py_ = column_of(pool["y"], D)
        go_d, go_u = py_ == (r0 - 1) % my, py_ == r1 % my
        ad, au = _exchange_pair(dist, torch, O, down, up, pool[go_d], pool[go_u])
        arrivals = cat([pool[~go_d & ~go_u], ad, au])
        acx, acy = column_of(arrivals["x"], D), column_of(arrivals["y"], D)
        assert np.all((acx >= c0) & (acx < c1) & (acy >= r0) & (acy < r1))
        sim.add_agents(arrivals)
        sim.field.lazy_update()
    gathered = [None] * world
    dist.all_gather_object(gathered, sim.agents().view(np.uint8).tobytes())
    if rank == 0:
        allb = np.concatenate([np.frombuffer(g, dtype=O.BIRD) for g in gathered])
        q.put(np.sort(allb, order="id").view(np.uint8).tobytes())
    dist.destroy_process_group()


@pytest.mark.parametrize("Px,Py", [(2, 2), (3, 2)])
def test_tile_two_hop_protocol_matches_single_process_oracle(Px, Py):
    """world_size 4 and 6 over gloo: 2-D blocks, ghosts and migrants of the diagonal neighbours in two hops"""
    import torch.multiprocessing as mp
    from oracle import oracle as O
    O.build()
    n, W, H, ticks = 1800, 120.0, 100.0, 20
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tile_worker, args=(r, Px, Py, port, n, W, H, ticks, q)) for r in range(Px * Py)]
    for p in procs:
        p.start()
    data = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = np.frombuffer(data, dtype=O.BIRD)
    sim = O.Sim(W, H)
    sim.init(uniform_birds(n, W, H))
    for _ in range(ticks):
        sim.step()
    assert len(got) == n
    assert bits_equal(got, sim.agents())
