#!/usr/bin/env python
"""bench.py — bird-updates/s of the Flockers hot path (Bird::step + Field2D lazy_update) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm on the host cores

A "step" is one tick: every bird's Bird::step (flockers/src/model/bird.rs:27-145) followed by the Field2D
lazy_update re-binning (flockers/src/model/state.rs:54-56).  Prints ONE JSON line on rank 0.

Workloads (BASELINE.json configs):
    weak16m   (default) 16 M birds per GPU, 672 cell columns x 5376 rows per GPU, uniform random, density ~0.1
              (config 4 "weak scaling 16M birds per GPU"; at N=1 it is the 16 M-bird single-B200 target case)
    strong16m 16 M birds total on 1920 x 1920 cells (config 3)
    1m        1 M birds, 480 x 480 cells (config 2)
    clustered 16 M birds per GPU in gaussian flocks (config 5)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D = float(np.float32(10.0) / np.float32(1.5))
INIT_SEED, STEP_SEED = 0xF10C, 0xB12D
ALGO_BYTES_PER_UPDATE = 40  # one 20-byte Bird read + one written per bird-update


def workload_geometry(name: str, world: int, birds_per_gpu: int | None):
    """-> (cells_x, cells_y, n_total, label)"""
    if name == "weak16m":
        per = birds_per_gpu or 16_000_000
        gx, gy = 672 * world, 5376
        if per != 16_000_000:
            scale = (per / 16_000_000) ** 0.5
            gx, gy = max(int(672 * scale), 3) * world, max(int(5376 * scale), 3)
        return gx, gy, per * world, f"flockers weak scaling, {per} birds/GPU, {gx}x{gy} cells, uniform random, density~0.1"
    if name == "strong16m":
        return 1920, 1920, 16_000_000, "flockers 16M birds strong scaling, 1920x1920 cells, uniform random"
    if name == "1m":
        return 480, 480, 1_000_000, "flockers 1M birds, 480x480 cells, uniform random"
    if name == "clustered":
        per = birds_per_gpu or 16_000_000
        return 672 * world, 5376, per * world, f"flockers clustered flocks, {per} birds/GPU, {672 * world}x5376 cells"
    raise SystemExit(f"unknown workload {name}")


def field_size(cells: int) -> float:
    """W such that f32 ceil(W/d) == cells (W = cells*d rounds correctly for the sizes used; checked)."""
    w = float(np.float32(np.float32(cells) * np.float32(D)))
    assert int(np.ceil(np.float32(w) / np.float32(D))) == cells, (cells, w)
    return w


def make_birds(n: int, W: float, H: float, clustered: bool, x0: float = 0.0, x1: float | None = None, seed: int = INIT_SEED):
    """Synthetic init (flockers/src/model/state.rs:41-48: loc = dim * rand, last_d = 0) restricted to x in [x0,x1)."""
    rng = np.random.Generator(np.random.Philox(seed))
    x1 = W if x1 is None else x1
    if not clustered:
        loc = np.empty((n, 2), dtype=np.float32)
        loc[:, 0] = (x0 + (x1 - x0) * rng.random(n, dtype=np.float32)).astype(np.float32)
        loc[:, 1] = (H * rng.random(n, dtype=np.float32)).astype(np.float32)
    else:
        k = max(n // 4096, 1)
        c = rng.random((k, 2)) * [x1 - x0, H] + [x0, 0]
        which = rng.integers(0, k, n)
        p = c[which] + rng.normal(0.0, 10.0, (n, 2))
        p[:, 0] = x0 + np.mod(p[:, 0] - x0, x1 - x0)
        p[:, 1] = np.mod(p[:, 1], H)
        loc = p.astype(np.float32)
    loc[:, 0] = np.minimum(loc[:, 0], np.nextafter(np.float32(x1), np.float32(0)))
    loc[:, 1] = np.minimum(loc[:, 1], np.nextafter(np.float32(H), np.float32(0)))
    return loc


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([p.strip() for p in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
            except Exception:
                continue
            for nme, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_launch(birds_per_launch: int, key: str = "bytes_per_launch"):
    """DRAM bytes (or, with key="warp_instructions_per_launch", warp-instructions) of one k_step launch from the
    committed ncu --set full capture (profiles/), or None when the capture was taken on a different launch size."""
    p = os.path.join(ROOT, "profiles", "k_step_traffic.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            if int(d["birds_per_launch"]) == int(birds_per_launch):
                return float(d[key])
        except Exception:
            return None
    return None


def issue_roofline(warp_instr, kernel_ms: float, sm_count: int, sm_mhz):
    """What actually bounds k_step (DESIGN.md roofline): warp-instructions issued per second against the issue peak of
    the device, 4 schedulers per SM x 1 instruction per cycle at the SM clock sampled during the timed region."""
    if not warp_instr or not sm_mhz:
        return None
    achieved = warp_instr / (kernel_ms * 1e-3) / 1e9
    peak = sm_count * 4 * sm_mhz * 1e6 / 1e9
    return {"achieved": achieved, "peak": peak, "unit": "G warp-instructions/s", "frac": achieved / peak,
            "warp_instructions_per_launch": warp_instr,
            "what": "ncu smsp__inst_executed.sum per k_step launch (committed capture) / live launch duration; "
                    "peak = SMs x 4 schedulers x sampled SM clock"}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle (a port of the reference algorithm) on the host cores
# --------------------------------------------------------------------------------------------------

def cpu_oracle_run(n_birds: int, ticks: int, warm: int, clustered: bool, threads: int | None = None):
    threads = threads or (os.cpu_count() or 1)
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is entitled to all host threads, and libgomp
    # reads the variable when the oracle library is loaded (below)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from oracle import oracle as O
    O.build()
    cells = max(int(round((n_birds / 0.0996) ** 0.5 / D)), 3)
    W = field_size(cells)
    loc = make_birds(n_birds, W, W, clustered)
    b = np.zeros(n_birds, dtype=O.BIRD)
    b["id"] = np.arange(n_birds, dtype=np.uint32)
    b["x"], b["y"] = loc[:, 0], loc[:, 1]
    sim = O.Sim(W, W)
    sim.init(b)
    for _ in range(warm):
        sim.step(seed=STEP_SEED, threads=threads)
    t0 = time.perf_counter()
    for _ in range(ticks):
        sim.step(seed=STEP_SEED, threads=threads)
    dt = time.perf_counter() - t0
    return n_birds * ticks / dt, dt, threads, f"{n_birds} birds on {cells}x{cells} cells, {ticks} ticks (+{warm} warm-up), OpenMP over agents"


def run_reference(args, world: int, rank: int):
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank, and libgomp started that way runs the oracle's parallel regions
    # several times slower even when they ask for more threads.  The reference arm is entitled to all host threads:
    # re-run it in a child process whose environment says so from the start.
    want = str(os.cpu_count() or 1)
    if os.environ.get("OMP_NUM_THREADS", want) != want and not os.environ.get("FLK_REFERENCE_CHILD"):
        env = dict(os.environ, OMP_NUM_THREADS=want, FLK_REFERENCE_CHILD="1")
        res = subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env)
        if res.returncode != 0:
            raise SystemExit(res.returncode)
        return
    gx, gy, n_total, label = workload_geometry(args.workload, world, args.birds_per_gpu)
    n_sample = min(n_total, args.cpu_birds)
    value, dt, threads, sample = cpu_oracle_run(n_sample, args.steps, args.warmup, args.workload == "clustered")
    line = {
        "impl": "reference", "metric": "bird-updates/sec", "value": value, "unit": "bird-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": label, "note": "reference algorithm on host cores, bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": "bird-updates/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "bird-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------

def run_ours(args, world: int, rank: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from examples_b200 import Field2D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA path is the only implementation (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    gx, gy, n_total, label = workload_geometry(args.workload, world, args.birds_per_gpu)
    W, H = field_size(gx), field_size(gy)
    clustered = args.workload == "clustered"
    n_local = n_total // world
    cap = int(n_local * 1.25) + 65536

    # ---- build this rank's field and its birds (each rank generates the birds of its own strip) ----
    if world == 1:
        f = Field2D(W, H, D, True, capacity=cap, device=local_rank)
        loc = make_birds(n_local, W, H, clustered)
        id_base = 0
    else:
        from examples_b200 import distributed as fdist
        f = fdist.DistField2D(W, H, D, capacity=cap, device=local_rank, rank=rank, world=world, tile_rows=args.tile_rows)
        c0, c1, r0, r1 = f.tile()
        x0 = float(np.float32(c0) * np.float32(D))
        x1 = float(np.float32(c1) * np.float32(D)) if c1 < gx else W
        loc = make_birds(n_local, W, H, clustered, x0, x1, seed=INIT_SEED + rank)
        # f32 rounding at the strip edges: keep every generated bird inside the columns this rank owns
        col = fdist.column_of(loc[:, 0], D)
        off = (col < c0) | (col >= c1)
        loc[off, 0] = np.float32(0.5 * (x0 + x1))
        if args.tile_rows > 1:   # tiles: squeeze the y coordinates into the rows this rank owns
            y0 = float(np.float32(r0) * np.float32(D))
            y1 = float(np.float32(r1) * np.float32(D)) if r1 < gy else H
            loc[:, 1] = (np.float32(y0) + np.float32(y1 - y0) * (loc[:, 1] / np.float32(H))).astype(np.float32)
            row = fdist.column_of(loc[:, 1], D)
            off = (row < r0) | (row >= r1)
            loc[off, 1] = np.float32(0.5 * (y0 + y1))
        id_base = rank * n_local
    # pinned host buffers: the host-side copy of the agents (what krABMaga's Schedule holds)
    h_id = torch.arange(id_base, id_base + n_local, dtype=torch.int64).to(torch.uint32).pin_memory() \
        if hasattr(torch, "uint32") else None
    h_loc = torch.from_numpy(loc).pin_memory()
    h_ld = torch.zeros((n_local, 2), dtype=torch.float32).pin_memory()
    if h_id is None:
        h_id = torch.from_numpy(np.arange(id_base, id_base + n_local, dtype=np.uint32).view(np.int32)).pin_memory()
    f.set_objects_raw(n_local, h_id.data_ptr(), h_loc.data_ptr(), h_ld.data_ptr())
    f.set_math_mode(args.math)
    f.commit()  # lazy_update (+ first halo exchange on >1 GPU)
    f.synchronize()
    assert f.num_owned() == n_local, (f.num_owned(), n_local)

    # ---- warm-up ----
    f.run(0, args.warmup, STEP_SEED)
    f.synchronize()

    # ---- timed region: K ticks, device-timed on the handle's stream, step kernel bracketed by events ----
    launches0 = f.launch_count()
    f.profile_enable(1)   # CUDA events around every k_step launch of the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    f.run(args.warmup, args.steps, STEP_SEED)
    f.synchronize()
    barrier()
    ms = max_over_ranks(f.elapsed_ms())
    clocks = sampler.stop() if rank == 0 else None
    k_ms, k_n = f.profile_step_kernel()
    f.profile_enable(False)
    launches = f.launch_count() - launches0
    # the other kernels of a tick (scan, scatter, rank), timed over a few more ticks outside the timed region
    f.profile_enable(2)
    f.run(args.warmup + args.steps, min(args.steps, 20), STEP_SEED)
    f.synchronize()
    rebin_ms = f.profile_rebin()
    f.profile_enable(False)
    n_now = sum_over_ranks(float(f.num_owned()))
    value = n_total * args.steps / (ms * 1e-3)

    # roofline of the dominant kernel (k_step): algorithmic bytes per launch / average launch duration
    k_avg_ms = max_over_ranks(k_ms / max(k_n, 1))
    peak, peak_src = measured_peak_gbs()
    achieved = ALGO_BYTES_PER_UPDATE * n_local / (k_avg_ms * 1e-3) / 1e9
    traffic = ncu_traffic_per_launch(n_total // world)
    warp_instr = ncu_traffic_per_launch(n_total // world, "warp_instructions_per_launch")
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count

    # ---- e2e: per step, host -> device copy of every agent, one tick, device -> host read of the new state ----
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    e2e_extra = {}
    if world == 1:
        # The host keeps the agents as a list indexed by id (flockers/src/model/state.rs:40-51), so the dense C-ABI
        # calls move Bird minus its id: 16 B per bird each way.  Every step of every repetition uploads all agents
        # from pinned host memory, publishes them (lazy_update), runs Bird::step + lazy_update and downloads all
        # agents; the download is the next step's upload (a dependent chain per repetition).  `reps` independent
        # repetitions (simulate!'s third argument, flockers/src/main.rs:47) are pipelined over one handle each, so
        # one repetition's download overlaps the next one's upload on the two PCIe directions.
        rec0 = np.concatenate([loc, np.zeros_like(loc)], axis=1)

        def e2e_run(fields, steps):
            R = len(fields)
            h_in = [torch.from_numpy(rec0.copy()).pin_memory() for _ in range(R)]
            h_out = [torch.empty((n_local, 4), dtype=torch.float32).pin_memory() for _ in range(R)]
            assert all(t.is_pinned() for t in h_in + h_out)
            t0 = 0.0
            for it in range(steps + 1):          # iteration 0 is untimed (first-touch, graph/alloc warm-up)
                if it == 1:
                    for fr in fields:
                        fr.synchronize()
                    barrier()
                    t0 = time.perf_counter()
                for r, fr in enumerate(fields):
                    fr.synchronize()             # this repetition's previous download has landed
                    fr.set_objects_dense_raw(n_local, h_in[r].data_ptr())
                    fr.commit()
                    fr.run(1000 + it, 1, STEP_SEED + r)
                    fr.snapshot_dense_async_raw(n_local, h_out[r].data_ptr())
                    h_in[r], h_out[r] = h_out[r], h_in[r]
            for fr in fields:
                fr.synchronize()
            barrier()
            dt = time.perf_counter() - t0
            for r in range(R):                   # every agent came back (no NaN rows), and it moved
                assert not bool(torch.isnan(h_in[r][:: max(n_local // 4096, 1)]).any())
            return R * n_local * steps / dt

        serial_value = e2e_run([f], e2e_steps)
        reps = max(1, args.e2e_reps)
        extra = [Field2D(W, H, D, True, capacity=cap, device=local_rank) for _ in range(reps - 1)]
        for fx in extra:
            fx.set_math_mode(args.math)
        e2e_value = e2e_run([f] + extra, e2e_steps)
        for fx in extra:
            fx.close()
        e2e_bytes = 16 * n_local
        e2e_extra = {"reps": reps, "serial_value": serial_value,
                     "what": f"{reps} independent repetitions pipelined over {reps} handles; per step and repetition: "
                             "flk_set_objects_dense(pinned host, 16 B/bird) + lazy_update + Bird::step + lazy_update + "
                             "flk_snapshot_dense_async(pinned host, 16 B/bird); output feeds the next step's input; "
                             "serial_value = one repetition alone (no overlap)"}
    else:
        # One strip per rank: the birds a rank owns change with migration, so ids travel with the records (20 B per
        # bird each way).  `reps` independent repetitions, each its own group of strip handles (own NCCL communicator),
        # are pipelined like the single-GPU case; every rank walks the repetitions in the same order (collectives).
        reps = max(1, args.e2e_reps)
        cap_h = n_local + n_local // 8 + 65536

        def e2e_run_dist(fields, steps):
            R = len(fields)
            bufs = []
            for r in range(R):
                pair = []
                for _ in range(2):
                    pair.append((torch.empty(cap_h, dtype=torch.int32).pin_memory(),
                                 torch.empty((cap_h, 2), dtype=torch.float32).pin_memory(),
                                 torch.empty((cap_h, 2), dtype=torch.float32).pin_memory()))
                pair[0][0][:n_local].copy_(torch.from_numpy(np.arange(id_base, id_base + n_local, dtype=np.uint32).view(np.int32)))
                pair[0][1][:n_local].copy_(torch.from_numpy(loc))
                pair[0][2][:n_local].zero_()
                bufs.append(pair)
            counts = [n_local] * R
            t0 = 0.0
            for it in range(steps + 1):
                if it == 1:
                    for fr in fields:
                        fr.synchronize()
                    barrier()
                    t0 = time.perf_counter()
                for r, fr in enumerate(fields):
                    fr.synchronize()
                    (i_id, i_loc, i_ld), (o_id, o_loc, o_ld) = bufs[r]
                    fr.set_objects_raw(counts[r], i_id.data_ptr(), i_loc.data_ptr(), i_ld.data_ptr())
                    fr.commit()
                    fr.run(1000 + it, 1, STEP_SEED + r)
                    counts[r] = fr.snapshot_owned_async_raw(o_id.data_ptr(), o_loc.data_ptr(), o_ld.data_ptr())
                    assert counts[r] <= cap_h, (counts[r], cap_h)
                    bufs[r] = [bufs[r][1], bufs[r][0]]
            for fr in fields:
                fr.synchronize()
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            total = sum_over_ranks(float(sum(counts)))
            assert int(total) == R * n_total, (total, R * n_total)
            return R * n_total * steps / dt

        extra = [fdist.DistField2D(W, H, D, capacity=cap, device=local_rank, rank=rank, world=world, tile_rows=args.tile_rows)
                 for _ in range(reps - 1)]
        for fx in extra:
            fx.set_math_mode(args.math)
        serial_value = e2e_run_dist([f], e2e_steps)
        e2e_value = e2e_run_dist([f] + extra, e2e_steps)
        for fx in extra:
            fx.close()
        e2e_bytes = 20 * (n_total // world)
        e2e_extra = {"reps": reps, "serial_value": serial_value,
                     "what": f"{reps} independent repetitions pipelined over {reps} groups of strip handles; per rank, step and "
                             "repetition: flk_set_objects(pinned host, 20 B/bird) + commit + Bird::step + exchange + "
                             "lazy_update + flk_dist_snapshot_owned_async(pinned host, 20 B/bird); output feeds the next "
                             "step's input; bytes are per rank; serial_value = one repetition alone"}

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle on the host cores, bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, threads, sample = cpu_oracle_run(args.cpu_birds, args.cpu_ticks, 1, clustered)
        cpu = {"value": v, "unit": "bird-updates/s", "cores": threads, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "bird-updates/sec", "value": value, "unit": "bird-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.workload == "strong16m" else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": label, "birds_total": n_total, "birds_per_gpu": n_total // world,
                       "decomposition": "single GPU" if world == 1 else
                                        (f"x-strips over {world} GPUs, NCCL halo+migration" if args.tile_rows == 1 else
                                         f"{world // args.tile_rows} x {args.tile_rows} tiles over {world} GPUs, NCCL halo+migration in two hops"),
                       "math": "exact (IEEE f32, bit-identical to the oracle)" if args.math == 0 else "fast",
                       "l2": "inputs larger than L2 (>=20 B/bird x birds per GPU vs 126 MB)" if n_total // world >= 8_000_000
                             else "working set fits L2 (reference config size)",
                       "birds_after": int(n_now)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_step", "kernel_ms": k_avg_ms,
                         "kernel_share_of_step": k_avg_ms / (ms / args.steps), "peak_source": peak_src,
                         "other_kernels_ms": {"k_scan_x3": rebin_ms[0], "k_scatter": rebin_ms[1], "k_rank": rebin_ms[2]},
                         "algorithmic_bytes_per_update": ALGO_BYTES_PER_UPDATE,
                         "issue_slots": issue_roofline(warp_instr, k_avg_ms, sm_count, (clocks or {}).get("sm_mhz")),
                         "note": "path is FP32-issue bound, not HBM bound (DESIGN.md roofline): see issue_slots"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "bird-updates/s", "h2d_bytes_per_step": e2e_bytes,
                    "d2h_bytes_per_step": e2e_bytes, "steps": e2e_steps, **e2e_extra},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    f.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="weak16m", choices=["weak16m", "strong16m", "1m", "clustered"])
    ap.add_argument("--birds-per-gpu", type=int, default=None)
    ap.add_argument("--math", type=int, default=0)
    ap.add_argument("--tile-rows", type=int, default=1)   # > 1: split the rows too (tiles); default = x-strips
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-reps", type=int, default=3)           # independent repetitions pipelined in the e2e leg
    ap.add_argument("--cpu-birds", type=int, default=1_000_000)   # birds of the CPU sample (per tick)
    ap.add_argument("--cpu-ticks", type=int, default=200)        # ~10 s of CPU work on 16 cores
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun for --gpus > 1 (one rank per GPU)")
    if args.impl == "reference":
        run_reference(args, world, rank)
    else:
        run_ours(args, world, rank, local_rank)


if __name__ == "__main__":
    main()
