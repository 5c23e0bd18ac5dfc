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
    if name == "shipped":
        return 15, 15, 1000, "flockers as shipped: 1000 birds, 100x100 field = 15x15 cells (flockers/src/main.rs:40-44)"
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
    """SM clock + clock-event reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is polled from
    a thread every ~1 ms (a 20-tick timed region lasts ~20 ms: `nvidia-smi -lms 100` saw two samples of it); nvidia-smi
    is the fallback when the NVML binding is missing."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: resolve through the PCI bus id of the CUDA device
            try:
                import torch
                bus = torch.cuda.get_device_properties(index).pci_bus_id  # torch >= 2.x
                dom = getattr(torch.cuda.get_device_properties(index), "pci_domain_id", 0)
                dev = torch.cuda.get_device_properties(index).pci_device_id
                h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:{dev:02x}.0".encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nvml, self._h = pynvml, h
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def start(self):
        if self._nvml is not None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self._nvml, self._h
        bits = [(nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.samples.append((mhz, [name for bit, name in bits if r & bit]))
            except Exception:
                pass
            time.sleep(0.001)

    def _read(self):
        for line in self.proc.stdout:
            p = [q.strip() for q in line.split(",")]
            try:
                self.samples.append((float(p[0]), [n for n, v in zip(self.NAMES, p[2:6]) if v.lower().startswith("active")]))
                self.sm_max = max(self.sm_max or 0.0, float(p[1]))
            except Exception:
                continue

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
            self._thread = None
            src = "nvml, 1 ms polling"
        elif self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            src = "nvidia-smi -lms 20"
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        sm = [m for m, _ in self.samples]
        reasons = sorted({r for _, rs in self.samples for r in rs})
        self.samples = []
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(sm), "source": src}


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

CPU_BIRDS_CAP = 32_000_000   # largest population the CPU arm steps per tick (128 M birds x 25 ticks would take > 2 min)


def cpu_sample_geometry(name: str, world: int, birds_per_gpu, cap: int):
    """The CPU arm runs the workload its `config` names whenever that fits the bound (16 M birds at N=1, 32 M at N=2);
    beyond it, the largest slice of the same workload that does: the configuration of fewer GPUs of the same weak-scaling
    series (same rows, same density, fewer column strips).  -> (gx, gy, n, whole: bool)"""
    gx, gy, n, _ = workload_geometry(name, world, birds_per_gpu)
    if n <= cap:
        return gx, gy, n, True
    w = world
    while w > 1:
        w //= 2
        gx2, gy2, n2, _ = workload_geometry(name, w, birds_per_gpu)
        if n2 <= cap and n2 < n:
            return gx2, gy2, n2, False
    # a single-GPU configuration above the bound: shrink both axes, keep the density
    scale = (cap / n) ** 0.5
    return max(int(gx * scale), 3), max(int(gy * scale), 3), cap, False


def cpu_oracle_run(gx: int, gy: int, n_birds: int, ticks: int, warm: int, clustered: bool, threads: int | None = None):
    """The reference algorithm (oracle port) on the host cores: `n_birds` on gx x gy cells, `warm` + `ticks` ticks."""
    threads = threads or (os.cpu_count() or 1)
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is entitled to all host threads, and libgomp
    # reads the variable when the oracle library is loaded (below)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    from oracle import oracle as O
    O.build()
    W, H = field_size(gx), field_size(gy)
    loc = make_birds(n_birds, W, H, clustered)
    b = np.zeros(n_birds, dtype=O.BIRD)
    b["id"] = np.arange(n_birds, dtype=np.uint32)
    b["x"], b["y"] = loc[:, 0], loc[:, 1]
    del loc
    sim = O.Sim(W, H)
    sim.init(b)
    for _ in range(warm):
        sim.step(seed=STEP_SEED, threads=threads)
    t0 = time.perf_counter()
    for _ in range(ticks):
        sim.step(seed=STEP_SEED, threads=threads)
    dt = time.perf_counter() - t0
    return n_birds * ticks / dt, dt, threads, (f"{n_birds} birds on {gx}x{gy} cells, {ticks} ticks (+{warm} warm-up), "
                                                f"OpenMP over agents, {threads} threads")


def common_config(args, world: int):
    """the `config` object both arms print: it names the workload and nothing that depends on the implementation"""
    gx, gy, n_total, label = workload_geometry(args.workload, world, args.birds_per_gpu)
    return {"workload": label, "birds_total": n_total, "birds_per_gpu": n_total // world,
            "cells": f"{gx}x{gy}", "data": "synthetic", "seeds": {"init": INIT_SEED, "step": STEP_SEED},
            "l2": "inputs larger than L2 (>=20 B/bird x birds per GPU vs 126 MB)" if n_total // world >= 8_000_000
                  else "working set fits L2 (reference config size)"}


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
    gx, gy, n, whole = cpu_sample_geometry(args.workload, world, args.birds_per_gpu, args.cpu_birds)
    value, dt, threads, sample = cpu_oracle_run(gx, gy, n, args.steps, args.warmup, args.workload == "clustered")
    sample = ("the whole workload: " if whole else "bounded sample of the workload: ") + sample
    line = {
        "impl": "reference", "metric": "bird-updates/sec", "value": value, "unit": "bird-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.workload == "strong16m" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": common_config(args, world),
        "sample_is_whole_workload": whole,
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
        from examples_b200 import distributed as fdist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_over_ranks(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x: float) -> float:
        return reduce_over_ranks(x, dist.ReduceOp.MAX if world > 1 else None)

    def sum_over_ranks(x: float) -> float:
        return reduce_over_ranks(x, dist.ReduceOp.SUM if world > 1 else None)

    def build_field(workload: str, birds_per_gpu):
        """this rank's field (or strip) of `workload` with its birds uploaded and published"""
        gx, gy, n_total, label = workload_geometry(workload, world, birds_per_gpu)
        W, H = field_size(gx), field_size(gy)
        clustered = workload == "clustered"
        n_local = n_total // world
        cap = int(n_local * 1.25) + 65536
        if world == 1:
            f = Field2D(W, H, D, True, capacity=cap, device=local_rank)
            loc = make_birds(n_local, W, H, clustered)
            id_base = 0
        else:
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
        h_id = torch.from_numpy(np.arange(id_base, id_base + n_local, dtype=np.uint32).view(np.int32)).pin_memory()
        h_loc = torch.from_numpy(loc).pin_memory()
        h_ld = torch.zeros((n_local, 2), dtype=torch.float32).pin_memory()
        f.set_objects_raw(n_local, h_id.data_ptr(), h_loc.data_ptr(), h_ld.data_ptr())
        f.set_math_mode(args.math)
        f.commit()  # lazy_update (+ first halo exchange on >1 GPU)
        f.synchronize()   # also the end of the pinned buffers' lifetime rule (include/flk.h)
        assert f.num_owned() == n_local, (f.num_owned(), n_local)
        return dict(f=f, W=W, H=H, gx=gx, gy=gy, n_total=n_total, n_local=n_local, cap=cap, loc=loc, id_base=id_base,
                    label=label, clustered=clustered)

    def timed_run(f, n_total, first, warmup, steps, sampler=None, profile=True, sampler_wanted=False):
        """W warm-up ticks, then K ticks timed on the device (CUDA events on the handle's stream, max over ranks).
        profile: CUDA events around every k_step launch of the timed region (plain launches instead of graph replays)"""
        f.run(first, warmup, STEP_SEED)
        f.synchronize()
        launches0 = f.launch_count()
        f.profile_enable(1 if profile else 0)
        if sampler is not None:
            sampler.start()
        barrier()
        f.run(first + warmup, steps, STEP_SEED)
        f.synchronize()
        barrier()
        ms = max_over_ranks(f.elapsed_ms())
        clocks = sampler.stop() if sampler is not None else None
        k_ms, k_n = f.profile_step_kernel()
        f.profile_enable(False)
        # a 20-tick region lasts ~23 ms and NVML sometimes answers slower than that: keep the SAME load running (untimed)
        # until the sampler has seen it, and say so (rank 0 decides for everybody: the ticks of a strip are collective)
        short = 1.0 if (clocks is not None and clocks.get("samples", 0) < 5) else 0.0
        if sampler_wanted and max_over_ranks(short) > 0.0:
            more = max(200, steps)
            if sampler is not None:
                sampler.start()
            f.run(first + warmup + steps, more, STEP_SEED)
            f.synchronize()
            barrier()
            if sampler is not None:
                c2 = sampler.stop()
                if c2.get("samples", 0) > clocks.get("samples", 0):
                    c2["source"] = c2.get("source", "") + f"; timed region too short for the sampler, continued over {more} more untimed ticks of the same load"
                    clocks = c2
        return dict(ms=ms, value=n_total * steps / (ms * 1e-3), k_ms=k_ms, k_n=k_n, clocks=clocks,
                    launches=f.launch_count() - launches0)

    # ---- the benchmarked workload ----
    wl = build_field(args.workload, args.birds_per_gpu)
    f, W, H, n_total, n_local, cap, loc, id_base = (wl[k] for k in ("f", "W", "H", "n_total", "n_local", "cap", "loc", "id_base"))
    clustered = wl["clustered"]
    main = timed_run(f, n_total, 0, args.warmup, args.steps, ClockSampler(local_rank) if rank == 0 else None, sampler_wanted=True)
    ms, value, clocks, launches = main["ms"], main["value"], main["clocks"], main["launches"]
    # the other kernels of a tick, timed over a few more ticks outside the timed region
    f.profile_enable(2)
    f.run(args.warmup + args.steps, min(args.steps, 20), STEP_SEED)
    f.synchronize()
    rebin_ms = f.profile_rebin()
    f.profile_enable(False)
    n_now = sum_over_ranks(float(f.num_owned()))

    # roofline of the dominant kernel (k_step): algorithmic bytes per launch / average launch duration
    k_avg_ms = max_over_ranks(main["k_ms"] / max(main["k_n"], 1))
    peak, peak_src = measured_peak_gbs()
    achieved = ALGO_BYTES_PER_UPDATE * n_local / (k_avg_ms * 1e-3) / 1e9
    traffic = ncu_traffic_per_launch(n_total // world)
    warp_instr = ncu_traffic_per_launch(n_total // world, "warp_instructions_per_launch")
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count

    # ---- e2e: per step, host -> device copy of every agent, one tick, device -> host read of the new state ----
    e2e_steps = max(1, min(args.e2e_steps, args.steps))
    reps = max(1, args.e2e_reps)
    if world == 1:
        # The host keeps the agents as a list indexed by id (flockers/src/model/state.rs:40-51), so the dense C-ABI
        # calls move Bird minus its id: 16 B per bird each way.  Every step uploads all agents from pinned host memory,
        # publishes them (lazy_update), runs Bird::step + lazy_update and downloads all agents; the download is the next
        # step's upload (a dependent chain).  The headline is ONE repetition, what the reference ships
        # (simulate_old!(state, step, 1), flockers/src/main.rs:47); `pipelined_value` runs `reps` independent
        # repetitions (simulate!'s third argument) over one handle each, so that one repetition's download overlaps the
        # next one's upload on the two PCIe directions.
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

        e2e_value = e2e_run([f], e2e_steps)
        extra = [Field2D(W, H, D, True, capacity=cap, device=local_rank) for _ in range(reps - 1)]
        for fx in extra:
            fx.set_math_mode(args.math)
        pipelined = e2e_run([f] + extra, e2e_steps) if reps > 1 else e2e_value
        for fx in extra:
            fx.close()
        e2e_bytes = 16 * n_local
        e2e_what = ("ONE repetition (flockers/src/main.rs:47): per step flk_set_objects_dense(pinned host, 16 B/bird) + "
                    "lazy_update + Bird::step + lazy_update + flk_snapshot_dense_async(pinned host, 16 B/bird); the output "
                    f"feeds the next step's input; pipelined_value = {reps} independent repetitions over {reps} handles")
    else:
        # One strip per rank: the birds a rank owns change with migration, so ids travel with the records (20 B per
        # bird each way).  Headline: one repetition; pipelined_value: `reps` independent repetitions, each its own group
        # of strip handles (own NCCL communicator); every rank walks the repetitions in the same order (collectives).
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
        e2e_value = e2e_run_dist([f], e2e_steps)
        pipelined = e2e_run_dist([f] + extra, e2e_steps) if reps > 1 else e2e_value
        for fx in extra:
            fx.close()
        e2e_bytes = 20 * (n_total // world)
        e2e_what = ("ONE repetition: per rank and step flk_set_objects(pinned host, 20 B/bird) + commit + Bird::step + "
                    "exchange + lazy_update + flk_dist_snapshot_owned_async(pinned host, 20 B/bird); the output feeds the "
                    f"next step's input; bytes are per rank; pipelined_value = {reps} independent repetitions over {reps} "
                    "groups of strip handles")
    f.close()

    # ---- parity leg: every bench line carries a correctness check of the code that was just timed ----
    parity = parity_check(args, world, rank, local_rank) if not args.no_verify else None

    # ---- the other BASELINE configs, short legs (driver-visible numbers for configs 2, 3 and 5) ----
    extras = {}
    if not args.no_extras and args.workload == "weak16m" and args.birds_per_gpu is None:
        legs = [("shipped", 2000), ("1m", 200), ("clustered", 10)] if world == 1 else [("strong16m", 100), ("clustered", 10)]
        for name, k in legs:
            try:
                w2 = build_field(name, None)
                r2 = timed_run(w2["f"], w2["n_total"], 0, 10, k, profile=False)     # the way flk_run runs: graph replays
                r3 = timed_run(w2["f"], w2["n_total"], 10 + k, 0, min(k, 20))       # a few more ticks with k_step events
                extras[name] = {"value": r2["value"], "unit": "bird-updates/s", "ms_per_step": r2["ms"] / k, "steps": k,
                                "warmup": 10, "workload": w2["label"], "birds_after": int(sum_over_ranks(float(w2["f"].num_owned()))),
                                "k_step_ms": max_over_ranks(r3["k_ms"] / max(r3["k_n"], 1))}
                w2["f"].close()
            except Exception as e:   # an extra leg never takes the headline down with it
                extras[name] = {"error": repr(e)}

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle on the host cores, bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gx, gy, n, whole = cpu_sample_geometry(args.workload, 1, args.birds_per_gpu, args.cpu_birds)
        v, dt, threads, sample = cpu_oracle_run(gx, gy, n, args.cpu_ticks, 1, clustered)
        cpu = {"value": v, "unit": "bird-updates/s", "cores": threads, "kind": "port",
               "sample": ("the whole workload: " if whole else "bounded sample of the workload: ") + sample}

    if rank == 0:
        line = {
            "metric": "bird-updates/sec", "value": value, "unit": "bird-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.workload == "strong16m" else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": common_config(args, world),
            "details": {"decomposition": "single GPU" if world == 1 else
                                         (f"x-strips over {world} GPUs, NCCL halo+migration" if args.tile_rows == 1 else
                                          f"{world // args.tile_rows} x {args.tile_rows} tiles over {world} GPUs, NCCL halo+migration in two hops"),
                        "math": "exact (IEEE f32, bit-identical to the oracle)" if args.math == 0 else "fast",
                        "birds_after": int(n_now)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_step", "kernel_ms": k_avg_ms,
                         "kernel_share_of_step": k_avg_ms / (ms / args.steps), "peak_source": peak_src,
                         "other_kernels_ms": {"k_scan": rebin_ms[0], "k_scatter": rebin_ms[1], "k_rank": rebin_ms[2]},
                         "algorithmic_bytes_per_update": ALGO_BYTES_PER_UPDATE,
                         "issue_slots": issue_roofline(warp_instr, k_avg_ms, sm_count, (clocks or {}).get("sm_mhz")),
                         "note": "path is FP32-issue bound, not HBM bound (DESIGN.md roofline): see issue_slots"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "bird-updates/s", "h2d_bytes_per_step": e2e_bytes,
                    "d2h_bytes_per_step": e2e_bytes, "steps": e2e_steps, "reps": 1, "pipelined_value": pipelined,
                    "pipelined_reps": reps, "what": e2e_what},
            "parity_check": parity,
            "extra_configs": extras,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def parity_check(args, world: int, rank: int, local_rank: int):
    """Correctness of the code that was just timed, inside the bench line.
    N = 1: the CUDA path against the CPU oracle, 200 000 birds, 3 ticks, whole population bit for bit.
    N > 1: 1 M birds (BASELINE config 2 geometry) as `world` strips over NCCL against the single field on rank 0's GPU
           AND against the oracle, 4 ticks, bit for bit; the digest is the sha256 of the id-sorted final state."""
    import hashlib
    import torch
    import torch.distributed as dist
    from examples_b200 import Field2D
    try:
        if world == 1:
            cells, n, ticks = 212, 200_000, 3
        else:
            cells, n, ticks = 480, 1_000_000, 4
        Wp = field_size(cells)
        loc = make_birds(n, Wp, Wp, False, seed=INIT_SEED + 77)
        ids = np.arange(n, dtype=np.uint32)
        ld = np.zeros_like(loc)
        got = None
        if world > 1:
            from examples_b200 import distributed as fdist
            fd = fdist.DistField2D(Wp, Wp, D, capacity=n // world * 2 + 65536, device=local_rank, rank=rank, world=world,
                                   tile_rows=args.tile_rows)
            c0, c1, r0, r1 = fd.tile()
            col, row = fdist.column_of(loc[:, 0], D), fdist.column_of(loc[:, 1], D)
            mine = (col >= c0) & (col < c1) & (row >= r0) & (row < r1)
            fd.set_objects(ids[mine], loc[mine], ld[mine])
            fd.set_math_mode(args.math)
            fd.commit()
            fd.run(0, ticks, STEP_SEED)
            o_id, o_loc, o_ld = fd.snapshot_owned()
            fd.close()
            k = torch.tensor([len(o_id)], dtype=torch.int64, device="cuda")
            ks = [torch.zeros_like(k) for _ in range(world)]
            dist.all_gather(ks, k)
            kmax = int(max(int(t.item()) for t in ks))
            pack = np.zeros((kmax, 5), dtype=np.float32)
            pack[: len(o_id), 0] = o_id.view(np.float32)
            pack[: len(o_id), 1:3] = o_loc
            pack[: len(o_id), 3:5] = o_ld
            mine_t = torch.from_numpy(pack).cuda()
            parts = [torch.zeros_like(mine_t) for _ in range(world)]
            dist.all_gather(parts, mine_t)
            if rank == 0:
                rows = np.concatenate([p.cpu().numpy()[: int(kk.item())] for p, kk in zip(parts, ks)])
                o = np.argsort(rows[:, 0].view(np.uint32), kind="stable")
                got = np.ascontiguousarray(rows[o])
        if rank != 0:
            return None
        f1 = Field2D(Wp, Wp, D, True, capacity=n, device=local_rank)
        f1.set_objects(ids, loc, ld)
        f1.set_math_mode(args.math)
        f1.lazy_update()
        f1.run(0, ticks, STEP_SEED)
        s_id, s_loc, s_ld = f1.snapshot_by_id()
        f1.close()
        single = np.zeros((n, 5), dtype=np.float32)
        single[:, 0] = s_id.view(np.float32)
        single[:, 1:3], single[:, 3:5] = s_loc, s_ld
        from oracle import oracle as O
        b = np.zeros(n, dtype=O.BIRD)
        b["id"], b["x"], b["y"] = ids, loc[:, 0], loc[:, 1]
        sim = O.Sim(Wp, Wp)
        sim.init(b)
        for _ in range(ticks):
            sim.step(seed=STEP_SEED, threads=os.cpu_count() or 1)
        a = sim.agents()
        orc = np.zeros((n, 5), dtype=np.float32)
        orc[:, 0] = a["id"].view(np.float32)
        orc[:, 1], orc[:, 2], orc[:, 3], orc[:, 4] = a["x"], a["y"], a["ldx"], a["ldy"]
        same = lambda u, v: u.shape == v.shape and np.array_equal(u.view(np.uint32), v.view(np.uint32))
        res = {"birds": n, "ticks": ticks, "cells": f"{cells}x{cells}",
               "digest": hashlib.sha256(single.view(np.uint8).tobytes()).hexdigest()[:16],
               "single_gpu_vs_oracle": "bit-exact" if (args.math == 0 and same(single, orc)) else
                                       ("fast math: max abs err %.3g" % float(np.abs(single[:, 1:] - orc[:, 1:]).max())
                                        if args.math else "MISMATCH")}
        ok = res["single_gpu_vs_oracle"] != "MISMATCH"
        if world > 1:
            res["strips_over_nccl_vs_single_gpu"] = "bit-exact" if same(got, single) else "MISMATCH"
            res["ranks"] = world
            ok = ok and res["strips_over_nccl_vs_single_gpu"] == "bit-exact"
        res["status"] = "ok" if ok else "FAILED"
        return res
    except Exception as e:
        return {"status": "FAILED", "error": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="weak16m", choices=["weak16m", "strong16m", "1m", "clustered", "shipped"])
    ap.add_argument("--birds-per-gpu", type=int, default=None)
    ap.add_argument("--math", type=int, default=0)
    ap.add_argument("--tile-rows", type=int, default=1)   # > 1: split the rows too (tiles); default = x-strips
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-reps", type=int, default=3)           # independent repetitions pipelined in the e2e leg
    ap.add_argument("--cpu-birds", type=int, default=CPU_BIRDS_CAP)   # largest population the CPU arm steps per tick
    ap.add_argument("--cpu-ticks", type=int, default=12)         # cpu_baseline leg of our arm: ~10 s on 16 cores at 16 M birds
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")          # skip the parity leg
    ap.add_argument("--no-extras", action="store_true")          # skip the short legs on the other BASELINE configs
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
