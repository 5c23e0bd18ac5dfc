"""Run under torchrun (one rank per GPU): strips over NCCL must reproduce the golden fixture / a single-GPU run.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/nccl_worker.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from examples_b200 import Field2D
    from examples_b200.distributed import DistField2D
    from tests.helpers import D, from_arrays, to_arrays, bits_equal
    from oracle import oracle as O

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True

    def gather(f):
        ids, loc, ld = f.snapshot_owned()
        mine = from_arrays(ids, loc, ld)
        parts = [None] * world
        dist.all_gather_object(parts, mine.view(np.uint8).tobytes())
        allb = np.concatenate([np.frombuffer(p, dtype=O.BIRD) for p in parts])
        return np.sort(allb, order="id")

    # 1. golden fixture (BASELINE config 1): every pinned tick bit-exact
    g = np.load(os.path.join(ROOT, "tests", "golden", "flockers_1000.npz"))
    birds = g["init"].view(O.BIRD).reshape(-1)
    f = DistField2D(100.0, 100.0, D, capacity=1000, device=local, rank=rank, world=world)
    f.set_objects(*to_arrays(birds))
    f.commit()
    t = 0
    for stop in list(g["ticks"]):
        f.run(t, stop - t, int(g["seed"]))
        t = stop
        got = gather(f)
        if rank == 0:
            same = bits_equal(got, g[f"state_{stop}"].view(O.BIRD).reshape(-1))
            print(f"[nccl x{world}] golden tick {stop}: {'ok' if same else 'MISMATCH'}", flush=True)
            ok &= same
    f.close()

    # 2. 200k birds, injected draws, against a single-GPU field on rank 0
    n, W = 200_000, 1400.0
    birds = O.make_uniform(0xF10C, W, W, n)
    f = DistField2D(W, W, D, capacity=n, device=local, rank=rank, world=world)
    f.set_objects(*to_arrays(birds))
    f.commit()
    s = None
    if rank == 0:
        s = Field2D(W, W, D, True, capacity=n, device=local)
        s.set_objects(*to_arrays(birds))
        s.lazy_update()
    rng = np.random.default_rng(4)
    for t in range(8):
        r = rng.random((n, 2), dtype=np.float32)
        f.step(t, injected=r)
        got = gather(f)
        if rank == 0:
            s.step(t, injected=r)
            s.lazy_update()
            same = bits_equal(got, from_arrays(*s.snapshot_by_id()))
            ok &= same
            if not same:
                print(f"[nccl x{world}] 200k tick {t}: MISMATCH", flush=True)
    if rank == 0:
        print(f"[nccl x{world}] 200k birds, 8 ticks vs single GPU: {'ok' if ok else 'MISMATCH'}", flush=True)
    owned = torch.tensor([f.num_owned()], device="cuda")
    dist.all_reduce(owned)
    if rank == 0:
        print(f"[nccl x{world}] owned total {int(owned.item())}", flush=True)
        ok &= int(owned.item()) == n
    f.close()
    # 2b. tiles (columns and rows split, two-hop diagonal exchange) when the ranks form a 2-row grid
    if world >= 4 and world % 2 == 0:
        ft = DistField2D(W, W, D, capacity=n, device=local, rank=rank, world=world, tile_rows=2)
        ft.set_objects(*to_arrays(birds))
        ft.commit()
        st = None
        if rank == 0:
            st = Field2D(W, W, D, True, capacity=n, device=local)
            st.set_objects(*to_arrays(birds))
            st.lazy_update()
        same = True
        for t0 in (0, 10, 20):
            ft.run(t0, 10)
            got = gather(ft)
            if rank == 0:
                st.run(t0, 10)
                same &= bits_equal(got, from_arrays(*st.snapshot_by_id()))
        if rank == 0:
            print(f"[nccl x{world}] tiles {world // 2} x 2, 200k birds, 30 ticks vs single GPU: {'ok' if same else 'MISMATCH'}", flush=True)
            ok &= same
        ft.close()
    # 3. the flockers_mpi model as shipped (flockers_mpi/src/main.rs:38-44: 1000 birds, 100x100, 100 steps) through the
    #    State-shaped mirror, against the single-GPU Flocker with the same placement
    from examples_b200 import Flocker, Schedule
    from examples_b200.distributed import DistFlocker
    st = DistFlocker((100.0, 100.0), 1000, device=local, rank=rank, world=world)
    st.init()
    for _ in range(100):
        st.tick()
    got = gather(st.field1)
    if rank == 0:
        ref = Flocker((100.0, 100.0), 1000, device=local)
        sch = Schedule()
        ref.init(sch)
        for _ in range(100):
            sch.step(ref)
        same = bits_equal(got, from_arrays(*ref.field1.snapshot_by_id()))
        print(f"[nccl x{world}] flockers_mpi shipped config, 100 steps vs flockers: {'ok' if same else 'MISMATCH'}", flush=True)
        ok &= same
    st.field1.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    if rank == 0:
        print("NCCL_WORKER_OK" if ok else "NCCL_WORKER_FAIL", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
