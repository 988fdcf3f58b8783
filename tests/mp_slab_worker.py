"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), NCCL transport of the slab exchange.
Rank 0 also runs the single-GPU reference run and compares.  Exit code 0 = parity."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t2d = importlib.import_module("2dtissue_b200")
    chart = t2d.load_chart(os.path.join(ROOT, "tests", "golden", "ellipsoid_x4.t2dchart"))
    N, steps = 40000, 5
    prec = t2d.PRECISION_FP64
    uv, n = t2d.seed_particles(N, seed=41)
    sigma = float(np.sqrt(0.5 * 451.3 / (np.pi * N)))
    kw = dict(v0=0.1, k=1.0, sigma=sigma, step_size=0.001, eta=0.03, seed=5, neigh_mode=t2d.NEIGH_EUCLID, precision=prec)
    # every rank projects the same seeded particles (deterministic), so the cuts agree without communication
    c0 = t2d.Context(chart, capacity=N, device=local, **kw)
    c0.set_particles(uv, n)
    s0 = c0.download(("uv", "n", "vid", "r3d"))
    ref = None
    if rank == 0:
        c0.step(steps)
        ref = c0.download()
    c0.close()
    cuts = t2d.slab_cuts(s0["r3d"][:N], world)
    uid = [t2d.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx = t2d.Context(chart, capacity=N, device=local, **kw)
    ctx.comm_init(rank, world, uid[0], cuts)
    p = t2d.partition_by_slab(s0, cuts, rank)
    ctx.set_state(p["uv"], p["n"], p["vid"], p["r3d"], ids=p["ids"])
    fault = ctx.step(steps)
    part = ctx.download()
    part["ids"] = ctx.download_ids()
    parts = [None] * world
    dist.all_gather_object(parts, part)
    ok = True
    if rank == 0:
        out = t2d.merge_by_id(parts, N)
        ok = fault == 0 and sum(q["ids"].size for q in parts) == N
        for k in ("n", "vid", "face", "color", "uv", "rdot", "r3d"):
            if not np.array_equal(out[k], ref[k]):
                print("MISMATCH", k, int(np.sum(out[k] != ref[k])))
                ok = False
        print("nccl slabs world=%d: owned %s, parity %s" % (world, [q["ids"].size for q in parts], ok))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
