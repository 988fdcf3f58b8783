"""Host-side logic of the multi-GPU slabs on CPU, world_size 2 over gloo: equal-count cuts, ownership, partition
by slab with global ids, gather + merge by id.  (The device side is covered by tests/test_gpu_slabs.py on one GPU and
tests/test_gpu_multi.py on several.)"""
import importlib
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    host = importlib.import_module("2dtissue_b200.host")
    N = 5000
    rng = np.random.default_rng(7)            # same seed on every rank, like the seeded particles of bench.py
    state = dict(uv=rng.random(2 * N), n=rng.integers(0, 360, N).astype(np.int32), vid=rng.integers(0, 4725, N).astype(np.int32),
                 r3d=rng.normal(size=3 * N))
    cuts = host.slab_cuts(state["r3d"][:N], world)
    part = host.partition_by_slab(state, cuts, rank)
    # what a rank would download after stepping: its owned particles (here unchanged) + their global ids
    mine = dict(uv=part["uv"], n=part["n"], vid=part["vid"], r3d=part["r3d"], ids=part["ids"])
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    uid = [b"x" * 128 if rank == 0 else None]   # the NCCL unique id travels the same way
    dist.broadcast_object_list(uid, src=0)
    merged = host.merge_by_id(parts, N)
    ok = all(np.array_equal(merged[k], state[k]) for k in ("uv", "n", "vid", "r3d"))
    ok = ok and uid[0] == b"x" * 128 and abs(part["n"].size - N / world) <= 1
    ok = ok and np.all(host.slab_of(part["r3d"][:part["n"].size], cuts) == rank)
    q.put((rank, bool(ok), int(part["n"].size)))
    dist.destroy_process_group()


def test_partition_gather_merge_world2():
    world, port = 2, 29541
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert sum(cnt for _, _, cnt in res) == 5000


def test_slab_cuts_and_ownership():
    sys.path.insert(0, ROOT)
    host = importlib.import_module("2dtissue_b200.host")
    x = np.random.default_rng(1).normal(size=10001)
    for world in (1, 2, 4, 8):
        cuts = host.slab_cuts(x, world)
        assert len(cuts) == world - 1 and np.all(np.diff(cuts) > 0)
        owner = host.slab_of(x, cuts)
        cnt = np.bincount(owner, minlength=world)
        assert cnt.max() - cnt.min() <= 1
    assert list(host.slab_of([0.0, 1.0, 2.0], [1.0])) == [0, 1, 1]     # slab r owns [cuts[r-1], cuts[r])
