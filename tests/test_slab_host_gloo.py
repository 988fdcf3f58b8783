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


def test_routing_rule_is_complete():
    """The slab exchange's routing rule (host restatement of slab_classify / k_comm_unpack_far): after ONE exchange every
    particle has exactly one owner, and every rank holds a copy of every foreign particle within `halo` of its slab —
    also when particles jump any number of slabs in one step (seam re-entry)."""
    sys.path.insert(0, ROOT)
    host = importlib.import_module("2dtissue_b200.host")
    rng = np.random.default_rng(11)
    for world in (2, 3, 4, 8):
        N, halo = 40000, 0.01
        x_old = rng.random(N) * world                       # slabs of width ~1 >= 4 halo
        cuts = host.slab_cuts(x_old, world)
        owner_old = host.slab_of(x_old, cuts)
        x_new = x_old + rng.normal(scale=0.004, size=N)     # ordinary motion: a fraction of the halo width
        jump = rng.random(N) < 0.02                         # 2 %: land anywhere (seam re-entry, very fast particles)
        x_new[jump] = rng.random(int(jump.sum())) * world
        near = rng.random(N) < 0.02                         # 2 %: land right next to some cut
        x_new[near] = rng.choice(cuts, int(near.sum())) + rng.normal(scale=0.3 * halo, size=int(near.sum()))
        owned = [set() for _ in range(world)]
        halo_copies = [set() for _ in range(world)]
        far_pool = []
        for r in range(world):
            ids = np.nonzero(owner_old == r)[0]
            rt = host.route_after_step(x_new[ids], cuts, r, halo)
            assert np.all(rt["stay"] ^ rt["to_left"] ^ rt["to_right"] ^ rt["far"])          # exactly one route
            owned[r].update(ids[rt["stay"]])
            halo_copies[r].update(ids[rt["keep_halo"]])
            if r > 0:
                owned[r - 1].update(ids[rt["to_left"]])
                halo_copies[r - 1].update(ids[rt["halo_left"]])
            else:
                assert not rt["to_left"].any()
            if r < world - 1:
                owned[r + 1].update(ids[rt["to_right"]])
                halo_copies[r + 1].update(ids[rt["halo_right"]])
            else:
                assert not rt["to_right"].any()
            far_pool.append((ids[rt["far"]], rt["dest"][rt["far"]], r))
        nfar = 0
        for ids, dest, src in far_pool:
            nfar += ids.size
            for r in range(world):
                if r == src:
                    assert not np.any(dest == r)             # a far record never comes back to its sender
                    continue
                adopt, copy = host.far_receive(x_new[ids], dest, cuts, r, halo)
                owned[r].update(ids[adopt])
                halo_copies[r].update(ids[copy])
        owner_new = host.slab_of(x_new, cuts)
        assert sum(len(o) for o in owned) == N
        for r in range(world):
            assert owned[r] == set(np.nonzero(owner_new == r)[0])                              # one owner, the right one
            lo = -np.inf if r == 0 else cuts[r - 1]
            hi = np.inf if r == world - 1 else cuts[r]
            needed = set(np.nonzero((owner_new != r) & (x_new >= lo - halo) & (x_new < hi + halo))[0])
            assert needed <= halo_copies[r], (world, r, len(needed - halo_copies[r]))          # no neighbour is missed
        assert world == 2 or nfar > 0
