"""Multi-GPU row on ONE GPU: P logical slabs with the loopback transport (t2d_comm_init_local /
t2d_step_local) must reproduce the single-context run bit for bit — neighbour sets, faces, headings, and in fp64
also every float — because global ids travel with the particles (summation order and noise are partition
independent).  The NCCL transport shares every kernel with this path; tests/test_gpu_multi.py runs it on 2 GPUs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run_single(t2d, chart, uv, n, sigma, steps, precision, eta=0.0):
    N = n.size
    ctx = t2d.Context(chart, v0=0.1, k=1.0, sigma=sigma, step_size=0.001, eta=eta, seed=99, neigh_mode=t2d.NEIGH_EUCLID,
                      precision=precision, capacity=N)
    ctx.set_particles(uv, n)
    s0 = ctx.download(("uv", "n", "vid", "r3d"))
    fault = ctx.step(steps)
    out = ctx.download()
    ctx.close()
    return s0, out, fault


@pytest.mark.parametrize("world", [2, 3])
def test_local_slabs_equal_single_context_fp64(t2d, chart, world):
    N, steps = 20000, 4
    uv, n = t2d.seed_particles(N, seed=31)
    sigma = float(np.sqrt(0.5 * 451.3 / (np.pi * N)))
    s0, ref, fault0 = _run_single(t2d, chart, uv, n, sigma, steps, t2d.PRECISION_FP64, eta=0.05)
    cuts = t2d.slab_cuts(s0["r3d"][:N], world)
    grp = t2d.LocalSlabGroup(chart, world, cuts, v0=0.1, k=1.0, sigma=sigma, step_size=0.001, eta=0.05, seed=99,
                             neigh_mode=t2d.NEIGH_EUCLID, precision=t2d.PRECISION_FP64, capacity=N)
    grp.set_state(s0)
    fault = grp.step(steps)
    out, owned = grp.download()
    assert fault == fault0 == 0
    assert sum(owned) == N and min(owned) > 0.2 * N / world
    for k in ("n", "vid", "face", "color"):
        assert np.array_equal(out[k], ref[k]), k                     # integer state: bit-exact
    for k in ("uv", "rdot", "r3d"):
        assert np.array_equal(out[k], ref[k]), k                     # fp64 floats: bit-exact too
    # particles did change owner during the run, and halo copies were exchanged
    moved = t2d.slab_of(out["r3d"][:N], cuts) != t2d.slab_of(s0["r3d"][:N], cuts)
    print("world %d: owned %s, %d particles changed slab" % (world, owned, int(moved.sum())))
    c = grp.ctxs[0].counters()
    assert c["cell_fallbacks"] == 0
    grp.close()


def test_local_slabs_fp32_fast_path(t2d, chart):
    N, steps, world = 50000, 3, 4
    uv, n = t2d.seed_particles(N, seed=32)
    sigma = float(np.sqrt(0.5 * 451.3 / (np.pi * N)))
    s0, ref, _ = _run_single(t2d, chart, uv, n, sigma, steps, t2d.PRECISION_FP32)
    cuts = t2d.slab_cuts(s0["r3d"][:N], world)
    grp = t2d.LocalSlabGroup(chart, world, cuts, v0=0.1, k=1.0, sigma=sigma, step_size=0.001, seed=99,
                             neigh_mode=t2d.NEIGH_EUCLID, precision=t2d.PRECISION_FP32, capacity=N)
    grp.set_state(s0)
    assert grp.step(steps) == 0
    out, owned = grp.download()
    assert sum(owned) == N
    assert np.array_equal(out["color"], ref["color"])                # neighbour sets identical
    # fp32 sums run in visiting order, which depends on the partition: compare within the fast-path tolerance
    same = out["n"] == ref["n"]
    assert same.mean() > 0.999
    s2 = np.concatenate([same, same])
    assert np.max(np.abs(out["uv"][s2] - ref["uv"][s2])) < 1e-5
    assert (out["face"][same] == ref["face"][same]).mean() > 0.999
    grp.close()


def test_slab_observables_and_counts(t2d, chart):
    N, world = 8000, 2
    uv, n = t2d.seed_particles(N, seed=33)
    sigma = float(np.sqrt(0.5 * 451.3 / (np.pi * N)))
    s0, ref, _ = _run_single(t2d, chart, uv, n, sigma, 2, t2d.PRECISION_FP64)
    cuts = t2d.slab_cuts(s0["r3d"][:N], world)
    grp = t2d.LocalSlabGroup(chart, world, cuts, v0=0.1, k=1.0, sigma=sigma, step_size=0.001, seed=99,
                             neigh_mode=t2d.NEIGH_EUCLID, precision=t2d.PRECISION_FP64, capacity=N)
    grp.set_state(s0)
    assert grp.step(2) == 0
    obs = [c.observables() for c in grp.ctxs]
    assert sum(o["count"] for o in obs) == N                          # halo copies are not counted
    sc, ss = sum(o["sum_cos"] for o in obs), sum(o["sum_sin"] for o in obs)
    phi = np.hypot(sc, ss) / N
    c1 = t2d.Context(chart, sigma=sigma, neigh_mode=t2d.NEIGH_EUCLID, capacity=N)
    c1.set_state(ref["uv"], ref["n"], ref["vid"], ref["r3d"])
    assert abs(phi - np.hypot(*[np.sum(f(np.deg2rad(ref["n"].astype(np.float64)))) for f in (np.cos, np.sin)]) / N) < 1e-12
    grp.close()


def test_far_migration_is_exact(t2d, chart):
    """Seam re-entry (and the rare very fast particle) is not continuous in 3-D under the reference's lift: a particle
    can land several slabs away in ONE step.  Such particles travel through the far channel (every rank receives every
    rank's far migrants); with 8 slabs the run must still equal the single-context run bit for bit."""
    N, steps, world = 8000, 4, 8
    uv, n = t2d.seed_particles(N, seed=1234)
    sigma = float(np.sqrt(0.5 * 451.3 / (np.pi * N)))
    # v0 = 700: 0.7 chart widths per step — every particle re-enters through the seam and lands somewhere else
    kw = dict(v0=700.0, k=1.0, sigma=sigma, step_size=0.001, eta=0.02, seed=7, neigh_mode=t2d.NEIGH_EUCLID,
              precision=t2d.PRECISION_FP64)
    c0 = t2d.Context(chart, capacity=N, **kw)
    c0.set_particles(uv, n)
    s0 = c0.download(("uv", "n", "vid", "r3d"))
    cuts = t2d.slab_cuts(s0["r3d"][:N], world)
    grp = t2d.LocalSlabGroup(chart, world, cuts, capacity=N, **kw)
    grp.set_state(s0)
    prev, far = s0, 0
    for s in range(steps):
        assert c0.step(1) == 0
        ref = c0.download()
        assert grp.step(1) == 0, "no migration / overflow fault"
        out, owned = grp.download()
        assert sum(owned) == N
        for k in ("n", "vid", "face", "color", "uv", "rdot", "r3d"):
            assert np.array_equal(out[k], ref[k]), (s, k)
        far += int(np.sum(np.abs(t2d.slab_of(ref["r3d"][:N], cuts) - t2d.slab_of(prev["r3d"][:N], cuts)) >= 2))
        prev = ref
    print("particles that jumped >= 2 slabs in one step: %d" % far)
    assert far > 0, "the configuration no longer exercises the far channel"
    grp.close()
    c0.close()
