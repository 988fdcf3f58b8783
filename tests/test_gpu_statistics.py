"""Long runs are chaotic, so the fp32 fast path is compared with the reference semantics STATISTICALLY
(BASELINE.json north_star): polar order parameter and mean speed, mean over seeds within 3 standard errors.
The fp64 path needs no such test — it is bit-identical to the oracle step by step (test_gpu_parity.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _series_gpu(t2d, chart, uv, n, sigma, eta, seed, steps, precision):
    N = n.size
    ctx = t2d.Context(chart, v0=0.1, k=1.0, sigma=sigma, step_size=0.001, eta=eta, seed=seed, neigh_mode=t2d.NEIGH_EUCLID,
                      precision=precision, capacity=N)
    ctx.set_particles(uv, n)
    phi, spd = [], []
    for _ in range(steps):
        assert ctx.step(1) == 0
        o = ctx.observables()
        phi.append(o["phi"])
        spd.append(o["mean_speed"])
    ctx.close()
    return np.array(phi), np.array(spd)


def _series_oracle(oracle, t2d, chart, uv, n, sigma, eta, seed, steps):
    r3d, vid, _ = oracle.get_r3d(uv)
    st = dict(uv=uv, n=n, vid=vid, r3d=r3d)
    phi, spd = [], []
    for s in range(steps):
        st = oracle.step(st["uv"], st["n"], st["vid"], st["r3d"], 0.1, 1.0, sigma, 0.001, eta=eta, seed=seed, mode=1,
                         step_index=s, threads=0)
        assert st["fault"] == 0
        p, v = oracle.observables(st["n"], st["rdot"])
        phi.append(p)
        spd.append(v)
    return np.array(phi), np.array(spd)


def test_fp32_long_run_statistics_match_reference_semantics(t2d, chart, oracle):
    N, steps, eta, nseeds = 3000, 150, 0.1, 6
    sigma = float(np.sqrt(0.5 * 451.3 / (np.pi * N)))
    tail = slice(steps // 2, steps)
    g_phi, g_spd, o_phi, o_spd = [], [], [], []
    for s in range(nseeds):
        uv, n = t2d.seed_particles(N, seed=500 + s)
        a, b = _series_gpu(t2d, chart, uv, n, sigma, eta, 1000 + s, steps, t2d.PRECISION_FP32)
        c, d = _series_oracle(oracle, t2d, chart, uv, n, sigma, eta, 1000 + s, steps)
        # the first step is still deterministic to fp32 accuracy
        assert abs(a[0] - c[0]) < 5e-3 and abs(b[0] - d[0]) < 1e-3 * max(1.0, d[0])
        g_phi.append(a[tail].mean()); g_spd.append(b[tail].mean())
        o_phi.append(c[tail].mean()); o_spd.append(d[tail].mean())
    for name, g, o in (("phi", np.array(g_phi), np.array(o_phi)), ("mean speed", np.array(g_spd), np.array(o_spd))):
        se = np.sqrt(g.var(ddof=1) / nseeds + o.var(ddof=1) / nseeds)
        print("%s: fp32 %.5f +- %.5f | oracle %.5f +- %.5f | diff %.2e, 3 s.e. %.2e" %
              (name, g.mean(), g.std(ddof=1) / np.sqrt(nseeds), o.mean(), o.std(ddof=1) / np.sqrt(nseeds),
               abs(g.mean() - o.mean()), 3 * se))
        assert abs(g.mean() - o.mean()) <= 3 * se + 1e-4 * max(1.0, abs(o.mean()))


def _stat_compare(name, g, o, nseeds):
    g, o = np.array(g), np.array(o)
    se = np.sqrt(g.var(ddof=1) / nseeds + o.var(ddof=1) / nseeds)
    print("%s: GPU %.5f +- %.5f | oracle %.5f +- %.5f | diff %.2e, 3 s.e. %.2e" %
          (name, g.mean(), g.std(ddof=1) / np.sqrt(nseeds), o.mean(), o.std(ddof=1) / np.sqrt(nseeds), abs(g.mean() - o.mean()), 3 * se))
    assert abs(g.mean() - o.mean()) <= 3 * se + 1e-4 * max(1.0, abs(o.mean()))


def test_barycentric_lift_statistics(t2d, chart, oracle_mod):
    """Row f4: the barycentric-lift extension (T2D_LIFT_BARYCENTRIC) on the fp32 fast path against the oracle's implementation
    of the same option in fp64 — long runs compared on the polar order parameter and the mean speed, like the stock lift."""
    N, steps, eta, nseeds = 3000, 120, 0.1, 5
    sigma = float(np.sqrt(0.5 * 451.3 / (np.pi * N)))
    tail = slice(steps // 2, steps)
    orc = oracle_mod.Oracle(chart)
    orc.set_lift_mode(1)
    g_phi, g_spd, o_phi, o_spd = [], [], [], []
    for s in range(nseeds):
        uv, n = t2d.seed_particles(N, seed=700 + s)
        ctx = t2d.Context(chart, v0=0.1, k=1.0, sigma=sigma, step_size=0.001, eta=eta, seed=2000 + s, neigh_mode=t2d.NEIGH_EUCLID,
                          precision=t2d.PRECISION_FP32, capacity=N, lift_mode=t2d.LIFT_BARYCENTRIC)
        ctx.set_particles(uv, n)
        a, b = [], []
        for _ in range(steps):
            assert ctx.step(1) == 0
            o = ctx.observables()
            a.append(o["phi"]); b.append(o["mean_speed"])
        ctx.close()
        c, d = _series_oracle(orc, t2d, chart, uv, n, sigma, eta, 2000 + s, steps)
        g_phi.append(np.mean(a[tail])); g_spd.append(np.mean(b[tail]))
        o_phi.append(c[tail].mean()); o_spd.append(d[tail].mean())
    _stat_compare("barycentric phi", g_phi, o_phi, nseeds)
    _stat_compare("barycentric mean speed", g_spd, o_spd, nseeds)


def test_metric_table_statistics(t2d, chart, oracle_mod):
    """Row f4: the metric geodesic table (edge-length Dijkstra, DijkstraDistanceHelper.cpp:64-79) as thresholded CSR rows on
    the fp32 path against the oracle with the same distances as a dense table in fp64: order parameter and mean speed of
    long runs agree statistically."""
    N, steps, eta, nseeds = 1500, 100, 0.1, 5
    sigma = 0.35
    csr = t2d.TableCSR.geodesic(chart, 2.4 * sigma + 1e-9)
    dense = csr.to_dense(1e9)
    orc = oracle_mod.Oracle(chart)
    orc.set_table(dense)
    tail = slice(steps // 2, steps)
    g_phi, g_spd, o_phi, o_spd = [], [], [], []
    for s in range(nseeds):
        uv, n = t2d.seed_particles(N, seed=900 + s)
        ctx = t2d.Context(chart, table=csr, v0=0.1, k=0.05, sigma=sigma, step_size=0.001, eta=eta, seed=3000 + s,
                          neigh_mode=t2d.NEIGH_TABLE, precision=t2d.PRECISION_FP32, capacity=N)
        ctx.set_particles(uv, n)
        a, b = [], []
        for _ in range(steps):
            assert ctx.step(1) == 0
            o = ctx.observables()
            a.append(o["phi"]); b.append(o["mean_speed"])
        ctx.close()
        r3d, vid, _ = orc.get_r3d(uv)
        st = dict(uv=uv, n=n, vid=vid, r3d=r3d)
        c, d = [], []
        for k in range(steps):
            st = orc.step(st["uv"], st["n"], st["vid"], st["r3d"], 0.1, 0.05, sigma, 0.001, eta=eta, seed=3000 + s, mode=0, step_index=k)
            assert st["fault"] == 0
            p, v = orc.observables(st["n"], st["rdot"])
            c.append(p); d.append(v)
        g_phi.append(np.mean(a[tail])); g_spd.append(np.mean(b[tail]))
        o_phi.append(np.mean(c[tail])); o_spd.append(np.mean(d[tail]))
    _stat_compare("metric table phi", g_phi, o_phi, nseeds)
    _stat_compare("metric table mean speed", g_spd, o_spd, nseeds)


def test_replicas_are_independent(t2d, chart):
    """Config 5 (noise sweep as independent replicas, bench.py --workload c5): contexts that share a GPU and are stepped in
    turn do not see each other — each one reproduces the run of the same (eta, seed) alone, bit for bit in fp64 (the exact
    path sums in ascending id and its noise is keyed by (seed, step, id))."""
    N = 6000
    sigma = float(np.sqrt(0.5 * 451.3 / (np.pi * N)))
    reps = [(0.0, 11), (0.3, 12), (0.9, 13)]
    kw = dict(v0=0.1, k=1.0, sigma=sigma, step_size=0.001, neigh_mode=t2d.NEIGH_EUCLID, precision=t2d.PRECISION_FP64, capacity=N)

    def start(eta, seed):
        uv, n = t2d.seed_particles(N, seed=seed)
        ctx = t2d.Context(chart, eta=eta, seed=seed, **kw)
        ctx.set_particles(uv, n)
        return ctx

    alone = []
    for eta, seed in reps:
        ctx = start(eta, seed)
        assert ctx.step(12) == 0
        alone.append((ctx.download(), ctx.observables()))
        ctx.close()
    ctxs = [start(eta, seed) for eta, seed in reps]
    for _ in range(4):           # interleaved: 3 steps of every replica in turn
        for ctx in ctxs:
            assert ctx.step(3) == 0
    for ctx, (ref, obs) in zip(ctxs, alone):
        out = ctx.download()
        for k in ("uv", "n", "vid", "r3d", "rdot", "color", "face"):
            assert np.array_equal(out[k], ref[k]), k
        assert abs(ctx.observables()["phi"] - obs["phi"]) < 1e-12   # the reduction's atomic order is free, the state is not
        ctx.close()
    # more noise, less order
    phis = [o["phi"] for _, o in alone]
    print("phi(eta):", list(zip([e for e, _ in reps], phis)))
