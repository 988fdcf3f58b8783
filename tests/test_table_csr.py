"""Row f1 / g2: the vertex-distance table as thresholded CSR rows instead of a dense V x V array.
CPU part: the host helpers (from_dense / hops / geodesic / binary cache).  GPU part: a context fed CSR rows computes exactly
what a context fed the dense table computes, the GPU hop-CSR builder equals the dense builder, and config 3's shape
(1 M particles, table criterion, refined chart, no V x V array anywhere) runs with fault mask 0."""
import os

import numpy as np
import pytest


def test_csr_helpers_cpu(t2d, chart, oracle, tmp_path):
    D = oracle.build_hop_table()
    for radius in (1, 3):
        a = t2d.TableCSR.from_dense(D, radius)
        b = t2d.TableCSR.hops(chart, radius)
        assert a.val.dtype == np.uint8 and b.val.dtype == np.uint8
        assert np.array_equal(a.start, b.start) and np.array_equal(a.col, b.col) and np.array_equal(a.val, b.val)
        assert np.all(np.diff(a.start) >= 1)                       # the diagonal is always there
    g = t2d.TableCSR.geodesic(chart, 0.9)
    assert g.val.dtype == np.float64 and g.val.max() <= 0.9
    rows = np.repeat(np.arange(g.V), np.diff(g.start))
    assert np.all(g.val[rows == g.col] == 0.0)
    # symmetric: D(v, u) listed iff D(u, v) listed, same value (to rounding of the two Dijkstra runs)
    G = g.to_dense(np.inf)
    assert np.array_equal(np.isfinite(G), np.isfinite(G.T))
    m = np.isfinite(G)
    assert np.max(np.abs(G[m] - G.T[m])) < 1e-12
    p = str(tmp_path / "g.t2dcsr")
    g.save(p)
    h = t2d.TableCSR.load(p)
    assert h.radius == g.radius and np.array_equal(h.start, g.start) and np.array_equal(h.col, g.col) and np.array_equal(h.val, g.val)


def _same(a, b, precision, tag):
    """fp64: bit-identical.  fp32 fast path: the order of the particles inside a bucket comes from atomics, so sums differ
    in the last bits from run to run — neighbour sets (colour) exact, the rest to fp32 rounding."""
    if precision == 0:
        for k in ("uv", "n", "vid", "r3d", "rdot", "color", "face"):
            assert np.array_equal(a[k], b[k]), (tag, k)
        return
    assert np.array_equal(a["color"], b["color"]), tag
    assert np.mean(a["n"] != b["n"]) < 2e-3, tag
    N = a["n"].size
    sp = np.maximum(np.hypot(a["rdot"][:N], a["rdot"][N:]), 0.1)
    assert np.max(np.hypot(a["rdot"][:N] - b["rdot"][:N], a["rdot"][N:] - b["rdot"][N:]) / sp) < 1e-4, tag


def _run(t2d, chart, table, kind, N, sigma, precision, steps=1, color_factor=2.4):
    uv, n = t2d.seed_particles(N, seed=77)
    ctx = t2d.Context(chart, table=table, table_kind=kind, sigma=sigma, color_factor=color_factor, neigh_mode=t2d.NEIGH_TABLE,
                      precision=precision, capacity=N)
    ctx.set_particles(uv, n)
    fault = ctx.step(steps)
    out = ctx.download()
    c = ctx.counters()
    ctx.close()
    return fault, out, c


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [0, 1])
def test_csr_input_equals_dense_input(t2d, chart, hop_table, precision):
    """Hop table (uint8) and a metric table (double): CSR rows complete up to the interaction radius give bit-identical
    results to the dense table, in both precisions (the table predicates are evaluated on doubles on both paths)."""
    N = 3000
    for dense, sigma, radius, cf in ((hop_table, 0.4166666666666667, 1.0, 2.4), (hop_table, 1.2, 3.0, 2.4)):
        csr = t2d.TableCSR.from_dense(dense, radius)
        f0, a, c0 = _run(t2d, chart, dense, None, N, sigma, precision, color_factor=cf)
        f1, b, c1 = _run(t2d, chart, csr, None, N, sigma, precision, color_factor=cf)
        assert f0 == f1
        _same(a, b, precision, ("hops", sigma))
        assert c0["pairs_in_range"] == c1["pairs_in_range"] > 0
    g = t2d.TableCSR.geodesic(chart, 1.0)
    G = g.to_dense(1e9)
    sigma = 0.4
    f0, a, _ = _run(t2d, chart, G, None, N, sigma, precision)
    f1, b, _ = _run(t2d, chart, g, None, N, sigma, precision)
    assert f0 == f1
    _same(a, b, precision, "metric")
    with pytest.raises(t2d.T2DError):                      # rows complete up to 1.0 cannot serve an interaction radius of 1.2
        _run(t2d, chart, g, None, 100, 0.5, precision)


@pytest.mark.gpu
def test_gpu_hop_csr_builder_equals_dense_builder(t2d, chart, hop_table):
    """T2D_TABLE_HOPS_FROM_MESH through the sparse builder (k_hop_csr, forced here on the small chart) against the dense
    builder (k_hop_bfs, whose table is sha256-identical to the reference's): same step results; and a larger sigma later
    (t2d_set_params) rebuilds the rows for the new radius."""
    N = 3000
    f0, a, _ = _run(t2d, chart, hop_table, None, N, 0.4166666666666667, 0)
    os.environ["T2D_HOPS_SPARSE"] = "1"
    try:
        f1, b, _ = _run(t2d, chart, None, t2d.TABLE_HOPS_FROM_MESH, N, 0.4166666666666667, 0)
        uv, n = t2d.seed_particles(N, seed=77)
        ctx = t2d.Context(chart, table_kind=t2d.TABLE_HOPS_FROM_MESH, sigma=0.4166666666666667, neigh_mode=t2d.NEIGH_TABLE, capacity=N)
        ctx.set_params(sigma=1.2)                           # radius 1 -> 3 hops
        ctx.set_particles(uv, n)
        f2 = ctx.step(2)
        c = ctx.download()
        ctx.close()
    finally:
        os.environ.pop("T2D_HOPS_SPARSE", None)
    assert f0 == f1
    for k in ("uv", "n", "vid", "r3d", "rdot", "color", "face"):
        assert np.array_equal(a[k], b[k]), k
    f3, d, _ = _run(t2d, chart, hop_table, None, N, 1.2, 0, steps=2)
    assert f2 == f3
    for k in ("uv", "n", "vid", "color", "face"):
        assert np.array_equal(c[k], d[k]), k


@pytest.mark.gpu
def test_config3_1M_table_mode_refined_chart(t2d, chart):
    """BASELINE.json configs[2]: 1 M particles, table criterion — on the refined chart (74.9 k vertices, ~13 particles per
    vertex; the dense table would be 5.6 GB as uint8, 45 GB in the reference's format) with hop-count rows built on the GPU.
    20 steps, fault mask 0, nobody lost."""
    fine = t2d.refine_chart(chart, 2)
    N = 1_000_000
    uv, n = t2d.seed_particles(N, seed=4321)
    # k = 0.01 as in bench.py --workload c3: co-located particles feel 1000 k |u_i - u_j| each (d = 0 -> 0.001 rule)
    ctx = t2d.Context(fine, table_kind=t2d.TABLE_HOPS_FROM_MESH, k=0.01, neigh_mode=t2d.NEIGH_TABLE, precision=t2d.PRECISION_FP32, capacity=N)
    ctx.set_particles(uv, n)
    assert ctx.step(20) == 0
    s = ctx.download(("uv", "n", "vid", "color"))
    inside = (s["uv"][:N] >= 0) & (s["uv"][:N] <= 1) & (s["uv"][N:] >= 0) & (s["uv"][N:] <= 1)
    assert inside.all()
    c = ctx.counters()
    assert c["wrap_cap_hits"] == 0 and c["pairs_in_range"] > N
    assert np.all((s["vid"] >= 0) & (s["vid"] < ctx.V))
    o = ctx.observables()
    assert o["lost"] == 0 and np.isfinite(o["mean_speed"])
    print("c3 on the refined chart: pairs/particle %.1f, mean speed %.3g, phi %.3f" % (c["pairs_in_range"] / 20 / N, o["mean_speed"], o["phi"]))
    ctx.close()
