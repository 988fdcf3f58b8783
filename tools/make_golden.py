#!/usr/bin/env python
"""Generate tests/golden/* from the UNMODIFIED reference (oracle/_ref/libt2d_ref.so).

Runs only in the build container (needs /root/reference to have been compiled by `make -C oracle ref`).
The fixtures it writes are committed; the GPU box never runs this.

    python tools/make_golden.py            # writes tests/golden/*.t2dchart, *.npz

Fixtures
  ellipsoid_x4.t2dchart      chart produced by SurfaceParametrization::create_uv_surface on the default mesh
  table_pins.npz             sha256 + sampled rows of the reference's vertex-distance table (hop counts)
  kat_tiling.npz             the reference's own EuclideanTiling KAT evaluated by the reference + random cases
  get_r3d_N4000.npz          CellHelper::get_r3d on seeded UV points (incl. points on edges/vertices)
  step_*.npz                 one-step in/out pairs of the full timestep in table / euclid mode
"""
import hashlib
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402

chart_mod = importlib.import_module("2dtissue_b200.chart")
GOLD = os.path.join(ROOT, "tests", "golden")


def seeded_state(N, seed, box=None):
    rng = np.random.default_rng(seed)
    u = rng.random(N)
    v = rng.random(N)
    if box is not None:
        u = box[0] + u * (box[1] - box[0])
        v = box[2] + v * (box[3] - box[2])
    uv = np.concatenate([u, v])
    n = rng.integers(0, 360, N).astype(np.int32)
    return uv, n


def metric_table(x3d):
    """Synthetic metric table: float32(Euclidean vertex distance), widened to double."""
    X = np.asarray(x3d, dtype=np.float64)
    D = np.zeros((len(X), len(X)), dtype=np.float32)
    for i0 in range(0, len(X), 512):
        d = X[i0:i0 + 512, None, :] - X[None, :, :]
        D[i0:i0 + 512] = np.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]).astype(np.float32)
    return D


def run_steps(ref, uv, n, nsteps, v0, k, sigma, h, mode, eta_fn=None):
    r3d, vid = ref.get_r3d(uv)
    out = dict(uv0=uv, n0=n, vid0=vid, r3d0=r3d, params=np.array([v0, k, sigma, h]), mode=np.int32(mode))
    for s in range(nsteps):
        eta = eta_fn(s, n.size) if eta_fn else None
        o = ref.step(uv, n, vid, r3d, v0, k, sigma, h, eta=eta, mode=mode)
        for key in ("uv", "n", "vid", "r3d", "rdot", "color", "F"):
            out["%s%d" % (key, s + 1)] = o[key]
        out["fault%d" % (s + 1)] = np.int32(o["fault"])
        if eta is not None:
            out["eta%d" % (s + 1)] = eta
        uv, n, vid, r3d = o["uv"], o["n"], o["vid"], o["r3d"]
    out["nsteps"] = np.int32(nsteps)
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = refbind.Ref()
    chart = ref.chart_from_mesh("ellipsoid_x4.off")
    # coordinates are float32 values widened to double (pmp::Scalar = float)
    assert np.array_equal(chart["uv"], chart["uv"].astype(np.float32).astype(np.float64))
    assert np.array_equal(chart["x3d"], chart["x3d"].astype(np.float32).astype(np.float64))
    chart_mod.save_chart(os.path.join(GOLD, "ellipsoid_x4.t2dchart"), chart)
    print("chart: V=%d F=%d P=%d" % (ref.V, ref.F, ref.P))

    D = ref.table_build()
    assert np.array_equal(D, np.round(D)) and D.max() < 255
    D8 = D.astype(np.uint8)
    rows = np.array([0, 1, 17, 100, 2000, 4000, ref.V - 1])
    np.savez_compressed(os.path.join(GOLD, "table_pins.npz"), sha256=hashlib.sha256(D8.tobytes()).hexdigest(),
                        rows=rows, row_values=D8[rows], max_hop=np.int32(D8.max()), V=np.int32(ref.V))
    print("table: max hop %d sha %s" % (D8.max(), hashlib.sha256(D8.tobytes()).hexdigest()[:16]))

    # inside-predicate pins (check_point_in_polygon on the real polygon)
    rng = np.random.default_rng(7)
    pts = rng.uniform(-0.5, 1.5, (4000, 2))
    edge = np.array([[0, 0], [1, 0], [1, 1], [0, 1], [0.5, 0], [0, 0.5], [1, 0.25], [0.75, 1], [0.5, 0.5], [1.0000001, 0.5],
                     [-1e-12, 0.5], [0.5, 1 + 1e-15], [np.nextafter(1, 2), 0.3], [np.nextafter(0, -1), 0.3]], dtype=np.float64)
    pts = np.concatenate([pts, edge])
    uvp = np.concatenate([pts[:, 0], pts[:, 1]])
    np.savez_compressed(os.path.join(GOLD, "inside_pins.npz"), uv=uvp, inside=ref.inside(uvp))

    # --- tiling: the reference's own KAT (tests/simulation/test_EuclideanTiling.cpp:44-72) + random cases
    old = np.array([0.5, 0.5, 0.9, 0.5, 0.5, 0.4])  # column-major (x's then y's)
    new = np.array([2.5, 1.3, -0.3, 0.5, 1.2, 0.7])
    nk = np.array([80, 120, 42], dtype=np.int32)
    o_old, o_new, o_n = ref.tiling(old, new, nk)
    rng = np.random.default_rng(11)
    M = 3000
    r_old = rng.random((M, 2))
    step = rng.normal(0, 0.6, (M, 2))
    step[:500] *= 0.02  # short hops near the border
    r_old[:500] = np.where(rng.random((500, 2)) < 0.5, rng.random((500, 2)) * 0.01, 1 - rng.random((500, 2)) * 0.01)
    r_new = r_old + step
    # exact-border and corner cases
    r_old[-4:] = [[0.5, 0.5], [0.0, 0.5], [0.5, 1.0], [0.2, 0.2]]
    r_new[-4:] = [[1.5, 1.5], [-0.25, 0.5], [0.5, 1.75], [-0.3, -0.3]]
    rn = rng.integers(0, 360, M).astype(np.int32)
    ro = np.concatenate([r_old[:, 0], r_old[:, 1]])
    rw = np.concatenate([r_new[:, 0], r_new[:, 1]])
    t_old, t_new, t_n = ref.tiling(ro, rw, rn)
    np.savez_compressed(os.path.join(GOLD, "kat_tiling.npz"), kat_old=old, kat_new=new, kat_n=nk, kat_out_old=o_old,
                        kat_out_new=o_new, kat_out_n=o_n, old=ro, new=rw, n=rn, out_old=t_old, out_new=t_new, out_n=t_n)
    print("tiling KAT ->", o_new, o_n)

    # --- get_r3d
    uv, _ = seeded_state(4000, 21)
    # add points exactly on mesh vertices, edge midpoints and face centroids
    V = chart["uv"]
    Fc = chart["faces"]
    sel = rng.integers(0, len(Fc), 300)
    extra = np.concatenate([V[Fc[sel[:100], 0]], 0.5 * (V[Fc[sel[100:200], 0]] + V[Fc[sel[100:200], 1]]),
                            (V[Fc[sel[200:], 0]] + V[Fc[sel[200:], 1]] + V[Fc[sel[200:], 2]]) / 3.0,
                            np.array([[0, 0], [1, 1], [0, 1], [1, 0], [0.5, 0], [1, 0.5]], dtype=np.float64)])
    N = 4000
    uv = np.concatenate([np.concatenate([uv[:N], extra[:, 0]]), np.concatenate([uv[N:], extra[:, 1]])])
    r3d, vid = ref.get_r3d(uv)
    np.savez_compressed(os.path.join(GOLD, "get_r3d_N4306.npz"), uv=uv, r3d=r3d, vid=vid)
    print("get_r3d: N=%d" % (uv.size // 2))

    # --- full steps, table mode (the reference's own hop-count table)
    uv, n = seeded_state(100, 1)
    np.savez_compressed(os.path.join(GOLD, "step_table_N100.npz"), **run_steps(ref, uv, n, 3, 0.02, 1.0, 0.4166666666666667, 0.001, 0))
    # dense patch so that many particles share a nearest vertex (d == 0 -> 0.001 branch, high speeds, seam jumps)
    uv, n = seeded_state(1500, 2, box=(0.0, 0.12, 0.3, 0.5))
    np.savez_compressed(os.path.join(GOLD, "step_table_dense_N1500.npz"), **run_steps(ref, uv, n, 3, 0.1, 1.0, 0.4166666666666667, 0.001, 0))
    # wide cutoff: hops 0..2 interact, hops <= 3 are counted for the colour
    uv, n = seeded_state(1500, 3)
    np.savez_compressed(os.path.join(GOLD, "step_table_wide_N1500.npz"), **run_steps(ref, uv, n, 2, 0.1, 1.0, 1.3, 0.001, 0))
    # with injected noise
    def eta_fn(s, N):
        return np.random.default_rng(100 + s).uniform(-40, 40, N)
    uv, n = seeded_state(800, 4)
    np.savez_compressed(os.path.join(GOLD, "step_table_noise_N800.npz"), **run_steps(ref, uv, n, 2, 0.1, 1.0, 1.3, 0.001, 0, eta_fn))
    print("table-mode steps done")

    # --- metric table (synthetic float32 distances), injected through the harness
    Dm = metric_table(chart["x3d"])
    ref.table_import(Dm.astype(np.float64))
    uv, n = seeded_state(1500, 5)
    np.savez_compressed(os.path.join(GOLD, "step_metric_N1500.npz"), **run_steps(ref, uv, n, 2, 0.1, 1.0, 0.4166666666666667, 0.001, 0))
    print("metric-table steps done")

    # --- euclid mode (extension): reference ForceHelper/OrientationHelper on ||X_i - X_j||
    uv, n = seeded_state(2000, 6)
    A = 451.3
    sigma = float(np.sqrt(0.5 * A / (np.pi * 2000)))
    np.savez_compressed(os.path.join(GOLD, "step_euclid_N2000.npz"), **run_steps(ref, uv, n, 3, 0.1, 1.0, sigma, 0.001, 1))
    uv, n = seeded_state(1500, 8, box=(0.45, 0.55, 0.45, 0.55))
    np.savez_compressed(os.path.join(GOLD, "step_euclid_dense_N1500.npz"), **run_steps(ref, uv, n, 2, 0.1, 1.0, 0.05, 0.001, 1))
    print("euclid-mode steps done")
    tot = sum(os.path.getsize(os.path.join(GOLD, f)) for f in os.listdir(GOLD))
    print("golden dir: %.1f KB" % (tot / 1024))


if __name__ == "__main__":
    main()
