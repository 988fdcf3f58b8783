"""dev: table mode on the refined chart (config 3) — fault mask, speeds and kernel times for a few force constants."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t2d = importlib.import_module("2dtissue_b200")
chart = t2d.load_chart("tests/golden/ellipsoid_x4.t2dchart")
if len(sys.argv) > 1 and sys.argv[1] == "det":
    from oracle import oraclebind
    D = oraclebind.Oracle(chart).build_hop_table()
    N = 3000
    uv, n = t2d.seed_particles(N, seed=77)
    outs = []
    for tab in (D, D, t2d.TableCSR.from_dense(D, 1.0)):
        ctx = t2d.Context(chart, table=tab, sigma=0.4166666666666667, neigh_mode=t2d.NEIGH_TABLE, precision=0, capacity=N)
        ctx.set_particles(uv, n)
        f = ctx.step(2)
        outs.append((f, ctx.download(), ctx.counters()))
        ctx.close()
    for name, (f, o, c) in zip(("dense", "dense again", "csr"), outs):
        print(name, f, c["pairs_in_range"], {k: (bool(np.array_equal(o[k], outs[0][1][k])), float(np.max(np.abs(o[k].astype(float) - outs[0][1][k].astype(float))))) for k in o})
    sys.exit(0)
fine = t2d.refine_chart(chart, 2)
N = 1_000_000
uv, n = t2d.seed_particles(N, seed=1234)
for k in (1.0, 0.1, 0.01):
    ctx = t2d.Context(fine, table_kind=t2d.TABLE_HOPS_FROM_MESH, k=k, neigh_mode=t2d.NEIGH_TABLE, precision=t2d.PRECISION_FP32, capacity=N)
    ctx.set_particles(uv, n)
    t0 = time.time()
    faults = [ctx.step(5) for _ in range(6)]
    o = ctx.observables()
    print("k=%g faults %s mean speed %.4g phi %.3f lost %d  %.1f ms/step  %s" % (k, faults, o["mean_speed"], o["phi"], o["lost"], ctx.last_step_ms / 5, ctx.profile_step()))
    ctx.close()
