import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
t2d = importlib.import_module("2dtissue_b200")
chart = t2d.load_chart(os.path.join(ROOT, "tests", "golden", "ellipsoid_x4.t2dchart"))
def T(msg, t0): print("%-40s %.3f s" % (msg, time.perf_counter() - t0), flush=True)
for N in (100_000, 1_000_000):
    t0 = time.perf_counter()
    ctx = t2d.Context(chart, table_kind=t2d.TABLE_HOPS_FROM_MESH, neigh_mode=t2d.NEIGH_TABLE, precision=t2d.PRECISION_FP32, capacity=N)
    T("create N=%d" % N, t0)
    uv, n = t2d.seed_particles(N, seed=1234)
    t0 = time.perf_counter(); ctx.set_particles(uv, n); T("set_particles", t0)
    for k in range(3):
        t0 = time.perf_counter(); f = ctx.step(1); T("step fault=%d dev_ms=%.3f" % (f, ctx.last_step_ms), t0)
    print(ctx.profile_step(), ctx.counters(), flush=True)
    ctx.close()
