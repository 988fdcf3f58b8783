"""dev: local slab group (loopback transport) on the refined chart, fp32, to reproduce multi-GPU bench faults on one GPU"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
t2d = importlib.import_module("2dtissue_b200")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
refine = int(sys.argv[4]) if len(sys.argv) > 4 else 2
chart = t2d.load_chart(os.path.join(ROOT, "tests", "golden", "ellipsoid_x4.t2dchart"))
if refine:
    chart = t2d.refine_chart(chart, refine)
uv, n = t2d.seed_particles(N, seed=1234)
sigma = float(np.sqrt(0.5 * 451.3 / (np.pi * N)))
kw = dict(v0=0.1, k=1.0, sigma=sigma, step_size=0.001, neigh_mode=t2d.NEIGH_EUCLID, precision=t2d.PRECISION_FP32)
c0 = t2d.Context(chart, capacity=N, **kw)
c0.set_particles(uv, n)
s0 = c0.download(("uv", "n", "vid", "r3d"))
f0 = c0.step(steps)
ref = c0.download()
print("single: fault", f0, c0.counters())
c0.close()
cuts = t2d.slab_cuts(s0["r3d"][:N], world)
cap = int(N / world * 1.25) + 65536
grp = t2d.LocalSlabGroup(chart, world, cuts, capacity=cap, **kw)
grp.set_state(s0)
for s in range(steps):
    f = grp.step(1)
    print("step", s, "fault", f, [c.counters()["cell_fallbacks"] for c in grp.ctxs], [c.counters()["wrap_cap_hits"] for c in grp.ctxs])
    if f:
        break
out, owned = grp.download()
print("owned", owned, "n equal", float((out["n"] == ref["n"]).mean()), "finite", bool(np.all(np.isfinite(out["uv"]))))
