import sys, importlib
import numpy as np
sys.path.insert(0, "/root/repo")
t2d = importlib.import_module("2dtissue_b200")
chart = t2d.load_chart("/root/repo/tests/golden/ellipsoid_x4.t2dchart")
N = 20000
sigma = float(np.sqrt(0.5 * 451.3 / (np.pi * N)))
uv, n = t2d.seed_particles(N, seed=7)
ctx = t2d.Context(chart, v0=0.1, k=1.0, sigma=sigma, step_size=0.001, neigh_mode=t2d.NEIGH_EUCLID, precision=t2d.PRECISION_FP32, capacity=N)
ctx.set_particles(uv, n)
for s in range(6):
    print(s, ctx.step(1)); 
print(ctx.counters())
