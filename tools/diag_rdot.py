import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
t2d = importlib.import_module("2dtissue_b200")
from oracle import oraclebind
chart = t2d.load_chart(os.path.join(ROOT, "tests", "golden", "ellipsoid_x4.t2dchart"))
orc = oraclebind.Oracle(chart)
N = 30000
uv, n = t2d.seed_particles(N, seed=77 + N)
sigma = float(np.sqrt(0.5 * 451.3 / (np.pi * N)))
ctx = t2d.Context(chart, v0=0.1, k=1.0, sigma=sigma, step_size=0.001, neigh_mode=1, capacity=N)
ctx.set_particles(uv, n)
s0 = ctx.download()
o = orc.step(uv, n, s0["vid"], s0["r3d"], 0.1, 1.0, sigma, 0.001, mode=1)
ctx.set_state(uv, n, s0["vid"], s0["r3d"])
F, nh, col = ctx.forces()
print("color equal", np.array_equal(col, o["color"]), "F equal", np.array_equal(F, o["F"]))
bad = np.nonzero((F[:N] != o["F"][:N]) | (F[N:] != o["F"][N:]))[0]
print("F differs for", len(bad), "particles; max abs diff", np.abs(F - o["F"]).max())
for i in bad[:10]:
    print(i, "color", col[i], "F gpu", F[i], F[N + i], "F orc", o["F"][i], o["F"][N + i], "relx", (F[i] - o["F"][i]) / (abs(o["F"][i]) + 1e-300))
ctx.set_state(uv, n, s0["vid"], s0["r3d"])
ctx.step(1)
g = ctx.download()
badr = np.nonzero((g["rdot"][:N] != o["rdot"][:N]) | (g["rdot"][N:] != o["rdot"][N:]))[0]
print("rdot differs for", len(badr), "max abs", np.abs(g["rdot"] - o["rdot"]).max(), "subset of F-bad:", set(badr) <= set(bad))
print("counters", ctx.counters())
