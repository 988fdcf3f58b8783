"""dev: t2d_step_host_uv at the bench's size, a few calls (run under `ncu --metrics gpu__time_duration.sum` for the kernel list)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
t2d = importlib.import_module("2dtissue_b200")
chart = t2d.refine_chart(t2d.load_chart("tests/golden/ellipsoid_x4.t2dchart"), 2)
N = 2_000_000
sigma = float(np.sqrt(0.5 * 451.3 / (np.pi * N)))
uv, n = t2d.seed_particles(N, seed=1234)
ctx = t2d.Context(chart, v0=0.1, k=1.0, sigma=sigma, step_size=0.001, neigh_mode=t2d.NEIGH_EUCLID, precision=t2d.PRECISION_FP32, capacity=N)
ctx.set_particles(uv, n)
ctx.step(30)
s = ctx.download(("uv", "n", "vid", "r3d"))
pin = lambda a: torch.from_numpy(a.copy()).pin_memory().numpy()
h = [pin(s["uv"]), pin(s["n"]), pin(s["vid"]), pin(s["r3d"]), pin(np.zeros(2 * N)), pin(np.zeros(N, dtype=np.int32))]
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ctx.step_host(*h, reproject=True)
print("ok")
