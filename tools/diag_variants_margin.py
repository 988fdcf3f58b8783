"""dev: margins of tests/test_gpu_fastpath.py::test_fast2_variants_agree over repeated runs (run-to-run slot order differs)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
t2d = importlib.import_module("2dtissue_b200")
import test_gpu_fastpath as T
chart = t2d.load_chart("tests/golden/ellipsoid_x4.t2dchart")
N = 60000
sigma = T.sigma_for(N)
uv, n = t2d.seed_particles(N, seed=99)
ctx = t2d.Context(chart, v0=0.1, k=1.0, sigma=sigma, step_size=0.001, neigh_mode=t2d.NEIGH_EUCLID, precision=t2d.PRECISION_FP32, capacity=N)
ctx.set_particles(uv, n); ctx.step(30); state = ctx.download(); ctx.close()
for rep in range(6):
    a = T._one_step_variant(t2d, chart, {}, state, sigma, N)
    for name, env, tl in (("full", {"T2D_LEAN": "0"}, False), ("ties", {}, True), ("legacy", {"T2D_STEP": "legacy"}, False)):
        b = T._one_step_variant(t2d, chart, env, state, sigma, N, tie_log=tl)
        sp = np.maximum(np.hypot(a["rdot"][:N], a["rdot"][N:]), 0.1)
        same = a["n"] == b["n"]
        print(rep, name, "n mismatch %.2e (<1e-3)" % np.mean(a["n"] != b["n"]),
              "rdot %.2e (<1e-5)" % np.max(np.hypot(a["rdot"][:N] - b["rdot"][:N], a["rdot"][N:] - b["rdot"][N:]) / sp),
              "face %.2e (<1e-4)" % np.mean(a["face"][same] != b["face"][same]),
              "uv %.2e (<1e-6)" % np.max(np.abs(a["uv"] - b["uv"])[np.concatenate([same, same])]),
              "color eq", np.array_equal(a["color"], b["color"]))
