"""dev tool: how well does a row's weight (max-lane trips) at step s predict the weight of the same row index at s+1?
needs the -DT2D_F2_TIMELINE build.  usage: diag_rowshift.py LIB [steps]"""
import ctypes as C
import importlib
import shutil
import sys

import numpy as np

sys.path.insert(0, ".")
shutil.copy(sys.argv[1], "2dtissue_b200/lib2dtissue_b200.so")
t2d = importlib.import_module("2dtissue_b200")
from bench import load_chart, sigma_for  # noqa: E402

N = 2_000_000
ctx = t2d.Context(load_chart(t2d, 0), v0=0.1, k=1.0, sigma=sigma_for(N), step_size=0.001, neigh_mode=t2d.NEIGH_EUCLID,
                  precision=t2d.PRECISION_FP32, capacity=N)
uv, n = t2d.seed_particles(N, seed=1234)
ctx.set_particles(uv, n)
ctx.step(int(sys.argv[2]) if len(sys.argv) > 2 else 200)
L = C.CDLL("2dtissue_b200/lib2dtissue_b200.so")
rs = np.zeros(4 * 131072, dtype=np.uint32)
W = []
for rep in range(4):
    ctx.step(1)
    L.t2d_dev_rowstat(rs.ctypes.data_as(C.POINTER(C.c_uint)))
    W.append(rs.reshape(-1, 4)[:62500, 2].astype(np.int64).copy())
np.save("gpurun_out/rowweights.npy", np.array(W))
for lag in (1, 2):
    a, b = W[0], W[lag]
    print("lag", lag, "corr same index %.3f" % np.corrcoef(a, b)[0, 1])
    for win in (0, 1, 2, 4, 8, 16):
        pred = a.copy()
        for d in range(1, win + 1):
            pred = np.maximum(pred, np.roll(a, d))
            pred = np.maximum(pred, np.roll(a, -d))
        heavy = b > 3 * b.mean()
        miss = (heavy & (pred < 0.6 * b)).sum()
        print("  window +-%d: corr %.3f; truly heavy rows %d, under-predicted (<0.6x) %d; rows predicted >2x mean: %d" % (
            win, np.corrcoef(pred, b)[0, 1], heavy.sum(), miss, (pred > 2 * b.mean()).sum()))
    # best shift per block of 1000 rows
    sh = []
    for blk in range(0, 62000, 4000):
        best = max(range(-40, 41), key=lambda d: np.corrcoef(np.roll(a, d)[blk:blk + 4000], b[blk:blk + 4000])[0, 1])
        sh.append(best)
    print("  best shift per 4000-row block:", sh)
