"""dev tool: per-warp (start, end) timeline of one k_step_fast2 launch on the headline workload.  Needs a library built
with -DT2D_F2_TIMELINE (tools/build_variant.sh timeline "-DT2D_F2_TIMELINE=1").  usage: diag_timeline.py LIB"""
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
chart = load_chart(t2d, 0)
ctx = t2d.Context(chart, v0=0.1, k=1.0, sigma=sigma_for(N), step_size=0.001, neigh_mode=t2d.NEIGH_EUCLID,
                  precision=t2d.PRECISION_FP32, capacity=N)
uv, n = t2d.seed_particles(N, seed=1234)
ctx.set_particles(uv, n)
ctx.step(int(sys.argv[2]) if len(sys.argv) > 2 else 200)
L = C.CDLL("2dtissue_b200/lib2dtissue_b200.so")
buf = np.zeros(2 * 8192, dtype=np.uint64)
rs = np.zeros(4 * 131072, dtype=np.uint32)


def makespan(dur, order, workers):
    """list scheduling: `workers` warps pull rows in `order`; returns (makespan, mean finish)"""
    import heapq
    h = [0.0] * workers
    heapq.heapify(h)
    for r in order:
        t = heapq.heappop(h)
        heapq.heappush(h, t + dur[r])
    return max(h), float(np.mean(h))


for rep in range(3):
    ctx.step(1)
    rc = L.t2d_dev_timeline(buf.ctypes.data_as(C.POINTER(C.c_ulonglong)))
    tl = buf.reshape(-1, 2).astype(np.int64)
    tl = tl[tl[:, 1] > 0]
    t0 = tl[:, 0].min()
    st, en = (tl[:, 0] - t0) / 1e3, (tl[:, 1] - t0) / 1e3
    q = [0, 1, 5, 25, 50, 75, 95, 99, 100]
    print("rep", rep, "rc", rc, "warps", len(tl))
    print("  start us pct", dict(zip(q, np.percentile(st, q).round(1))))
    print("  end   us pct", dict(zip(q, np.percentile(en, q).round(1))))
    # occupancy integral: fraction of warp-time between global start and global end that warps are alive
    T = en.max()
    print("  kernel span %.1f us; mean warp alive %.1f us (%.1f%%)" % (T, (en - st).mean(), 100 * (en - st).mean() / T))
    # per SM last end: block b -> 4 warps; SM unknown, use block
    ends = np.sort(en)[::-1]
    print("  last 10 ends", ends[:10].round(1), " 100th", ends[100].round(1), "1000th", ends[1000].round(1))
    if len(sys.argv) > 3 and sys.argv[3] == "warps":
        continue
    L.t2d_dev_rowstat(rs.ctypes.data_as(C.POINTER(C.c_uint)))
    R = rs.reshape(-1, 4).astype(np.int64)
    R = R[R[:, 1] > 0]
    dur, mx, sm = R[:, 1] / 1e3, R[:, 2], R[:, 3]
    print("  rows", len(R), "dur us pct", dict(zip(q, np.percentile(dur, q).round(1))), "mean %.1f" % dur.mean())
    print("  max-lane trips pct", dict(zip(q, np.percentile(mx, q).round(0))), "mean %.1f" % mx.mean())
    print("  corr(dur, maxtrips) %.3f  corr(dur, sumtrips) %.3f" % (np.corrcoef(dur, mx)[0, 1], np.corrcoef(dur, sm)[0, 1]))
    st_r = (R[:, 0] - R[:, 0].min()) / 1e3
    o = np.argsort(st_r)
    print("  corr(row start, maxtrips) %.3f; mean trips of first 5000 started %.1f, middle %.1f, last 5000 %.1f" % (
        np.corrcoef(st_r, mx)[0, 1], mx[o[:5000]].mean(), mx[o[len(o) // 2 - 2500:len(o) // 2 + 2500]].mean(), mx[o[-5000:]].mean()))
    late = o[-20:]
    print("  last 20 started rows: start", st_r[late].round(0), "trips", mx[late], "dur", dur[late].round(0))
    nat = np.arange(len(R))
    print("  simulated makespan natural order: %.1f (mean finish %.1f)" % makespan(dur, nat, len(tl)))
    print("  simulated makespan longest-first by duration: %.1f (mean %.1f)" % makespan(dur, np.argsort(-dur), len(tl)))
    print("  simulated makespan longest-first by trips:    %.1f (mean %.1f)" % makespan(dur, np.argsort(-mx, kind="stable"), len(tl)))
    heavy = mx > 1.3 * mx.mean()
    order = np.concatenate([nat[heavy], nat[~heavy]])
    print("  simulated makespan heavy(>1.3 mean trips, %d rows) first, rest natural: %.1f (mean %.1f)" % ((heavy.sum(),) + makespan(dur, order, len(tl))))
