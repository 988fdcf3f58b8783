"""Vertex-distance tables in CSR form (include/t2d.h t2d_table_csr): only the vertex pairs that can ever interact.

The reference keeps the full V x V table of doubles in memory and caches it as CSV
(MeshCartographyLib CachedGeodesicDistanceHelper.h:22-72, DijkstraDistanceHelper.cpp:27-111); on a refined chart
(V = 75 k) that is 45 GB.  Stage 2 only ever reads d < 2 sigma and d <= color_factor sigma, so the table travels as rows
that are complete up to a radius:

    TableCSR.from_dense(D, radius)          thresholded copy of a dense table, min-symmetrised like Locomotion.cpp:110
    TableCSR.geodesic(chart, radius)        metric variant (DijkstraDistanceHelper.cpp:64-79): edge-length Dijkstra on the
                                            mesh's edge graph, cut off at the radius (host, scipy)
    TableCSR.hops(chart, radius)            hop counts, the stock table, cut off at the radius (host, scipy)
    save() / load()                         binary cache file (replaces the reference's CSV cache)

No arithmetic of the step happens here: these are setup inputs, like the chart itself.
"""
import struct

import numpy as np

MAGIC = b"T2DCSR1\0"


class TableCSR:
    def __init__(self, start, col, val, radius):
        self.start = np.ascontiguousarray(start, dtype=np.int32)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        val = np.ascontiguousarray(val)
        if val.dtype not in (np.dtype(np.float64), np.dtype(np.float32), np.dtype(np.uint8)):
            val = val.astype(np.float64)
        self.val = val
        self.radius = float(radius)
        self.V = self.start.size - 1
        if self.start[0] != 0 or self.start[-1] != self.col.size or self.col.size != self.val.size:
            raise ValueError("inconsistent CSR arrays")

    @property
    def nnz(self):
        return int(self.col.size)

    # ---- constructors -------------------------------------------------------------------------------------------
    @classmethod
    def from_dense(cls, D, radius, dtype=None):
        """Rows of min(D, D^T) (Locomotion.cpp:110) with entries <= radius; the diagonal is always kept."""
        D = np.asarray(D)
        S = np.minimum(D, D.T)
        keep = S <= radius
        np.fill_diagonal(keep, True)
        rows, cols = np.nonzero(keep)          # row-major: ascending column inside a row
        start = np.zeros(D.shape[0] + 1, dtype=np.int64)
        np.add.at(start, rows + 1, 1)
        start = np.cumsum(start)
        val = S[rows, cols]
        return cls(start, cols, val.astype(dtype or D.dtype), radius)

    @classmethod
    def _graph(cls, chart, metric):
        from scipy.sparse import coo_matrix
        f = np.asarray(chart["faces"], dtype=np.int64)
        V = len(chart["uv"])
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        e = np.unique(np.sort(e, axis=1), axis=0)
        if metric:
            x = np.asarray(chart["x3d"], dtype=np.float64)
            w = np.linalg.norm(x[e[:, 0]] - x[e[:, 1]], axis=1)
        else:
            w = np.ones(len(e))
        g = coo_matrix((np.concatenate([w, w]), (np.concatenate([e[:, 0], e[:, 1]]), np.concatenate([e[:, 1], e[:, 0]]))), shape=(V, V))
        return g.tocsr()

    @classmethod
    def _dijkstra(cls, chart, radius, metric, chunk=4096):
        from scipy.sparse.csgraph import dijkstra
        g = cls._graph(chart, metric)
        V = g.shape[0]
        starts, cols, vals = [0], [], []
        for a in range(0, V, chunk):
            idx = np.arange(a, min(V, a + chunk))
            d = dijkstra(g, directed=False, indices=idx, limit=radius)      # inf beyond the limit
            r, c = np.nonzero(np.isfinite(d))
            cnt = np.bincount(r, minlength=len(idx))
            for k in cnt:
                starts.append(starts[-1] + int(k))
            cols.append(c)
            vals.append(d[r, c])
        col = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
        val = np.concatenate(vals) if vals else np.zeros(0)
        if metric:
            # D(v, u) and D(u, v) come from two Dijkstra runs and differ in the last bits: symmetrise by min over the union
            # of the two patterns, exactly what Locomotion.cpp:110 does with the dense table
            row = np.repeat(np.arange(V), np.diff(np.asarray(starts)))
            r2, c2, v2 = np.concatenate([row, col]), np.concatenate([col, row]), np.concatenate([val, val])
            order = np.lexsort((c2, r2))
            r2, c2, v2 = r2[order], c2[order], v2[order]
            first = np.ones(r2.size, dtype=bool)
            first[1:] = (r2[1:] != r2[:-1]) | (c2[1:] != c2[:-1])
            idx = np.nonzero(first)[0]
            val = np.minimum.reduceat(v2, idx)
            col = c2[idx]
            starts = np.concatenate([[0], np.cumsum(np.bincount(r2[idx], minlength=V))])
        return cls(np.asarray(starts), col, val if metric else val.astype(np.uint8), radius)

    @classmethod
    def geodesic(cls, chart, radius):
        """Edge-length shortest paths over the mesh's edge graph (the metric variant of DijkstraDistanceHelper.cpp:64-79),
        complete up to `radius` (mesh units).  Symmetric by construction (undirected graph, one Dijkstra per source)."""
        return cls._dijkstra(chart, radius, True)

    @classmethod
    def hops(cls, chart, radius):
        """Hop counts over the mesh's edge graph (the stock table), complete up to `radius` hops."""
        return cls._dijkstra(chart, float(int(radius)), False)

    # ---- binary cache -------------------------------------------------------------------------------------------
    def save(self, path):
        """MAGIC, V (int64), nnz (int64), radius (double), value type code (int64: 8 double, 4 float, 1 uint8), then start
        (int32[V + 1]), col (int32[nnz]), val — little endian, no padding; the C++ driver reads the same file."""
        with open(path, "wb") as f:
            f.write(MAGIC)
            f.write(struct.pack("<qqdq", self.V, self.nnz, self.radius, self.val.dtype.itemsize))
            f.write(self.start.tobytes())
            f.write(self.col.tobytes())
            f.write(self.val.tobytes())

    @classmethod
    def load(cls, path):
        with open(path, "rb") as f:
            if f.read(8) != MAGIC:
                raise ValueError("not a T2DCSR1 table cache: " + path)
            V, nnz, radius, item = struct.unpack("<qqdq", f.read(32))
            start = np.frombuffer(f.read(4 * (V + 1)), dtype=np.int32)
            col = np.frombuffer(f.read(4 * nnz), dtype=np.int32)
            dt = {8: np.float64, 4: np.float32, 1: np.uint8}[item]
            val = np.frombuffer(f.read(item * nnz), dtype=dt)
        return cls(start, col, val, radius)

    def to_dense(self, fill):
        """Dense V x V copy with `fill` where the CSR has no entry (tests on small charts only)."""
        D = np.full((self.V, self.V), fill, dtype=np.float64 if self.val.dtype != np.uint8 else np.uint8)
        rows = np.repeat(np.arange(self.V), np.diff(self.start))
        D[rows, self.col] = self.val
        return D
