"""Chart files: the UV chart of a closed triangle mesh as the reference's setup produces it.

A chart is what `_2DTissue`'s constructor leaves behind for the stepping loop
(/root/reference/src/simulation/2DTissue.cpp:85-107): `vertice_UV`, `vertice_3D` (both V rows, indexed by
the cut-open mesh's vertex id), `face_UV` (F x 3 vertex ids) and the border polygon.  The chart is made by
the reference's host-side MeshCartographyLib (which stays the host-side surface, BASELINE.json north_star);
this module only stores / loads / refines it.

Binary layout (little endian):  8s magic "T2DCHART" | u32 version=1 | u32 V | u32 F | u32 P | 8 pad bytes |
f64 uv[V][2] | f64 x3d[V][3] | i32 faces[F][3] | f64 polygon[P][2].   The C++ host driver reads the same file
(2dtissue_b200/csrc/host/chart_io.h).
"""
import struct

import numpy as np

MAGIC = b"T2DCHART"


def save_chart(path, chart):
    uv = np.ascontiguousarray(chart["uv"], dtype="<f8")
    x3d = np.ascontiguousarray(chart["x3d"], dtype="<f8")
    faces = np.ascontiguousarray(chart["faces"], dtype="<i4")
    poly = np.ascontiguousarray(chart.get("polygon", np.zeros((0, 2))), dtype="<f8")
    with open(path, "wb") as f:
        f.write(struct.pack("<8sIIII8x", MAGIC, 1, len(uv), len(faces), len(poly)))
        f.write(uv.tobytes())
        f.write(x3d.tobytes())
        f.write(faces.tobytes())
        f.write(poly.tobytes())


def load_chart(path):
    with open(path, "rb") as f:
        magic, ver, V, F, P = struct.unpack("<8sIIII8x", f.read(32))
        if magic != MAGIC or ver != 1:
            raise ValueError("%s is not a T2DCHART v1 file" % path)
        uv = np.frombuffer(f.read(16 * V), dtype="<f8").reshape(V, 2).copy()
        x3d = np.frombuffer(f.read(24 * V), dtype="<f8").reshape(V, 3).copy()
        faces = np.frombuffer(f.read(12 * F), dtype="<i4").reshape(F, 3).copy()
        poly = np.frombuffer(f.read(16 * P), dtype="<f8").reshape(P, 2).copy()
    return dict(uv=uv, x3d=x3d, faces=faces, polygon=poly)


def ellipsoid_axes(x3d):
    """Semi-axes and centre of the axis-aligned ellipsoid the chart's 3-D vertices lie on."""
    lo, hi = x3d.min(axis=0), x3d.max(axis=0)
    return 0.5 * (hi - lo), 0.5 * (hi + lo)


def refine_chart(chart, levels=1, project_to_ellipsoid=True):
    """1->4 midpoint subdivision of a chart ("high-resolution ellipsoid", SURVEY.md §8d synthetic inputs).

    New UV vertices are edge midpoints in UV; their 3-D positions are edge midpoints pushed back onto the
    analytic ellipsoid fitted to the original vertices (so UV<->3-D stays a consistent chart).  All
    coordinates are rounded to float32 like the reference's pmp::Scalar=float mesh storage.
    Vertex ids 0..V-1 of the input keep their ids.
    """
    uv = np.asarray(chart["uv"], dtype=np.float64)
    x3d = np.asarray(chart["x3d"], dtype=np.float64)
    faces = np.asarray(chart["faces"], dtype=np.int64)
    axes, centre = ellipsoid_axes(x3d)
    for _ in range(levels):
        V = len(uv)
        e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=0)
        e.sort(axis=1)
        key = e[:, 0] * V + e[:, 1]
        ukey, inv = np.unique(key, return_inverse=True)
        a, b = ukey // V, ukey % V
        mid_uv = 0.5 * (uv[a] + uv[b])
        mid_x = 0.5 * (x3d[a] + x3d[b])
        if project_to_ellipsoid:
            q = (mid_x - centre) / axes
            q /= np.linalg.norm(q, axis=1, keepdims=True)
            mid_x = centre + q * axes
        uv = np.concatenate([uv, mid_uv]).astype(np.float32).astype(np.float64)
        x3d = np.concatenate([x3d, mid_x]).astype(np.float32).astype(np.float64)
        F = len(faces)
        m01, m12, m20 = V + inv[:F], V + inv[F:2 * F], V + inv[2 * F:]
        v0, v1, v2 = faces[:, 0], faces[:, 1], faces[:, 2]
        faces = np.concatenate([
            np.stack([v0, m01, m20], 1), np.stack([m01, v1, m12], 1),
            np.stack([m20, m12, v2], 1), np.stack([m01, m12, m20], 1)], axis=0)
    return dict(uv=uv, x3d=x3d, faces=faces.astype(np.int32), polygon=np.asarray(chart.get("polygon", np.zeros((0, 2)))))
