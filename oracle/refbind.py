"""ctypes binding for oracle/_ref/libt2d_ref.so — TEST INFRASTRUCTURE, not product code.

libt2d_ref.so is the UNMODIFIED reference (compiled where it lies under /root/reference by
oracle/Makefile) plus oracle/ref_harness.cpp.  Only tests/, tools/make_golden.py,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libt2d_ref.so")
MESH_DIR = os.path.join(HERE, "_ref", "mcl", "meshes")


def available():
    return os.path.exists(LIB_PATH)


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class Ref:
    """One process-wide reference world (the harness keeps a single static instance)."""

    def __init__(self):
        if not available():
            raise RuntimeError("oracle/_ref/libt2d_ref.so not built (make -C oracle ref)")
        self.L = C.CDLL(LIB_PATH)
        self.L.t2dref_last_error.restype = C.c_char_p
        self.L.t2dref_step.argtypes = [C.c_int, _dp, _ip, _ip, _dp, _dp, _ip, _dp, C.c_double, C.c_double,
                                       C.c_double, C.c_double, _dp, C.c_int]
        self.L.t2dref_force_orientation.argtypes = [C.c_int, _dp, _ip, _dp, C.c_double, C.c_double, _dp]
        self.V = self.F = self.P = 0

    def _chk(self, r, what):
        if r < 0:
            raise RuntimeError("%s failed (%d): %s" % (what, r, self.L.t2dref_last_error().decode()))
        return r

    # --- chart ---------------------------------------------------------------------------------
    def chart_from_mesh(self, mesh_name="ellipsoid_x4.off"):
        path = os.path.join(MESH_DIR, mesh_name)
        self._chk(self.L.t2dref_chart_from_mesh(path.encode()), "chart_from_mesh")
        return self.chart_export()

    def chart_export(self):
        V, F, P = C.c_int(), C.c_int(), C.c_int()
        self._chk(self.L.t2dref_chart_sizes(C.byref(V), C.byref(F), C.byref(P)), "chart_sizes")
        self.V, self.F, self.P = V.value, F.value, P.value
        uv = np.zeros((self.V, 2))
        x3d = np.zeros((self.V, 3))
        faces = np.zeros((self.F, 3), dtype=np.int32)
        poly = np.zeros((self.P, 2))
        self._chk(self.L.t2dref_chart_export(_d(uv), _d(x3d), _i(faces), _d(poly)), "chart_export")
        return dict(uv=uv, x3d=x3d, faces=faces, polygon=poly)

    def chart_import(self, chart):
        uv = np.ascontiguousarray(chart["uv"], dtype=np.float64)
        x3d = np.ascontiguousarray(chart["x3d"], dtype=np.float64)
        faces = np.ascontiguousarray(chart["faces"], dtype=np.int32)
        poly = np.ascontiguousarray(chart["polygon"], dtype=np.float64)
        self.V, self.F, self.P = len(uv), len(faces), len(poly)
        self._chk(self.L.t2dref_chart_import(self.V, self.F, self.P, _d(uv), _d(x3d), _i(faces), _d(poly)),
                  "chart_import")

    # --- table ---------------------------------------------------------------------------------
    def table_build(self):
        self._chk(self.L.t2dref_table_build(), "table_build")
        D = np.zeros((self.V, self.V))
        self._chk(self.L.t2dref_table_export(_d(D)), "table_export")
        return D

    def table_import(self, D):
        D = np.ascontiguousarray(D)
        if D.dtype == np.uint8:
            self._chk(self.L.t2dref_table_import_u8(D.shape[0], D.ctypes.data_as(C.POINTER(C.c_ubyte))), "table_import")
        else:
            D = np.ascontiguousarray(D, dtype=np.float64)
            self._chk(self.L.t2dref_table_import_f64(D.shape[0], _d(D)), "table_import")

    # --- per-function entry points ------------------------------------------------------------
    def inside(self, uv):
        uv = np.ascontiguousarray(uv, dtype=np.float64)
        N = uv.size // 2
        out = np.zeros(N, dtype=np.int32)
        self.L.t2dref_inside(N, _d(uv), _i(out))
        return out

    def angles_to_unit_vectors(self, n):
        n = np.ascontiguousarray(n, dtype=np.int32)
        out = np.zeros(2 * n.size)
        self.L.t2dref_angles_to_unit_vectors(n.size, _i(n), _d(out))
        return out

    def get_r3d(self, uv):
        uv = np.ascontiguousarray(uv, dtype=np.float64)
        N = uv.size // 2
        r3d = np.zeros(3 * N)
        vid = np.zeros(N, dtype=np.int32)
        self._chk(self.L.t2dref_get_r3d(N, _d(uv), _d(r3d), _i(vid)), "get_r3d")
        return r3d, vid

    def tiling(self, uv_old, uv, n):
        uv_old = np.array(uv_old, dtype=np.float64).copy()
        uv = np.array(uv, dtype=np.float64).copy()
        n = np.array(n, dtype=np.int32).copy()
        self._chk(self.L.t2dref_tiling(n.size, _d(uv_old), _d(uv), _i(n)), "tiling")
        return uv_old, uv, n

    def force_orientation(self, uv, n, dist_length, k, sigma):
        uv = np.ascontiguousarray(uv, dtype=np.float64)
        n = np.array(n, dtype=np.int32).copy()
        dl = np.ascontiguousarray(dist_length, dtype=np.float64)
        F = np.zeros(2 * n.size)
        self._chk(self.L.t2dref_force_orientation(n.size, _d(uv), _i(n), _d(dl), k, sigma, _d(F)), "force_orientation")
        return F, n

    def step(self, uv, n, vid, r3d, v0, k, sigma, step_size, eta=None, mode=0):
        """One reference timestep.  All arrays are copied; returns a dict of outputs + fault mask."""
        uv = np.array(uv, dtype=np.float64).copy()
        n = np.array(n, dtype=np.int32).copy()
        vid = np.array(vid, dtype=np.int32).copy()
        r3d = np.array(r3d, dtype=np.float64).copy()
        N = n.size
        rdot = np.zeros(2 * N)
        color = np.zeros(N, dtype=np.int32)
        F = np.zeros(2 * N)
        eta_p = None
        if eta is not None:
            eta = np.ascontiguousarray(eta, dtype=np.float64)
            eta_p = _d(eta)
        fault = self._chk(self.L.t2dref_step(N, _d(uv), _i(n), _i(vid), _d(r3d), _d(rdot), _i(color), _d(F),
                                             v0, k, sigma, step_size, eta_p, mode), "step")
        return dict(uv=uv, n=n, vid=vid, r3d=r3d, rdot=rdot, color=color, F=F, fault=fault)
