"""ctypes binding for oracle/libt2d_oracle.so (the plain-C restatement) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libt2d_oracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_uint32)


class Params(C.Structure):
    _fields_ = [("v0", C.c_double), ("k", C.c_double), ("sigma", C.c_double), ("step_size", C.c_double),
                ("eta", C.c_double), ("color_factor", C.c_double), ("seed", C.c_uint64), ("mode", C.c_int),
                ("brute", C.c_int), ("threads", C.c_int)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("pairs_in_range", "ties_cutoff", "ties_trunc", "wraps", "wrap_cap_hits",
                                         "locate_fallbacks", "lost", "nonfinite")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


_L = None


def lib():
    global _L
    if _L is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.t2do_create.restype = C.c_void_p
        L.t2do_create.argtypes = [C.c_int, C.c_int, _dp, _dp, _ip]
        L.t2do_destroy.argtypes = [C.c_void_p]
        L.t2do_set_table_f64.argtypes = [C.c_void_p, C.c_int, _dp]
        L.t2do_set_table_u8.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_ubyte)]
        L.t2do_build_hop_table.argtypes = [C.c_void_p]
        L.t2do_table_u8.restype = C.POINTER(C.c_ubyte)
        L.t2do_table_u8.argtypes = [C.c_void_p]
        L.t2do_inject_noise.argtypes = [C.c_void_p, _dp]
        L.t2do_set_angle_out.argtypes = [C.c_void_p, _dp]
        L.t2do_set_lift_mode.argtypes = [C.c_void_p, C.c_int]
        L.t2do_get_r3d.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _ip, _ip, C.c_int, C.POINTER(Stats)]
        L.t2do_tiling.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _ip, C.POINTER(Stats)]
        L.t2do_step.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int, _dp, _ip, _ip, _dp, _up, C.c_uint64, _dp, _ip,
                                _dp, _ip, C.POINTER(Stats)]
        L.t2do_observables.argtypes = [C.c_int, _ip, _dp, _dp]
        L.t2do_mean_angle_deg.restype = C.c_double
        L.t2do_mean_angle_deg.argtypes = [_dp, C.c_int, _ip]
        L.t2do_repulsive_adhesion.argtypes = [C.c_double] * 7 + [_dp]
        L.t2do_angles_to_unit_vectors.argtypes = [C.c_int, _ip, _dp]
        L.t2do_get_dist_vect.argtypes = [C.c_int, _dp, _dp, _dp]
        L.t2do_symmetrize_min.argtypes = [C.c_int, _dp]
        L.t2do_inside.argtypes = [C.c_double, C.c_double]
        L.t2do_point_triangle_distance.restype = C.c_double
        L.t2do_point_triangle_distance.argtypes = [_dp, _dp, _dp, _dp]
        L.t2do_philox_uniform.restype = C.c_double
        L.t2do_philox_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.t2do_noise_deg.restype = C.c_double
        L.t2do_noise_deg.argtypes = [C.c_double, C.c_uint64, C.c_uint64, C.c_uint32]
        _L = L
    return _L


class Oracle:
    def __init__(self, chart):
        self.L = lib()
        self.uv = np.ascontiguousarray(chart["uv"], dtype=np.float64)
        self.x3d = np.ascontiguousarray(chart["x3d"], dtype=np.float64)
        self.faces = np.ascontiguousarray(chart["faces"], dtype=np.int32)
        self.V, self.F = len(self.uv), len(self.faces)
        self.ctx = self.L.t2do_create(self.V, self.F, _d(self.uv), _d(self.x3d), _i(self.faces))
        self.last_stats = None

    def set_lift_mode(self, mode):
        """0: the reference's distance-weighted lift; 1: barycentric (the extension of include/t2d.h T2D_LIFT_BARYCENTRIC)."""
        self.L.t2do_set_lift_mode(self.ctx, int(mode))

    def __del__(self):
        try:
            self.L.t2do_destroy(self.ctx)
        except Exception:
            pass

    def set_table(self, D):
        D = np.ascontiguousarray(D)
        if D.dtype == np.uint8:
            self.L.t2do_set_table_u8(self.ctx, D.shape[0], D.ctypes.data_as(C.POINTER(C.c_ubyte)))
        else:
            D = np.ascontiguousarray(D, dtype=np.float64)
            self.L.t2do_set_table_f64(self.ctx, D.shape[0], _d(D))

    def build_hop_table(self):
        self.L.t2do_build_hop_table(self.ctx)
        p = self.L.t2do_table_u8(self.ctx)
        return np.ctypeslib.as_array(p, shape=(self.V, self.V)).copy()

    def get_r3d(self, uv, brute=False):
        uv = np.ascontiguousarray(uv, dtype=np.float64)
        N = uv.size // 2
        r3d = np.zeros(3 * N)
        vid = np.zeros(N, dtype=np.int32)
        face = np.zeros(N, dtype=np.int32)
        st = Stats()
        self.L.t2do_get_r3d(self.ctx, N, _d(uv), _d(r3d), _i(vid), _i(face), int(brute), C.byref(st))
        self.last_stats = st.as_dict()
        return r3d, vid, face

    def tiling(self, uv_old, uv, n):
        uv_old = np.array(uv_old, dtype=np.float64).copy()
        uv = np.array(uv, dtype=np.float64).copy()
        n = np.array(n, dtype=np.int32).copy()
        st = Stats()
        self.L.t2do_tiling(self.ctx, n.size, _d(uv_old), _d(uv), _i(n), C.byref(st))
        self.last_stats = st.as_dict()
        return uv_old, uv, n

    def step(self, uv, n, vid, r3d, v0, k, sigma, step_size, eta=0.0, seed=0, mode=0, brute=False, threads=0,
             ids=None, step_index=0, color_factor=2.4, eta_inject=None):
        uv = np.array(uv, dtype=np.float64).copy()
        n = np.array(n, dtype=np.int32).copy()
        vid = np.array(vid, dtype=np.int32).copy()
        r3d = np.array(r3d, dtype=np.float64).copy()
        N = n.size
        rdot = np.zeros(2 * N)
        color = np.zeros(N, dtype=np.int32)
        F = np.zeros(2 * N)
        face = np.zeros(N, dtype=np.int32)
        P = Params(v0, k, sigma, step_size, eta, color_factor, seed, mode, int(brute), threads)
        st = Stats()
        ids_p = None
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.uint32)
            ids_p = ids.ctypes.data_as(_up)
        if eta_inject is not None:
            eta_inject = np.ascontiguousarray(eta_inject, dtype=np.float64)
            self.L.t2do_inject_noise(self.ctx, _d(eta_inject))
        else:
            self.L.t2do_inject_noise(self.ctx, None)
        angle = np.zeros(N)
        self.L.t2do_set_angle_out(self.ctx, _d(angle))
        fault = self.L.t2do_step(self.ctx, C.byref(P), N, _d(uv), _i(n), _i(vid), _d(r3d), ids_p, step_index, _d(rdot),
                                 _i(color), _d(F), _i(face), C.byref(st))
        self.L.t2do_set_angle_out(self.ctx, None)
        if fault < 0:
            raise RuntimeError("oracle step failed: %d" % fault)
        self.last_stats = st.as_dict()
        return dict(uv=uv, n=n, vid=vid, r3d=r3d, rdot=rdot, color=color, F=F, face=face, fault=fault, angle=angle,
                    stats=st.as_dict())

    def observables(self, n, rdot):
        n = np.ascontiguousarray(n, dtype=np.int32)
        rdot = np.ascontiguousarray(rdot, dtype=np.float64)
        out = np.zeros(2)
        self.L.t2do_observables(n.size, _i(n), _d(rdot), _d(out))
        return out
