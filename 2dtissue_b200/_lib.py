"""Loader + ctypes signatures for lib2dtissue_b200.so (include/t2d.h).  There is no CPU fallback: if the
shared library is missing, or no sm_100 GPU is present when a context is created, this raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib2dtissue_b200.so")

OBS_LEN = 8
UNIQUE_ID_BYTES = 128


class Mesh(C.Structure):
    _fields_ = [("V", C.c_int32), ("F", C.c_int32), ("uv", C.POINTER(C.c_double)), ("x3d", C.POINTER(C.c_double)),
                ("faces", C.POINTER(C.c_int32))]


class Table(C.Structure):
    _fields_ = [("V", C.c_int32), ("kind", C.c_int32), ("data", C.c_void_p)]


class TableCSRStruct(C.Structure):   # t2d_table_csr
    _fields_ = [("nnz", C.c_int64), ("start", C.POINTER(C.c_int32)), ("col", C.POINTER(C.c_int32)), ("val", C.c_void_p),
                ("radius", C.c_double)]


class Params(C.Structure):
    _fields_ = [("v0", C.c_double), ("k", C.c_double), ("sigma", C.c_double), ("step_size", C.c_double),
                ("eta", C.c_double), ("color_factor", C.c_double), ("seed", C.c_uint64), ("neigh_mode", C.c_int32),
                ("precision", C.c_int32), ("capacity", C.c_int32), ("lift_mode", C.c_int32)]


class Counters(C.Structure):
    _names = ("steps", "kernel_launches", "pairs_in_range", "ties_cutoff", "ties_trunc", "wraps", "wrap_cap_hits",
              "order_fallbacks", "trig_fallbacks", "locate_fallbacks", "max_row", "cell_fallbacks", "buckets")
    _fields_ = [(k, C.c_int64) for k in _names] + [("reserved", C.c_int64 * 3)]

    def as_dict(self):
        return {k: getattr(self, k) for k in self._names}


# every symbol include/t2d.h declares (tests/test_abi.py checks the built library exports all of them)
SYMBOLS = [
    "t2d_create", "t2d_destroy", "t2d_last_error", "t2d_version", "t2d_set_particles", "t2d_set_state", "t2d_download",
    "t2d_seed_particles", "t2d_export_begin", "t2d_export_wait", "t2d_particle_count", "t2d_step", "t2d_step_host", "t2d_step_host_uv", "t2d_observables", "t2d_get_counters", "t2d_reset_counters", "t2d_set_tie_log",
    "t2d_get_step", "t2d_set_step", "t2d_set_params", "t2d_get_r3d", "t2d_tiling", "t2d_angles_to_unit_vectors",
    "t2d_forces", "t2d_build_hop_table", "t2d_last_step_ms", "t2d_profile_step", "t2d_pinned_alloc", "t2d_pinned_free",
    "t2d_comm_unique_id", "t2d_comm_init", "t2d_comm_init_local", "t2d_step_local", "t2d_comm_destroy", "t2d_owned_count",
    "t2d_download_ids",
]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing — build it with `make -C 2dtissue_b200/csrc` (or __graft_entry__.build()); "
                           "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    dp, ip, up = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
    vp = C.c_void_p
    L.t2d_create.argtypes = [C.POINTER(Mesh), C.POINTER(Table), C.POINTER(Params), C.c_int, C.POINTER(vp)]
    L.t2d_destroy.argtypes = [vp]
    L.t2d_destroy.restype = None
    L.t2d_last_error.argtypes = [vp]
    L.t2d_last_error.restype = C.c_char_p
    L.t2d_set_particles.argtypes = [vp, C.c_int32, dp, ip, up]
    L.t2d_set_state.argtypes = [vp, C.c_int32, dp, ip, ip, dp, up]
    L.t2d_download.argtypes = [vp, dp, ip, ip, dp, dp, ip, ip]
    L.t2d_particle_count.argtypes = [vp]
    L.t2d_step.argtypes = [vp, C.c_int32]
    L.t2d_step_host.argtypes = [vp, C.c_int32, dp, ip, ip, dp, dp, ip]
    L.t2d_step_host_uv.argtypes = [vp, C.c_int32, dp, ip, ip, dp, dp, ip]
    L.t2d_observables.argtypes = [vp, dp]
    L.t2d_get_counters.argtypes = [vp, C.POINTER(Counters)]
    L.t2d_reset_counters.argtypes = [vp]
    L.t2d_set_tie_log.argtypes = [vp, C.c_int]
    L.t2d_export_begin.argtypes = [vp, C.POINTER(C.c_int32)]
    pdp, pip = C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.POINTER(C.c_int32))
    L.t2d_export_wait.argtypes = [vp, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int64), pdp, pip, pip, pdp, pdp, pip]
    L.t2d_seed_particles.argtypes = [vp, C.c_int32, C.c_uint64, C.c_int32, C.c_uint32]
    L.t2d_get_step.argtypes = [vp]
    L.t2d_get_step.restype = C.c_int64
    L.t2d_set_step.argtypes = [vp, C.c_int64]
    L.t2d_set_params.argtypes = [vp, C.POINTER(Params)]
    L.t2d_get_r3d.argtypes = [vp, C.c_int32, dp, dp, ip, ip]
    L.t2d_tiling.argtypes = [vp, C.c_int32, dp, dp, ip]
    L.t2d_angles_to_unit_vectors.argtypes = [vp, C.c_int32, ip, dp]
    L.t2d_forces.argtypes = [vp, dp, ip, ip]
    L.t2d_build_hop_table.argtypes = [vp, C.POINTER(C.c_ubyte)]
    L.t2d_last_step_ms.argtypes = [vp]
    L.t2d_last_step_ms.restype = C.c_double
    L.t2d_profile_step.argtypes = [vp, C.POINTER(C.c_char_p), dp, C.c_int]
    L.t2d_pinned_alloc.argtypes = [C.c_size_t]
    L.t2d_pinned_alloc.restype = vp
    L.t2d_pinned_free.argtypes = [vp]
    L.t2d_pinned_free.restype = None
    L.t2d_comm_unique_id.argtypes = [C.POINTER(C.c_ubyte)]
    L.t2d_comm_init.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_ubyte), dp]
    L.t2d_comm_destroy.argtypes = [vp]
    L.t2d_comm_init_local.argtypes = [C.POINTER(vp), C.c_int, dp]
    L.t2d_step_local.argtypes = [C.POINTER(vp), C.c_int, C.c_int32]
    L.t2d_owned_count.argtypes = [vp]
    L.t2d_owned_count.restype = C.c_int32
    L.t2d_download_ids.argtypes = [vp, up]
    _lib = L
    return L
