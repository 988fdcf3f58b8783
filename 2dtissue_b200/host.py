"""Python host side above the C ABI.  No arithmetic of the step happens here: every stage runs in
lib2dtissue_b200.so on the GPU; this file only marshals the reference's array layouts."""
import ctypes as C
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import _lib
from ._lib import Counters, Mesh, Params, Table, TableCSRStruct
from .table import TableCSR

TABLE_NONE, TABLE_DENSE_F64, TABLE_DENSE_F32, TABLE_DENSE_U8, TABLE_HOPS_FROM_MESH = 0, 1, 2, 3, 4
TABLE_CSR_F64, TABLE_CSR_F32, TABLE_CSR_U8 = 5, 6, 7
NEIGH_TABLE, NEIGH_EUCLID = 0, 1
PRECISION_FP64, PRECISION_FP32 = 0, 1
FAULT_LOST, FAULT_NONFINITE, FAULT_WRAP_CAP, FAULT_MIGRATION, FAULT_COMM_OVERFLOW = 1, 2, 4, 8, 16
LIFT_REFERENCE, LIFT_BARYCENTRIC = 0, 1   # T2D_LIFT_*

_dp, _ip, _up = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)


class T2DError(RuntimeError):
    pass


class LostParticlesError(T2DError):
    """Validation::error_lost_particles (/root/reference/src/simulation/Validation.cpp:66-72)."""


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


class Context:
    """One GPU context (t2d_ctx).  Arrays use the reference's layouts: uv = N x's then N y's, r3d = N x,N y,N z."""

    def __init__(self, chart, table=None, table_kind=None, v0=0.1, k=1.0, sigma=0.4166666666666667, step_size=0.001,
                 eta=0.0, color_factor=2.4, seed=0, neigh_mode=NEIGH_TABLE, precision=PRECISION_FP64, capacity=1024,
                 device=0, lift_mode=LIFT_REFERENCE):
        self.L = _lib.load()
        self._uv = np.ascontiguousarray(chart["uv"], dtype=np.float64)
        self._x3d = np.ascontiguousarray(chart["x3d"], dtype=np.float64)
        self._faces = np.ascontiguousarray(chart["faces"], dtype=np.int32)
        self.V, self.F = len(self._uv), len(self._faces)
        mesh = Mesh(self.V, self.F, _d(self._uv), _d(self._x3d), _i(self._faces))
        tab = Table(self.V, TABLE_NONE, None)
        self._table = None
        if table_kind == TABLE_HOPS_FROM_MESH:
            tab = Table(self.V, TABLE_HOPS_FROM_MESH, None)
        elif isinstance(table, TableCSR):   # thresholded rows (include/t2d.h t2d_table_csr)
            kind = {np.dtype(np.float64): TABLE_CSR_F64, np.dtype(np.float32): TABLE_CSR_F32, np.dtype(np.uint8): TABLE_CSR_U8}[table.val.dtype]
            self._table = table
            self._csr = TableCSRStruct(table.nnz, _i(table.start), _i(table.col), table.val.ctypes.data_as(C.c_void_p), table.radius)
            tab = Table(table.V, kind, C.cast(C.pointer(self._csr), C.c_void_p))
        elif table is not None:
            t = np.ascontiguousarray(table)
            kind = {np.dtype(np.float64): TABLE_DENSE_F64, np.dtype(np.float32): TABLE_DENSE_F32,
                    np.dtype(np.uint8): TABLE_DENSE_U8}[t.dtype]
            self._table = t
            tab = Table(t.shape[0], kind, t.ctypes.data_as(C.c_void_p))
        self.params = Params(v0, k, sigma, step_size, eta, color_factor, seed, neigh_mode, precision, capacity, lift_mode)
        h = C.c_void_p()
        rc = self.L.t2d_create(C.byref(mesh), C.byref(tab), C.byref(self.params), device, C.byref(h))
        if rc != 0:
            raise T2DError("t2d_create failed: %s" % self.L.t2d_last_error(None).decode())
        self.h = h
        self._table = None  # the library keeps its own copy

    def close(self):
        if getattr(self, "h", None):
            self.L.t2d_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        if rc < 0:
            raise T2DError("%s failed: %s" % (what, self.L.t2d_last_error(self.h).decode()))
        return rc

    @property
    def N(self):
        return self.L.t2d_particle_count(self.h)

    def set_params(self, **kw):
        for k_, v in kw.items():
            setattr(self.params, k_, v)
        self._chk(self.L.t2d_set_params(self.h, C.byref(self.params)), "t2d_set_params")

    def set_particles(self, uv, heading, ids=None):
        uv = np.ascontiguousarray(uv, dtype=np.float64)
        heading = np.ascontiguousarray(heading, dtype=np.int32)
        idp = None
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.uint32)
            idp = ids.ctypes.data_as(_up)
        self._chk(self.L.t2d_set_particles(self.h, heading.size, _d(uv), _i(heading), idp), "t2d_set_particles")

    def seed_on_device(self, N, seed=1234, mode=0, first_id=0):
        """t2d_seed_particles: Philox-seeded particles + initial projection entirely on the GPU (mode 0: uniform in the chart,
        mode 1: face centres like CellHelper::init_particle_position)."""
        self._chk(self.L.t2d_seed_particles(self.h, int(N), int(seed), int(mode), int(first_id)), "t2d_seed_particles")

    def set_state(self, uv, heading, vid, r3d, ids=None):
        uv = np.ascontiguousarray(uv, dtype=np.float64)
        heading = np.ascontiguousarray(heading, dtype=np.int32)
        vid = np.ascontiguousarray(vid, dtype=np.int32)
        r3d = np.ascontiguousarray(r3d, dtype=np.float64)
        idp = None
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.uint32)
            idp = ids.ctypes.data_as(_up)
        self._chk(self.L.t2d_set_state(self.h, heading.size, _d(uv), _i(heading), _i(vid), _d(r3d), idp), "t2d_set_state")

    # ---- multi-GPU slabs (include/t2d.h "multi-GPU") ----
    def comm_init(self, rank, world, unique_id, cuts):
        """NCCL transport, one process per GPU; collective.  unique_id: 128 bytes from comm_unique_id() of rank 0."""
        cuts = np.ascontiguousarray(cuts, dtype=np.float64)
        uid = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        self._chk(self.L.t2d_comm_init(self.h, rank, world, uid, _d(cuts)), "t2d_comm_init")

    def owned_count(self):
        return self._chk(self.L.t2d_owned_count(self.h), "t2d_owned_count")

    def download_ids(self):
        ids = np.zeros(self.owned_count(), dtype=np.uint32)
        self._chk(self.L.t2d_download_ids(self.h, ids.ctypes.data_as(_up)), "t2d_download_ids")
        return ids

    def download_into(self, bufs):
        """Download into caller-provided (e.g. pinned) arrays sized for the context's capacity; returns N owned.
        bufs: dict with any of uv, n, vid, r3d, rdot, color, face, ids.  Layout = the reference's, with stride N."""
        N = self.owned_count()
        g = bufs.get
        self._chk(self.L.t2d_download(self.h, _d(g("uv")), _i(g("n")), _i(g("vid")), _d(g("r3d")), _d(g("rdot")),
                                      _i(g("color")), _i(g("face"))), "t2d_download")
        if g("ids") is not None:
            self._chk(self.L.t2d_download_ids(self.h, g("ids").ctypes.data_as(_up)), "t2d_download_ids")
        return N

    def set_state_raw(self, N, uv, heading, vid, r3d, ids=None):
        """set_state without any host-side copy: arrays may be larger than needed (first 2N / N / 3N entries are used)."""
        self._chk(self.L.t2d_set_state(self.h, N, _d(uv), _i(heading), _i(vid), _d(r3d),
                                       ids.ctypes.data_as(_up) if ids is not None else None), "t2d_set_state")

    def download(self, fields=("uv", "n", "vid", "r3d", "rdot", "color", "face")):
        N = self.owned_count()
        out = {}
        if "uv" in fields:
            out["uv"] = np.zeros(2 * N)
        if "n" in fields:
            out["n"] = np.zeros(N, dtype=np.int32)
        if "vid" in fields:
            out["vid"] = np.zeros(N, dtype=np.int32)
        if "r3d" in fields:
            out["r3d"] = np.zeros(3 * N)
        if "rdot" in fields:
            out["rdot"] = np.zeros(2 * N)
        if "color" in fields:
            out["color"] = np.zeros(N, dtype=np.int32)
        if "face" in fields:
            out["face"] = np.zeros(N, dtype=np.int32)
        self._chk(self.L.t2d_download(self.h, _d(out.get("uv")), _i(out.get("n")), _i(out.get("vid")), _d(out.get("r3d")),
                                      _d(out.get("rdot")), _i(out.get("color")), _i(out.get("face"))), "t2d_download")
        return out

    def step(self, nsteps=1):
        return self._chk(self.L.t2d_step(self.h, nsteps), "t2d_step")

    def export_begin(self):
        """Start an asynchronous snapshot + device->pinned-host copy of the resident state on a side stream; returns a slot."""
        slot = C.c_int32(0)
        self._chk(self.L.t2d_export_begin(self.h, C.byref(slot)), "t2d_export_begin")
        return slot.value

    def export_wait(self, slot):
        """Block until the slot's copy has landed; numpy views into the pinned ring (valid until the slot is reused)."""
        N, step = C.c_int32(0), C.c_int64(0)
        uv, r3d, rdot = C.POINTER(C.c_double)(), C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
        n, vid, col = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
        self._chk(self.L.t2d_export_wait(self.h, slot, C.byref(N), C.byref(step), C.byref(uv), C.byref(n), C.byref(vid), C.byref(r3d),
                                         C.byref(rdot), C.byref(col)), "t2d_export_wait")
        m = N.value
        view = lambda p, k: np.ctypeslib.as_array(p, shape=(k * m,)) if m else np.zeros(0)
        return dict(step=step.value, uv=view(uv, 2), n=view(n, 1), vid=view(vid, 1), r3d=view(r3d, 3), rdot=view(rdot, 2), color=view(col, 1))

    def step_host(self, uv, heading, vid, r3d, rdot, color, reproject=False):
        """In/out numpy arrays in the reference's layouts (the literal perform_particle_simulation drop-in).
        reproject=True: only uv and heading are uploaded, the device re-projects uv (vid, r3d are outputs only)."""
        fn = self.L.t2d_step_host_uv if reproject else self.L.t2d_step_host
        return self._chk(fn(self.h, heading.size, _d(uv), _i(heading), _i(vid), _d(r3d), _d(rdot), _i(color)), "t2d_step_host")

    def observables(self):
        out = np.zeros(_lib.OBS_LEN)
        self._chk(self.L.t2d_observables(self.h, _d(out)), "t2d_observables")
        return dict(phi=out[0], mean_speed=out[1], sum_cos=out[2], sum_sin=out[3], sum_speed=out[4], count=out[5],
                    lost=out[6], nonfinite=out[7])

    def counters(self):
        c = Counters()
        self._chk(self.L.t2d_get_counters(self.h, C.byref(c)), "t2d_get_counters")
        return c.as_dict()

    def reset_counters(self):
        self._chk(self.L.t2d_reset_counters(self.h), "t2d_reset_counters")

    def set_tie_log(self, on=True):
        """fp32 fast path: count candidates within 8 ulps of a squared cutoff in counters()["ties_cutoff"] (off by default)."""
        self._chk(self.L.t2d_set_tie_log(self.h, 1 if on else 0), "t2d_set_tie_log")

    @property
    def step_index(self):
        return self.L.t2d_get_step(self.h)

    @step_index.setter
    def step_index(self, v):
        self.L.t2d_set_step(self.h, int(v))

    @property
    def last_step_ms(self):
        return self.L.t2d_last_step_ms(self.h)

    # ---- single stages ----
    def get_r3d(self, uv):
        uv = np.ascontiguousarray(uv, dtype=np.float64)
        N = uv.size // 2
        r3d, vid, face = np.zeros(3 * N), np.zeros(N, dtype=np.int32), np.zeros(N, dtype=np.int32)
        self._chk(self.L.t2d_get_r3d(self.h, N, _d(uv), _d(r3d), _i(vid), _i(face)), "t2d_get_r3d")
        return r3d, vid, face

    def tiling(self, uv_old, uv, heading):
        uv_old = np.array(uv_old, dtype=np.float64).copy()
        uv = np.array(uv, dtype=np.float64).copy()
        heading = np.array(heading, dtype=np.int32).copy()
        fault = self._chk(self.L.t2d_tiling(self.h, heading.size, _d(uv_old), _d(uv), _i(heading)), "t2d_tiling")
        return uv_old, uv, heading, fault

    def angles_to_unit_vectors(self, heading):
        heading = np.ascontiguousarray(heading, dtype=np.int32)
        out = np.zeros(2 * heading.size)
        self._chk(self.L.t2d_angles_to_unit_vectors(self.h, heading.size, _i(heading), _d(out)), "t2d_angles_to_unit_vectors")
        return out

    def forces(self):
        N = self.N
        F, nh, col = np.zeros(2 * N), np.zeros(N, dtype=np.int32), np.zeros(N, dtype=np.int32)
        self._chk(self.L.t2d_forces(self.h, _d(F), _i(nh), _i(col)), "t2d_forces")
        return F, nh, col

    def build_hop_table(self):
        out = np.zeros((self.V, self.V), dtype=np.uint8)
        self._chk(self.L.t2d_build_hop_table(self.h, out.ctypes.data_as(C.POINTER(C.c_ubyte))), "t2d_build_hop_table")
        return out

    def profile_step(self):
        names = (C.c_char_p * 16)()
        ms = np.zeros(16)
        n = self._chk(self.L.t2d_profile_step(self.h, names, _d(ms), 16), "t2d_profile_step")
        return [(names[i].decode(), float(ms[i])) for i in range(n)]


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it and broadcasts it, e.g. with torch.distributed)."""
    L = _lib.load()
    buf = (C.c_ubyte * 128)()
    if L.t2d_comm_unique_id(buf) != 0:
        raise T2DError("t2d_comm_unique_id failed: %s" % L.t2d_last_error(None).decode())
    return bytes(buf)


def slab_cuts(x, world):
    """world-1 interior slab boundaries along x with equal particle counts (cuts chosen so that
    particle counts are equal at t = 0).  A cut is placed midway between two consecutive sorted x values."""
    x = np.sort(np.asarray(x, dtype=np.float64))
    cuts = []
    for r in range(1, world):
        k = (len(x) * r) // world
        cuts.append(0.5 * (x[k - 1] + x[k]) if 0 < k < len(x) else (x[-1] if len(x) else 0.0))
    return np.array(cuts, dtype=np.float64)


def slab_of(x, cuts):
    """Owner rank of every particle: slab r owns x in [cuts[r-1], cuts[r])."""
    return np.searchsorted(np.asarray(cuts, dtype=np.float64), np.asarray(x, dtype=np.float64), side="right").astype(np.int32)


def route_after_step(x, cuts, rank, halo):
    """Host-side restatement of `slab_classify` (csrc/kernels.cuh): where the particles OWNED by `rank` go after a step,
    from their new 3-D x.  Returns boolean masks over x:
        stay                      keeps its owner
        to_left / to_right        migrates through the message to slab-1 / slab+1
        keep_halo                 migrant that stays behind as a halo copy (still within `halo` of the cut it crossed)
        halo_left / halo_right    stays, and a halo copy goes to slab-1 / slab+1
        far                       goes into the far message every rank receives (landed beyond the adjacent slab, or
                                  inside it but within `halo` of its other cut)
    and dest = the slab that owns the new x.  Slabs must be at least 4 halo wide (t2d_comm_init checks it)."""
    x = np.asarray(x, dtype=np.float64)
    cuts = np.asarray(cuts, dtype=np.float64)
    world = len(cuts) + 1
    cut = lambda k: -np.inf if k < 0 else (np.inf if k >= world - 1 else cuts[k])
    lo, hi, lo2, hi2 = cut(rank - 1), cut(rank), cut(rank - 2), cut(rank + 1)
    far = (x < lo2 + halo) | (x >= hi2 - halo)
    left = ~far & (x < lo)
    right = ~far & (x >= hi)
    keep_halo = (left & (x >= lo - halo)) | (right & (x < hi + halo))
    stay = ~far & ~left & ~right
    return dict(stay=stay, to_left=left, to_right=right, keep_halo=keep_halo, far=far,
                halo_left=stay & (rank > 0) & (x < lo + halo), halo_right=stay & (rank < world - 1) & (x >= hi - halo),
                dest=slab_of(x, cuts))


def far_receive(x, dest, cuts, rank, halo):
    """What `rank` does with the far records of the other ranks (k_comm_unpack_far): (adopt, halo_copy) masks."""
    x = np.asarray(x, dtype=np.float64)
    cuts = np.asarray(cuts, dtype=np.float64)
    world = len(cuts) + 1
    lo = -np.inf if rank == 0 else cuts[rank - 1]
    hi = np.inf if rank == world - 1 else cuts[rank]
    mine = np.asarray(dest) == rank
    return mine, ~mine & (x >= lo - halo) & (x < hi + halo)


def partition_by_slab(state, cuts, rank):
    """The sub-state (with global ids) that `rank` owns.  state: dict(uv, n, vid, r3d) in the reference's layouts."""
    N = state["n"].size
    x = state["r3d"][:N]
    sel = np.nonzero(slab_of(x, cuts) == rank)[0]
    col = lambda a, k: np.concatenate([a[j * N:(j + 1) * N][sel] for j in range(k)])
    return dict(uv=col(state["uv"], 2), n=state["n"][sel], vid=state["vid"][sel], r3d=col(state["r3d"], 3),
                ids=sel.astype(np.uint32))


def merge_by_id(parts, N):
    """Inverse of partition_by_slab for downloaded per-rank results: dict of arrays in global-id order."""
    out = {}
    for p in parts:
        ids = p["ids"].astype(np.int64)
        m = ids.size
        for k, a in p.items():
            if k == "ids":
                continue
            cols = a.size // m if m else 1
            if k not in out:
                cols = {"uv": 2, "rdot": 2, "r3d": 3}.get(k, 1)
                out[k] = np.zeros(cols * N, dtype=a.dtype)
            cols = out[k].size // N
            for j in range(cols):
                out[k][j * N + ids] = a[j * m:(j + 1) * m]
    return out


class LocalSlabGroup:
    """`world` contexts driven by one host thread (t2d_comm_init_local / t2d_step_local): P logical slabs on one GPU
    (parity tests: P slabs == 1 GPU) or one context per GPU of a box."""

    def __init__(self, chart, world, cuts, devices=None, **ctx_kw):
        self.world, self.cuts = world, np.ascontiguousarray(cuts, dtype=np.float64)
        devices = devices or [0] * world
        self.ctxs = [Context(chart, device=devices[r], **ctx_kw) for r in range(world)]
        self.L = self.ctxs[0].L
        self._arr = (C.c_void_p * world)(*[c.h for c in self.ctxs])
        if self.L.t2d_comm_init_local(self._arr, world, _d(self.cuts)) != 0:
            raise T2DError("t2d_comm_init_local failed: %s" % self.L.t2d_last_error(self.ctxs[0].h).decode())
        self.N = 0

    def set_state(self, state):
        self.N = state["n"].size
        for r, c in enumerate(self.ctxs):
            p = partition_by_slab(state, self.cuts, r)
            c.set_state(p["uv"], p["n"], p["vid"], p["r3d"], ids=p["ids"])

    def step(self, nsteps=1):
        rc = self.L.t2d_step_local(self._arr, self.world, nsteps)
        if rc < 0:
            raise T2DError("t2d_step_local failed: %s" % self.L.t2d_last_error(self.ctxs[0].h).decode())
        return rc

    def download(self, fields=("uv", "n", "vid", "r3d", "rdot", "color", "face")):
        parts = []
        for c in self.ctxs:
            d = c.download(fields)
            d["ids"] = c.download_ids()
            parts.append(d)
        return merge_by_id(parts, self.N), [p["ids"].size for p in parts]

    def close(self):
        for c in self.ctxs:
            c.close()


def seed_particles(N, seed=1234, first_id=0):
    """Synthetic inputs of SURVEY.md §8d: u,v ~ U(0,1), heading ~ U{0..359}, from a seeded counter-based stream."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    if first_id:
        rng.bit_generator.advance(3 * first_id)
    u = rng.random(N)
    v = rng.random(N)
    n = rng.integers(0, 360, N).astype(np.int32)
    return np.concatenate([u, v]), n


# ---- mirror of the reference's Struct.h:9-29 -----------------------------------------------------------
@dataclass
class Particle:
    x_UV: float
    y_UV: float
    x_velocity_UV: float
    y_velocity_UV: float
    alignment_UV: float
    x_3D: float
    y_3D: float
    z_3D: float
    neighbor_count: int


@dataclass
class System:
    order_parameter: float
    particles: List[Particle] = field(default_factory=list)


class Tissue2D:
    """Mirror of `_2DTissue` (2DTissue.h:33-60) for the hot path: same ctor argument names/defaults, same
    start()/update()/is_finished()/get_order_parameter() protocol.  `chart` replaces `mesh_path`: the chart is
    produced by the reference's host-side MeshCartographyLib setup (2DTissue.cpp:85-107) and loaded from a
    .t2dchart file; the distance table is built on the GPU from the chart's edge graph unless `table` is given.
    """

    def __init__(self, save_data=False, particle_innenleben=False, free_boundary=False, chart=None, particle_count=1000,
                 step_count=1, v0=0.1, use_kafka=False, k=1.0, k_next=10, v0_next=0.1, sigma=0.4166666666666667, mu=1,
                 r_adh=1, k_adh=0.75, step_size=0.001, map_cache_count=30, *, table=None, neigh_mode=NEIGH_TABLE,
                 precision=PRECISION_FP64, eta=0.0, seed=0, device=0, build_particles=False):
        if use_kafka:
            raise T2DError("Kafka output is out of scope (SURVEY.md §2 row 12)")
        self.particle_count, self.step_count = particle_count, step_count
        self.current_step, self.finished = 0, False
        self.v_order = np.zeros(step_count)
        self.save_data, self.build_particles = save_data, build_particles
        kind = None
        if neigh_mode == NEIGH_TABLE and table is None:
            kind = TABLE_HOPS_FROM_MESH
        self.ctx = Context(chart, table=table, table_kind=kind, v0=v0, k=k, sigma=sigma, step_size=step_size, eta=eta,
                           seed=seed, neigh_mode=neigh_mode, precision=precision, capacity=max(1, particle_count),
                           device=device)

    def start(self, uv=None, heading=None):
        """_2DTissue::start (2DTissue.cpp:117-134).  The reference seeds from std::random_device; here the
        caller passes the state (or a seeded synthetic one is used) and the GPU does the initial get_r3d()."""
        if uv is None:
            uv, heading = seed_particles(self.particle_count)
        self.ctx.set_particles(uv, heading)

    def update(self):
        """_2DTissue::update (2DTissue.cpp:136-206) with perform_particle_simulation() on the GPU."""
        fault = self.ctx.step(1)
        if fault & FAULT_LOST:
            raise LostParticlesError("We lost particles after getting the original UV mesh coord")
        if fault & FAULT_NONFINITE:
            raise SystemExit(1)  # Validation::error_invalid_values -> std::exit(1)
        if fault & FAULT_WRAP_CAP:
            raise T2DError("seam re-entry did not terminate")
        obs = self.ctx.observables()
        self.v_order[self.current_step] = obs["phi"]
        system = System(order_parameter=obs["phi"])
        if self.build_particles:
            s = self.ctx.download(("uv", "n", "r3d", "rdot", "color"))
            N = self.particle_count
            system.particles = [Particle(s["uv"][i], s["uv"][N + i], s["rdot"][i], s["rdot"][N + i], float(s["n"][i]),
                                         s["r3d"][i], s["r3d"][N + i], s["r3d"][2 * N + i], int(s["color"][i]))
                                for i in range(N)]
        self.current_step += 1
        if self.current_step >= self.step_count:
            self.finished = True
        return system

    def is_finished(self):
        return self.finished

    def get_order_parameter(self):
        return self.v_order
