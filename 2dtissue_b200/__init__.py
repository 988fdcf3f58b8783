"""2dtissue_b200 — B200-native (sm_100a) implementation of 2DTissue's per-timestep particle update.

Host-side mirror of the reference's driver surface for the hot path only:

    Context   thin object wrapper over the C ABI (include/t2d.h, lib2dtissue_b200.so)
    Tissue2D  mirrors `_2DTissue` (/root/reference/src/simulation/2DTissue.h:33-60): start() / update() /
              is_finished() / get_order_parameter(), with perform_particle_simulation() on the GPU

The directory name starts with a digit, so import it with
    t2d = importlib.import_module("2dtissue_b200")
"""
from . import chart  # noqa: F401
from .chart import load_chart, refine_chart, save_chart  # noqa: F401
from .table import TableCSR  # noqa: F401
from .host import (FAULT_LOST, FAULT_NONFINITE, FAULT_WRAP_CAP, NEIGH_EUCLID, NEIGH_TABLE, PRECISION_FP32, PRECISION_FP64, TABLE_DENSE_F32, TABLE_DENSE_F64,  # noqa: F401
                   TABLE_DENSE_U8, TABLE_HOPS_FROM_MESH, TABLE_NONE, Context, LostParticlesError, Particle, System,
                   T2DError, Tissue2D, seed_particles, FAULT_MIGRATION, FAULT_COMM_OVERFLOW, LocalSlabGroup,
                   comm_unique_id, merge_by_id, partition_by_slab, slab_cuts, slab_of, LIFT_REFERENCE, LIFT_BARYCENTRIC,
                   TABLE_CSR_F64, TABLE_CSR_F32, TABLE_CSR_U8)
