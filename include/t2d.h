/* include/t2d.h — C ABI of lib2dtissue_b200.so: 2DTissue's per-timestep particle update on B200 (sm_100a).
 *
 * Drop-in boundary.  The reference has no FFI layer; its practical seam is
 *     _2DTissue::perform_particle_simulation()            /root/reference/src/simulation/2DTissue.cpp:222-252
 * whose body is exactly the hot path and whose only caller is _2DTissue::update() (:136-206).  A reference
 * maintainer keeps `update()`'s bookkeeping and replaces that body by t2d_step()/t2d_step_host();
 * INTEGRATION.md shows the binding.  Plain pointers and sizes only — no C++/torch types cross this line.
 *
 * Host array layouts are the reference's Eigen column-major ones (2DTissue.h:89-106):
 *     uv   double[2N]  = N x's then N y's          (r_UV, r_UV_old, r_dot)
 *     r3d  double[3N]  = N x, N y, N z             (r_3D)
 *     heading, vid, color, face  int32[N]          (n, vertices_3D_active, particles_color)
 * Particle i of the host arrays is the particle with global id ids[i] (default: i); the library keeps
 * its own device order and always uploads/downloads in the caller's order.
 *
 * Status codes: 0 ok; < 0 argument / CUDA / NCCL error (t2d_last_error has the text); > 0 simulation fault
 * bitmask, the analogue of the reference's exceptions:
 *     T2D_FAULT_LOST      Validation::error_lost_particles throws std::runtime_error   (Validation.cpp:66-72)
 *     T2D_FAULT_NONFINITE Validation::error_invalid_values calls std::exit(1)          (Validation.cpp:40-46)
 *     T2D_FAULT_WRAP_CAP  the seam re-entry loop (EuclideanTiling.cpp:41-68) did not terminate in 4096 rounds
 */
#ifndef T2D_H
#define T2D_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct t2d_ctx t2d_ctx;

/* UV chart of the cut-open mesh == what _2DTissue's ctor leaves for the stepping loop (2DTissue.cpp:85-107):
 * vertice_UV (z dropped), vertice_3D, face_UV.  Coordinates are float32 values widened to double
 * (pmp::Scalar = float); the library stores them as float32 when that is lossless. */
typedef struct {
    int32_t V, F;
    const double* uv;     /* [V][2] row-major */
    const double* x3d;    /* [V][3] row-major */
    const int32_t* faces; /* [F][3] vertex ids */
} t2d_mesh;

/* The vertex-distance table of stage 2 "in its stored precision"
 * (CachedGeodesicDistanceHelper::get_mesh_distance_matrix, MeshCartographyLib CachedGeodesicDistanceHelper.cpp:29-40). */
enum {
    T2D_TABLE_NONE = 0,           /* Euclidean mode needs no table */
    T2D_TABLE_DENSE_F64 = 1,      /* double[V][V], what the reference holds in memory */
    T2D_TABLE_DENSE_F32 = 2,      /* float[V][V]; entries are widened to double before every comparison */
    T2D_TABLE_DENSE_U8 = 3,       /* uint8[V][V] hop counts (255 = farther than 254 hops) */
    T2D_TABLE_HOPS_FROM_MESH = 4, /* build the hop-count table on the GPU from the mesh's edge graph
                                     (replaces DijkstraDistanceHelper.cpp:27-111); data is ignored.  Meshes whose dense table
                                     would exceed 1 GB get the thresholded CSR built directly (no V x V array anywhere) */
    T2D_TABLE_CSR_F64 = 5,        /* thresholded table in CSR form, data -> t2d_table_csr, values double */
    T2D_TABLE_CSR_F32 = 6,        /* the same with float values (widened to double before every comparison) */
    T2D_TABLE_CSR_U8 = 7          /* the same with uint8 hop counts */
};
/* A vertex-distance table that only lists the pairs that can ever interact: row v = the vertices u with
 * D(v, u) <= radius, ascending u, the diagonal included, already symmetrised by min(D(v,u), D(u,v)) as Locomotion.cpp:110
 * does.  `radius` is the distance up to which the rows are complete; t2d_create / t2d_set_params refuse an interaction
 * radius max(2 sigma, color_factor sigma) beyond it.  What the dense kinds cost in memory (V = 75 k: 45 GB as double, the
 * reference's own format) this kind does not: it is what stage 2 actually reads. */
typedef struct {
    int64_t nnz;
    const int32_t* start; /* [V + 1] */
    const int32_t* col;   /* [nnz] ascending inside a row */
    const void* val;      /* [nnz] double / float / uint8 by kind */
    double radius;
} t2d_table_csr;
typedef struct {
    int32_t V;
    int32_t kind;
    const void* data;
} t2d_table;

enum { T2D_NEIGH_TABLE = 0, T2D_NEIGH_EUCLID = 1 };
enum { T2D_PRECISION_FP64 = 0, T2D_PRECISION_FP32 = 1 };
/* UV -> 3-D lift.  REFERENCE: weights = normalised UV distances to the face's corners (CellHelper.cpp:133-146 — not
 * barycentric: it pulls every point towards the middle of its face).  BARYCENTRIC: the true barycentric coordinates,
 * i.e. the piecewise-linear chart the code comments describe; validated against the oracle, which
 * implements the same option, and statistically — there is no reference output to compare with. */
enum { T2D_LIFT_REFERENCE = 0, T2D_LIFT_BARYCENTRIC = 1 };
enum {
    T2D_FAULT_LOST = 1, T2D_FAULT_NONFINITE = 2, T2D_FAULT_WRAP_CAP = 4,
    T2D_FAULT_MIGRATION = 8,      /* reserved (particles that land beyond an adjacent halo strip travel through the far channel) */
    T2D_FAULT_COMM_OVERFLOW = 16  /* multi-GPU: a halo/migration message or the context's capacity overflowed */
};

/* _2DTissue ctor arguments that reach the step (2DTissue.h:37-54) + the extensions of SURVEY.md App. A */
typedef struct {
    double v0;            /* self-propulsion speed (CLI --step-time is wired here, main.cpp:72-82) */
    double k;             /* repulsion strength */
    double sigma;         /* particle radius; interaction cutoff is 2*sigma */
    double step_size;     /* Euler dt, 0.001 in the reference */
    double eta;           /* Vicsek noise amplitude in turns (eta_i = eta*360*(u-0.5) deg); 0 = reference */
    double color_factor;  /* 2.4: count_particle_neighbors counts 0 != d <= color_factor*sigma; 0 disables */
    uint64_t seed;        /* Philox4x32-10 key; counter = (step, global particle id) */
    int32_t neigh_mode;   /* T2D_NEIGH_* */
    int32_t precision;    /* T2D_PRECISION_* */
    int32_t capacity;     /* max particles resident on this context (owned + halo) */
    int32_t lift_mode;    /* T2D_LIFT_*: 0 = the reference's distance-weighted lift (zero-initialised callers get it) */
} t2d_params;

/* indices into the array filled by t2d_observables */
enum {
    T2D_OBS_PHI = 0,        /* polar order |sum_i n_hat_i| / N   (SURVEY.md §8 a11 — the reference's own value is UB) */
    T2D_OBS_MEAN_SPEED = 1, /* <|r_dot|> */
    T2D_OBS_SUM_COS = 2,    /* partial sums so that slabs can be combined */
    T2D_OBS_SUM_SIN = 3,
    T2D_OBS_SUM_SPEED = 4,
    T2D_OBS_COUNT = 5,
    T2D_OBS_LOST = 6,
    T2D_OBS_NONFINITE = 7,
    T2D_OBS_LEN = 8
};

/* diagnostics accumulated since t2d_create / t2d_reset_counters */
typedef struct {
    int64_t steps;
    int64_t kernel_launches;   /* kernels of this library launched so far */
    int64_t pairs_in_range;    /* sum_i |{j != i : d_ij < 2 sigma}| */
    int64_t ties_cutoff;       /* fp64: pairs with d_ij == 2 sigma exactly; fp32 fast path: candidates within 8 ulps of a squared
                                  cutoff, counted only while the tie log is on (t2d_set_tie_log) */
    int64_t ties_trunc;        /* headings whose mean angle is within 1e-9 deg of an integer */
    int64_t wraps;             /* seam re-entries */
    int64_t wrap_cap_hits;
    int64_t order_fallbacks;   /* rows too long for the exact ascending-id accumulation (summed unordered) */
    int64_t trig_fallbacks;    /* headings outside the host-built cos/sin table */
    int64_t locate_fallbacks;  /* point locations that scanned all faces */
    int64_t max_row;           /* longest neighbour row seen */
    int64_t cell_fallbacks;    /* particles found outside the static 3-D cell index (kept exact via the overflow bucket) */
    int64_t buckets;           /* number of buckets of the counting sort (mesh vertices / surface cells + 1) */
    int64_t reserved[3];
} t2d_counters;

/* ---- lifetime -------------------------------------------------------------------------------------- */
/* replaces the device-relevant part of _2DTissue::_2DTissue (2DTissue.cpp:20-111): uploads chart, table, params */
int t2d_create(const t2d_mesh* mesh, const t2d_table* table, const t2d_params* params, int device, t2d_ctx** out);
void t2d_destroy(t2d_ctx* ctx);
const char* t2d_last_error(const t2d_ctx* ctx); /* ctx may be NULL: error of the last failed t2d_create */
int t2d_version(void);

/* ---- state ----------------------------------------------------------------------------------------- */
/* replaces _2DTissue::start after init_particle_position (2DTissue.cpp:117-134): uploads r_UV and n and
 * runs the initial projection (CellHelper::get_r3d) to obtain r_3D / vertices_3D_active.  ids may be NULL. */
int t2d_set_particles(t2d_ctx* ctx, int32_t N, const double* uv, const int32_t* heading, const uint32_t* ids);
/* Device-side seeding + initial projection, no host arrays (replaces CellHelper::init_particle_position, CellHelper.cpp:43-67).
 * Particle i gets global id first_id + i and draws from the Philox4x32-10 stream keyed by (seed, draw, id):
 * mode 0: u, v ~ U(0, 1), heading ~ U{0..359} (the synthetic inputs of SURVEY.md 8d);
 * mode 1: the reference's scheme — the gravity centre of a random face and a random heading. */
int t2d_seed_particles(t2d_ctx* ctx, int32_t N, uint64_t seed, int32_t mode, uint32_t first_id);
/* full state injection (uv, heading, vid, r3d as a previous step left them); used by t2d_step_host */
int t2d_set_state(t2d_ctx* ctx, int32_t N, const double* uv, const int32_t* heading, const int32_t* vid,
                  const double* r3d, const uint32_t* ids);
/* Asynchronous export (the output path of _2DTissue::update, 2DTissue.cpp:148-162, 270-280, without stalling the step):
 * t2d_export_begin snapshots the resident state into a device staging buffer on the step stream and starts its copy into
 * a pinned host ring on a SIDE stream; it returns at once with a slot number (two slots), and the caller keeps stepping.
 * t2d_export_wait blocks until that slot's copy has landed and hands out pointers into the pinned ring, in the reference's
 * layouts (r_UV and r_dot N x 2 column-major, r_3D N x 3, n, vertices_3D_active, particles_color); they stay valid until the
 * slot is reused by the second t2d_export_begin after this one.  Single-context mode only. */
int t2d_export_begin(t2d_ctx* ctx, int32_t* slot);
int t2d_export_wait(t2d_ctx* ctx, int32_t slot, int32_t* N, int64_t* step, const double** uv, const int32_t** heading,
                    const int32_t** vid, const double** r3d, const double** rdot, const int32_t** color);
/* any output pointer may be NULL.  Order = the order of the last upload (ascending position in ids). */
int t2d_download(t2d_ctx* ctx, double* uv, int32_t* heading, int32_t* vid, double* r3d, double* rdot, int32_t* color,
                 int32_t* face);
int32_t t2d_particle_count(const t2d_ctx* ctx);

/* ---- the hot path ---------------------------------------------------------------------------------- */
/* nsteps times the body of _2DTissue::perform_particle_simulation (2DTissue.cpp:222-252) on the resident
 * state; asynchronous work is finished before returning.  Returns the fault mask OR-ed over the steps. */
int t2d_step(t2d_ctx* ctx, int32_t nsteps);
/* the same, with HOST buffers in the reference's layouts, all in/out like the driver's Eigen members:
 * upload -> one step -> download.  This is the literal drop-in for perform_particle_simulation. */
int t2d_step_host(t2d_ctx* ctx, int32_t N, double* uv, int32_t* heading, int32_t* vid, double* r3d, double* rdot,
                  int32_t* color);
/* the same, but only r_UV and n travel to the device: r_3D and vertices_3D_active are functions of r_UV
 * (CellHelper::get_r3d, what the previous step left in the driver's members), so the device re-projects the uploaded
 * r_UV instead of receiving them — 20 instead of 48 bytes per particle over PCIe.  vid and r3d are outputs only. */
int t2d_step_host_uv(t2d_ctx* ctx, int32_t N, double* uv, int32_t* heading, int32_t* vid_out, double* r3d_out, double* rdot,
                     int32_t* color);
int t2d_observables(t2d_ctx* ctx, double out[T2D_OBS_LEN]);
int t2d_get_counters(t2d_ctx* ctx, t2d_counters* out);
int t2d_reset_counters(t2d_ctx* ctx);
/* fp32 fast path: switch the near-cutoff tie log on (1) or off (0, the default).  The log is how the parity tests tell a
   legitimate fp32 rounding difference of a neighbour set from an error (every difference against the fp64 oracle must be a
   logged tie); it runs the same kernel with one extra test per candidate (about +15 % step time), results are identical.
   T2D_COUNT_TIES=1 in the environment switches it on for every new context. */
int t2d_set_tie_log(t2d_ctx* ctx, int on);
/* current step index (the Philox counter's step word); t2d_set_step supports checkpoint/resume */
int64_t t2d_get_step(const t2d_ctx* ctx);
int t2d_set_step(t2d_ctx* ctx, int64_t step);
int t2d_set_params(t2d_ctx* ctx, const t2d_params* params); /* v0,k,sigma,step_size,eta,color_factor,seed may change */

/* ---- single stages, for parity tests against the reference's individual functions ------------------ */
/* CellHelper::get_r3d (CellHelper.cpp:71-85) on host points; does not touch the resident state */
int t2d_get_r3d(t2d_ctx* ctx, int32_t N, const double* uv, double* r3d, int32_t* vid, int32_t* face);
/* EuclideanTiling::diagonal_seam_edges_square_border (EuclideanTiling.cpp:31-69); all three in/out */
int t2d_tiling(t2d_ctx* ctx, int32_t N, double* uv_old, double* uv, int32_t* heading);
/* LinearAlgebra::angles_to_unit_vectors (LinearAlgebra.cpp:25-46): out = N cos then N sin */
int t2d_angles_to_unit_vectors(t2d_ctx* ctx, int32_t N, const int32_t* heading, double* out);
/* forces, new headings and colour of stage 2-3 on the resident state without moving it:
 * ForceHelper::calculate_forces_between_particles (ForceHelper.cpp:34-73),
 * OrientationHelper::calculate_average_n_within_distance (OrientationHelper.cpp:29-74),
 * _2DTissue::count_particle_neighbors (2DTissue.cpp:254-268).  Outputs in upload order; any may be NULL. */
int t2d_forces(t2d_ctx* ctx, double* F, int32_t* new_heading, int32_t* color);
/* hop-count table built on the GPU (replaces DijkstraDistanceHelper.cpp:27-111); out = uint8[V][V] */
int t2d_build_hop_table(t2d_ctx* ctx, uint8_t* out);

/* ---- timing ---------------------------------------------------------------------------------------- */
/* CUDA-event time of the last t2d_step call on the library's stream, in ms */
double t2d_last_step_ms(const t2d_ctx* ctx);
/* per-kernel CUDA-event breakdown of ONE extra step run with events between kernels (does advance the
 * state).  names/ms hold up to cap entries; returns the number of entries. */
int t2d_profile_step(t2d_ctx* ctx, const char** names, double* ms, int cap);

/* ---- pinned host memory for the reference-layout arrays of t2d_step_host (optional convenience) ------ */
void* t2d_pinned_alloc(size_t bytes);
void t2d_pinned_free(void* p);

/* ---- multi-GPU: one context per rank, spatial slabs along the 3-D x axis ----------------
 * Euclidean criterion only.  Rank r owns the particles whose 3-D x lies in [cuts[r-1], cuts[r]) (cuts[-1] = -inf,
 * cuts[world-1] = +inf); mesh, cos/sin tables and the cell index are replicated.  Once per step every rank sends
 * to slab r-1 and r+1 (a) the particles that migrated there (full state) and (b) copies of its particles within
 * r_max of the cut (halo: position, heading, uv, id), in ONE fixed-capacity message per direction; counts travel in
 * the message header, so the host never synchronises inside t2d_step.  Global ids travel with the particles: the
 * summation order and the (seed, step, id) noise are partition independent, results equal the single-GPU ones.
 * After t2d_comm_init*, t2d_set_state/t2d_set_particles take ONLY the particles this rank owns (with their global
 * ids); t2d_download returns the currently owned particles in device order, t2d_download_ids their global ids,
 * t2d_owned_count how many there are.  capacity must leave room for halo copies and migration imbalance. */
#define T2D_UNIQUE_ID_BYTES 128
int t2d_comm_unique_id(uint8_t id[T2D_UNIQUE_ID_BYTES]);
/* one process per GPU, NCCL transport (ncclSend/ncclRecv to rank-1 and rank+1 inside one group).
 * cuts: world-1 ascending interior boundaries.  Collective: every rank of the communicator must call it. */
int t2d_comm_init(t2d_ctx* ctx, int rank, int world, const uint8_t id[T2D_UNIQUE_ID_BYTES], const double* cuts);
/* one process driving `world` contexts (on the same or on different GPUs): the exchange is a device-to-device
 * copy into the neighbour context's receive buffer.  Lets P logical slabs run on ONE GPU (parity tests) or one
 * host thread drive a whole box.  Step the group with t2d_step_local, not t2d_step. */
int t2d_comm_init_local(t2d_ctx** ctxs, int world, const double* cuts);
int t2d_step_local(t2d_ctx** ctxs, int world, int32_t nsteps);   /* returns the OR of all ranks' fault masks */
int t2d_comm_destroy(t2d_ctx* ctx);
int32_t t2d_owned_count(t2d_ctx* ctx);                  /* synchronises; also refreshes t2d_particle_count */
int t2d_download_ids(t2d_ctx* ctx, uint32_t* ids);      /* global ids in the order t2d_download uses */

#ifdef __cplusplus
}
#endif
#endif /* T2D_H */
