// t2d_internal.h — device-side data layout shared by the kernel translation units and the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/t2d.h"
#include "hd_math.cuh"

namespace t2d {

// Programmatic dependent launch (sm_90+): the three kernels of a lean step are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the launch of the next one is processed while the previous
// one drains, and each begins with pdl_wait() (griddepcontrol.wait: the previous grid has completed and its writes
// are visible) before it touches anything.  Stream order is therefore unchanged; only the launch latency overlaps.
// T2D_PDL=0 goes back to plain launches.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// (An explicit griddepcontrol.launch_dependents at the top of every block, so that the next grid is resident even earlier,
// measured slightly worse: 0.350 vs 0.3485 ms on the headline step, 25.8 vs 24.5 us on c2.)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif


// ---- mesh / chart in HBM (replicated on every GPU; a few MB, L2-resident) -----------------------------
template <typename R> struct alignas(16) TriUV {   // one UV triangle, corners a,b,c
    R ax, ay, bx, by, cx, cy;
};
template <typename R> struct alignas(16) Pos3 {    // 3-D point; w carries the particle's heading (integer degrees)
    R x, y, z, w;
};

template <typename R> struct DevMesh {
    int V = 0, F = 0, G = 0;                 // G x G uniform grid over the UV square
    int lift_mode = 0;                       // T2D_LIFT_*
    const TriUV<R>* tri = nullptr;           // [F]
    const int4* tri_vid = nullptr;           // [F] vertex ids (a,b,c,-)
    const Pos3<R>* x3d = nullptr;            // [V]
    const int* gstart = nullptr;             // [G*G+1] CSR cell -> faces (ascending face id)
    const int* gfaces = nullptr;
};

// per-vertex CSR of table entries that can matter: d < 2σ or d <= color_factor·σ (symmetrised by min,
// Locomotion.cpp:110), values kept as double == the reference's in-memory precision
struct DevCSR {
    int V = 0;
    const int* start = nullptr;   // [V+1]
    const int* col = nullptr;     // [nnz] ascending u
    const double* d = nullptr;    // [nnz]
};

// Sparse row index of the 3-D cell list (Euclidean criterion).  Particles live ON the mesh surface, so the set of
// cells that can ever hold a particle is static: the cells within half a cell diagonal of a mesh triangle.  The
// grid is cut into rows along x; a row is a string of 32-cell words {occupancy bits, compact index of the word's
// first occupied cell}.  Compact indices — the sort key of the counting sort — ascend along x inside a row and
// rows follow a Morton curve over (y, z): the particles of ANY run of x-adjacent cells are one contiguous slot
// range [start[rank(x0)], start[rank(x1 + 1)]), found with one 8-byte load and two popcounts, and the bucket
// arrays have M ~ N entries.
template <typename R> struct DevVox {
    int ncx = 0, ncy = 0, ncz = 0;     // cells per axis
    int nwx = 0;                       // 32-cell words per row
    int M = 0;                         // compact cells; bucket M is the overflow bucket (cell not in the index)
    const uint2* words = nullptr;      // [ncz*ncy*nwx] {occupancy, base}
    R origin[3] = {0, 0, 0};
    R inv_cell = 0;
};

// Static neighbourhood table of the sparse row index (fp32 fast path, k_step_fast2).  The set of surface cells never
// changes, so WHICH compact cells the 3 x 3 rows of 3 x-adjacent cells around a cell cover is known when the index is
// built: nbr[c * NBR_STRIDE + m] = [lo, hi) in compact-cell numbers for row m = (dy + 1) + 3 (dz + 1), (0, 0) if empty.
// At run time the candidates are the slots [start[lo], start[hi]): two loads per row instead of a word lookup + popcounts.
constexpr int NBR_STRIDE = 10;   // 9 rows + one spare entry: 80 bytes per cell, 16-byte aligned

// ---- particle state, SoA, in "slot" order (sorted by bucket key) --------------------------------------
template <typename R> struct alignas(2 * sizeof(R)) Real2 { R x, y; };
template <typename R> struct ParticleArrays {
    Pos3<R>* pos = nullptr;   // 3-D position of the last projection (r_3D) + heading n in w
    Real2<R>* uv = nullptr;   // chart coordinates (r_UV)
    int4* aux = nullptr;      // x = nearest-vertex id (vertices_3D_active), y = face, z = global id, w = index in the caller's arrays
    Real2<R>* rdot = nullptr; // velocity of the last step (r_dot)
    int* color = nullptr;     // neighbour count of the last step (particles_color)
    double2* cs = nullptr;    // legacy fp32 Euclid kernel only: (cos, sin) of the heading as the reference's libm gives them; made by k_scatter
    float4* rec = nullptr;    // fp32 Euclid fast path: ONE 32-byte record per slot, made by k_scatter: {x, y, z, trig-table slot},
                              // {u, v, compact cell, heading} — everything the neighbour pass reads of a candidate, in one sector
};

// ---- multi-GPU slabs: message records and the device-side description of one rank's slab ----
constexpr uint32_t KEY_DROP = 0xffffffffu;   // slot leaves the resident state at the next counting sort
constexpr int ORIGIN_GHOST = -1;             // aux.w of a halo copy (read by the neighbour search, never updated)
constexpr int ORIGIN_DEAD = -2;              // aux.w of a slot to be dropped

template <typename R> struct alignas(16) GhostRec {   // what the neighbour search reads of a foreign particle
    Pos3<R> pos;
    Real2<R> uv;
    int id, pad;
};
template <typename R> struct alignas(16) MigRec {     // full state of a particle that changes owner
    Pos3<R> pos;
    Real2<R> uv, rdot;
    int4 aux;
    int color, pad[3];
};
struct CommHeader { int n_mig, n_ghost, pad0, pad1; };   // first 16 bytes of every message
struct DevCommState { int n, n_res; };                   // sorted residents in `cur`; residents after the unpack

constexpr int T2D_MAX_WORLD = 16;   // slabs per job (one box)
constexpr int FAR_CAP = 1024;       // particles per step and rank that may leave for a NON-adjacent slab

template <typename R> struct DevComm {
    int on = 0;
    int rank = 0, world = 1;
    R lo = 0, hi = 0;       // this rank owns x in [lo, hi)
    R lo2 = 0, hi2 = 0;     // outer boundaries of the two adjacent slabs
    R halo = 0;             // r_max (+ rounding margin)
    int mig_cap = 0, ghost_cap = 0, capacity = 0;
    unsigned char* send[2] = {nullptr, nullptr};         // 0: to rank-1, 1: to rank+1
    const unsigned char* recv[2] = {nullptr, nullptr};   // 0: from rank-1, 1: from rank+1
    DevCommState* state = nullptr;
    // far channel: seam re-entry and the rare very fast particle can land anywhere on the surface (the reference's
    // re-entry is not continuous in 3-D), so a particle that leaves for a non-adjacent slab goes into one small
    // message that EVERY rank receives; the new owner adopts it, a rank whose halo zone it lands in takes a halo copy
    R cuts[T2D_MAX_WORLD - 1] = {};             // slab k owns [cuts[k-1], cuts[k])
    unsigned char* far_send = nullptr;           // CommHeader + MigRec[FAR_CAP] (pad[0] = destination slab)
    const unsigned char* far_recv = nullptr;     // `world` slots of far_bytes each; slot r = what rank r sent
    size_t far_bytes = 0;
};

struct DevCounters {   // mirrors t2d_counters' device-updated fields
    unsigned long long pairs_in_range, ties_cutoff, ties_trunc, wraps, wrap_cap_hits, order_fallbacks,
        trig_fallbacks, locate_fallbacks, max_row, lost, nonfinite, cell_fallbacks;
    unsigned int fault;
    unsigned int pad;
};

// everything a kernel launch needs, passed by value
template <typename R> struct StepArgs {
    int N = 0;
    ParticleArrays<R> cur, alt;
    uint32_t* key = nullptr;      // bucket key of each particle of `cur`
    uint32_t* rank = nullptr;     // arrival rank inside its bucket
    int* count = nullptr;         // [M+1] bucket histogram (zero between sorts)
    int* start = nullptr;         // [M+2] exclusive scan
    int* blocksums = nullptr;
    int M = 0;                    // number of buckets (V in table mode, compact cells + 1 in Euclid mode)
    Real2<R>* uv_new = nullptr;   // table mode: position after the Euler step, before seam re-entry
    int* new_heading = nullptr;   // table mode / t2d_forces: heading after alignment
    Real2<R>* F = nullptr;        // t2d_forces only
    DevCounters* counters = nullptr;
    const double2* trig_d = nullptr;   // [TRIG_N] (cos, sin) of integer degrees, built by the host's libm
    const float2* trig_f = nullptr;
    const CrEntry* cr = nullptr;       // [181] double-double cos/sin/angle of integer degrees (cr_tables.h)
    DevMesh<R> mesh;
    DevCSR csr;
    DevVox<R> vox;
    const int2* nbr = nullptr;     // [(M + 1) * NBR_STRIDE] static neighbourhood table (see NBR_STRIDE)
    DevComm<R> comm;
    // parameters
    R v0, k, two_sigma, color_r, step_size;
    double eta360;
    double two_sigma_d, color_r_d;   // table predicates are evaluated on doubles (the table's stored precision)
    uint64_t seed, step;
    int mode;
    int write_F;
    int count_ties = 0;            // fp32 fast path: log candidates within 8 ulps of a cutoff in ties_cutoff (t2d_set_tie_log)
    int* work_counter = nullptr;   // [4] dynamic queues: [0] buckets of the table-mode kernel, [1], [2] chunk queues of k_step_fast2
    const float4* rec_sentinel = nullptr;   // one record that is nobody's neighbour: what the masked tail trips of k_step_fast2 read
    int* src = nullptr;            // lean pipeline: sorted slot -> index in the pre-sort arrays (r_dot, colour)
    int* inv = nullptr;            // host-buffer path: index in the caller's arrays -> sorted slot (written by the lean sorts when set)
    int lean = 0;                  // 1: k_step_fast2 writes records (alt.rec) instead of pos / uv / key (single context, fp32 Euclid)
    int ablate = 0;                // dev builds (-DT2D_F2_ABLATE) only: 1 no candidate loop, 2 no epilogue, 3 loads only
    int pdl = 1;                   // programmatic dependent launch of the lean step's kernels (T2D_PDL=0: plain launches)
    int queue_flip = 0;            // launch number of k_step_fast2: its low bit picks the queue counter (the launch zeroes the other one)
};

// kernel launchers implemented once per precision (step_f64.cu with --fmad=false, step_f32.cu with FMA)
template <typename R> struct Launch {
    static void voxelize(const DevMesh<R>& m, const double org[3], double cs, double reach, const int nc[3], int nwx,
                         unsigned* occ, cudaStream_t s);   // setup: surface cells of the sparse row index
    static void bin(const StepArgs<R>& a, cudaStream_t s);                      // key + rank + histogram of `cur`
    static void scatter(const StepArgs<R>& a, cudaStream_t s);                  // cur -> alt in bucket order
    static void scatter_lean(const StepArgs<R>& a, cudaStream_t s);             // fp32 fast path: records + aux + source index only
    static void expand(const StepArgs<R>& a, cudaStream_t s);                   // fp32 fast path: lean state -> pos / uv / r_dot / colour
    static void step_euclid(const StepArgs<R>& a, bool moving, cudaStream_t s); // stages 2-5 fused, cur -> alt (+ next keys)
    static bool step_fast2(const StepArgs<R>& a, bool moving, int sm_count, cudaStream_t s);   // fp32 only (step_fast2.cuh); false = not available
    static void build_nbr(const DevVox<R>& vx, int2* nbr, cudaStream_t s);      // setup: static neighbourhood table
    static void neigh_table(const StepArgs<R>& a, cudaStream_t s, int sm_count);
    static void wrap_project(const StepArgs<R>& a, cudaStream_t s);             // table mode stages 4b-5, in place (+ next keys)
    static void project_only(const StepArgs<R>& a, cudaStream_t s);             // initial projection (get_r3d)
    static void comm_pack(const StepArgs<R>& a, cudaStream_t s);                // slabs: classify, pack migrants + halo, keys
    static void comm_unpack(const StepArgs<R>& a, cudaStream_t s);              // slabs: append received particles, keys
    static void comm_unpack_far(const StepArgs<R>& a, cudaStream_t s);          // slabs: the far channel (all ranks' messages)
    static size_t comm_far_bytes();
    static size_t comm_message_bytes(int mig_cap, int ghost_cap);
    static void tiling_only(const StepArgs<R>& a, Real2<R>* uv_old, Real2<R>* uv, int* heading, int N, cudaStream_t s);
    static void unit_vectors(const StepArgs<R>& a, const int* heading, R* out, int N, cudaStream_t s);
};

// precision-independent kernels (common.cu)
void launch_scan(int* count, int* start, int* blocksums, int M, cudaStream_t s);   // exclusive scan, zeroes count
int scan_blocks(int M);
void launch_scan_onepass(int* count, int* start, unsigned long long* status, int* ticket, int ticket_base, unsigned seq, int M,
                         cudaStream_t s, bool pdl = false);   // the same in one launch (decoupled look-back); status: [scan_blocks(M)] words
void launch_observables(const void* pos, const void* rdot, const int4* aux, int is_f32, int N, const int* dN,
                        const double2* trig, double* out8, cudaStream_t s);   // aux/dN: slab mode (skip halo copies, device count)

}  // namespace t2d
