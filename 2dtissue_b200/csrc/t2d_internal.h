// t2d_internal.h — device-side data layout shared by the kernel translation units and the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/t2d.h"
#include "hd_math.cuh"

namespace t2d {

// ---- mesh / chart in HBM (replicated on every GPU; a few MB, L2-resident) -----------------------------
template <typename R> struct alignas(16) TriUV {   // one UV triangle, corners a,b,c
    R ax, ay, bx, by, cx, cy;
};
template <typename R> struct alignas(16) Pos3 {    // one 3-D vertex (w pads to 16/32 B)
    R x, y, z, w;
};

template <typename R> struct DevMesh {
    int V = 0, F = 0, G = 0;                 // G x G uniform grid over the UV square
    const TriUV<R>* tri = nullptr;           // [F]
    const int4* tri_vid = nullptr;           // [F] vertex ids (a,b,c,-)
    const Pos3<R>* x3d = nullptr;            // [V]
    const int* gstart = nullptr;             // [G*G+1] CSR cell -> faces (ascending face id)
    const int* gfaces = nullptr;
    R eucl_origin[3] = {0, 0, 0};            // 3-D cell-list origin (mesh bbox min minus one cell)
};

// per-vertex CSR of table entries that can matter: d < 2σ or d <= color_factor·σ (symmetrised by min,
// Locomotion.cpp:110), values kept as double == the reference's in-memory precision
struct DevCSR {
    int V = 0;
    const int* start = nullptr;   // [V+1]
    const int* col = nullptr;     // [nnz] ascending u
    const double* d = nullptr;    // [nnz]
};

// ---- particle state, SoA, in "slot" order (sorted by bucket key of the last binning) ------------------
template <typename R> struct alignas(2 * sizeof(R)) Real2 { R x, y; };
template <typename R> struct ParticleArrays {
    Real2<R>* uv = nullptr;      // chart coordinates (r_UV)
    int2* hv = nullptr;          // x = heading n (integer degrees), y = nearest-vertex id (vertices_3D_active)
    Pos3<R>* X = nullptr;        // 3-D position of the last projection (r_3D); w unused
    int* face = nullptr;         // face of the last projection
    uint32_t* id = nullptr;      // global particle id (RNG counter + accumulation order)
    uint32_t* origin = nullptr;  // index in the caller's arrays
};

struct DevCounters {   // mirrors t2d_counters' device-updated fields
    unsigned long long pairs_in_range, ties_cutoff, ties_trunc, wraps, wrap_cap_hits, order_fallbacks,
        trig_fallbacks, locate_fallbacks, max_row, lost, nonfinite;
    unsigned int fault;
    unsigned int pad;
};

// everything a kernel launch needs, passed by value
template <typename R> struct StepArgs {
    int N = 0;
    ParticleArrays<R> cur, alt;
    // temporaries (slot order)
    uint32_t* key = nullptr;      // bucket key of each particle
    uint32_t* rank = nullptr;     // arrival rank inside its bucket
    int* count = nullptr;         // [M] bucket histogram (zero between steps)
    int* start = nullptr;         // [M+1] exclusive scan
    int* blocksums = nullptr;
    int M = 0;                    // number of buckets (V in table mode, hash size in Euclid mode)
    Real2<R>* uv_new = nullptr;   // position after the Euler step, before seam re-entry
    Real2<R>* rdot = nullptr;
    Real2<R>* F = nullptr;
    int* new_heading = nullptr;
    int* color = nullptr;
    DevCounters* counters = nullptr;
    const double2* trig_d = nullptr;   // [TRIG_N] (cos, sin) of integer degrees, built by the host's libm
    const float2* trig_f = nullptr;
    DevMesh<R> mesh;
    DevCSR csr;
    // parameters
    R v0, k, two_sigma, color_r, step_size, cell_size, inv_cell;
    double eta360;
    double two_sigma_d, color_r_d;   // table predicates are evaluated on doubles (the table's stored precision)
    uint64_t seed, step;
    uint32_t hash_mask;
    int mode;
    int write_F;
    int* work_counter = nullptr;   // dynamic bucket queue for the table-mode kernel
};

// kernel launchers implemented once per precision (step_f64.cu with --fmad=false, step_f32.cu with FMA)
template <typename R> struct Launch {
    static void count_keys(const StepArgs<R>& a, cudaStream_t s);
    static void reorder(const StepArgs<R>& a, cudaStream_t s);
    static void neigh_table(const StepArgs<R>& a, cudaStream_t s, int sm_count);
    static void neigh_euclid(const StepArgs<R>& a, cudaStream_t s);
    static void wrap_project(const StepArgs<R>& a, cudaStream_t s);
    static void project_only(const StepArgs<R>& a, cudaStream_t s);    // initial projection (get_r3d)
    static void tiling_only(const StepArgs<R>& a, Real2<R>* uv_old, Real2<R>* uv, int* heading, int N, cudaStream_t s);
    static void unit_vectors(const StepArgs<R>& a, const int* heading, R* out, int N, cudaStream_t s);
};

// precision-independent kernels (common.cu)
void launch_scan(int* count, int* start, int* blocksums, int M, cudaStream_t s);   // exclusive scan, zeroes count
int scan_blocks(int M);
void launch_observables(const int2* hv, const void* rdot, int is_f32, int N, const double2* trig, double* out8,
                        cudaStream_t s);

}  // namespace t2d
