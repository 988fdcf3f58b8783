// kernels.cuh — the step's CUDA kernels, templated on the arithmetic type.
//
//   k_bin               bucket key of every resident particle + arrival rank + histogram (K1; after uploads only —
//                       during stepping the producing kernel emits the next keys itself)
//   k_scatter           counting-sort scatter of the SoA state into bucket order (K2); makes the 32-byte records of the fp32 path
//   k_scatter_lean(_comm), k_expand   the lean sort of the fp32 fast path: record + aux + source index only; full state on demand
//   k_build_nbr         setup: static neighbourhood table of the sparse row index (nine compact-cell runs per cell)
//   (k_step_fast2       the benchmarked fp32 kernel, stages 2-5 FUSED for the Euclidean criterion: step_fast2.cuh)
//   k_step_euclid_fast  the round-1 fp32 kernel (T2D_STEP=legacy): sparse row index of the cell list, 3 x 3 rows of 3-cell
//                       x-runs, force + alignment (+ noise), Euler, seam re-entry, re-projection, next key
//   k_step_euclid_exact the same, fp64 parity path: in-range neighbours summed in ascending global id
//   k_neigh_table       stages 2-4a for the vertex-distance-table criterion, fp64: one CTA per bucket, neighbour
//                       buckets from the thresholded CSR row, shared-memory staging, exact ascending-id sums (K3)
//   k_neigh_table_warp  the same for fp32: one warp per bucket, coalesced neighbour loads broadcast by shuffles
//   k_wrap_project      table mode: seam re-entry + UV point location + 3-D lift + validation + next key (K4+K5)
//   k_comm_pack / k_comm_unpack / k_comm_unpack_far   slab exchange (multi-GPU): classify + pack, append what arrived
//
// No tensor cores anywhere: the work is gather/scatter + O(10^2) flop per particle.
#pragma once
#include "t2d_internal.h"

namespace t2d {

// ---------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------
// (cos, sin) of an integer-degree heading exactly as the reference's libm call sees it (host-built table)
// outside the host-built table: CUDA libm (may differ from glibc in the last ulp) — counted by the caller.  Out of line:
// double-precision cos + sin are several hundred instructions and this never runs in practice.
static __device__ __noinline__ double2 trig_slow(int n)
{
    double r = (double)n * DEG_TO_RAD_D;
    return make_double2(cos(r), sin(r));
}
__device__ __forceinline__ double2 trig_lookup(const double2* __restrict__ tab, int n, unsigned long long& fb)
{
    if (n >= TRIG_MIN && n <= TRIG_MAX) return __ldg(&tab[n - TRIG_MIN]);
    fb++;
    return trig_slow(n);
}

template <typename R> __device__ __forceinline__ R dev_floor(R v);
template <> __device__ __forceinline__ double dev_floor<double>(double v) { return floor(v); }
template <> __device__ __forceinline__ float dev_floor<float>(float v) { return floorf(v); }

// block-wide accumulation of diagnostic counters: one global atomic per warp per non-zero counter
struct BlockCounters {
    unsigned long long pairs = 0, ties_cut = 0, ties_trunc = 0, wraps = 0, caps = 0, order_fb = 0, trig_fb = 0,
                       loc_fb = 0, max_row = 0, lost = 0, nonfinite = 0, cell_fb = 0;
    unsigned fault = 0;
};

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_max(unsigned long long v)
{
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_down_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// every thread of the block must call this (once, at the end of the kernel); only counters that some lane of
// the warp touched are reduced, so the common case costs a handful of votes
__device__ __forceinline__ void flush_counters(const BlockCounters& c, DevCounters* g)
{
    const unsigned long long v[11] = {c.pairs, c.ties_cut, c.ties_trunc, c.wraps, c.caps, c.order_fb,
                                      c.trig_fb, c.loc_fb, c.lost, c.nonfinite, c.cell_fb};
    unsigned long long* const dst[11] = {&g->pairs_in_range, &g->ties_cutoff, &g->ties_trunc, &g->wraps, &g->wrap_cap_hits,
                                         &g->order_fallbacks, &g->trig_fallbacks, &g->locate_fallbacks, &g->lost, &g->nonfinite,
                                         &g->cell_fallbacks};
    const bool lane0 = (threadIdx.x & 31) == 0;
#pragma unroll
    for (int q = 0; q < 11; ++q) {
        if (__any_sync(0xffffffffu, v[q] != 0)) {
            unsigned long long t = warp_sum(v[q]);
            if (lane0) atomicAdd(dst[q], t);
        }
    }
    if (__any_sync(0xffffffffu, c.max_row != 0)) {
        unsigned long long mr = warp_max(c.max_row);
        if (lane0) atomicMax(&g->max_row, mr);
    }
    if (__any_sync(0xffffffffu, c.fault != 0)) {
        unsigned f = c.fault;
        for (int o = 16; o > 0; o >>= 1) f |= __shfl_down_sync(0xffffffffu, f, o);
        if (lane0) atomicOr(&g->fault, f);
    }
}

// ---------------------------------------------------------------------------------------------------
// sparse row index: 3-D cell -> compact bucket id, x-run of cells -> one contiguous slot range
// ---------------------------------------------------------------------------------------------------
// cell of a point
template <typename R> __device__ __forceinline__ void cell_coords(const DevVox<R>& vx, const Pos3<R>& X, int c[3])
{
    c[0] = (int)dev_floor<R>((X.x - vx.origin[0]) * vx.inv_cell);
    c[1] = (int)dev_floor<R>((X.y - vx.origin[1]) * vx.inv_cell);
    c[2] = (int)dev_floor<R>((X.z - vx.origin[2]) * vx.inv_cell);
}

// compact index of cell (cx, cy, cz), or -1 if the cell is not in the index
template <typename R> __device__ __forceinline__ int vox_index(const DevVox<R>& vx, int cx, int cy, int cz)
{
    if ((unsigned)cx >= (unsigned)vx.ncx || (unsigned)cy >= (unsigned)vx.ncy || (unsigned)cz >= (unsigned)vx.ncz) return -1;
    const uint2 e = __ldg(&vx.words[((size_t)cz * vx.ncy + cy) * vx.nwx + (cx >> 5)]);
    const unsigned bit = cx & 31;
    if (!((e.x >> bit) & 1u)) return -1;
    return (int)e.y + __popc(e.x & ((1u << bit) - 1u));
}

// compact-cell run [lo, hi) of the cells (x0 - 1 .. x0 + 1, y, z).  Compact indices ascend along x inside a row (also
// across its words), so the particles of the run are contiguous in the sorted state whichever of its cells are
// occupied.  Every row ends with a spare word whose base is the row's end, so word (x >> 5) exists for x = ncx.
// Empty (lo == hi == 0) outside the grid.
template <typename R> __device__ __forceinline__ void row_cells(const DevVox<R>& vx, int x0, int y, int z, int& lo, int& hi)
{
    lo = hi = 0;
    if ((unsigned)y >= (unsigned)vx.ncy || (unsigned)z >= (unsigned)vx.ncz || x0 < 1 || x0 > vx.ncx - 2) return;
    const int xl = x0 - 1, xh = x0 + 2;   // ranks of xl and xh bound the run
    const uint2* row = vx.words + ((size_t)z * vx.ncy + y) * vx.nwx;
    const uint2 e0 = __ldg(row + (xl >> 5));
    const int l = (int)e0.y + __popc(e0.x & ((1u << (xl & 31)) - 1u));
    int h;
    if ((xh >> 5) == (xl >> 5)) {
        h = (int)e0.y + __popc(e0.x & ((1u << (xh & 31)) - 1u));
    } else {   // the run straddles two words of the row
        const uint2 e1 = __ldg(row + (xh >> 5));
        h = (int)e1.y + __popc(e1.x & ((1u << (xh & 31)) - 1u));
    }
    if (h > l) {
        lo = l;
        hi = h;
    }
}

// slot range [b, e) of the particles in those cells
template <typename R> __device__ __forceinline__ void row_range(const StepArgs<R>& a, int x0, int y, int z, int& b, int& e)
{
    int lo, hi;
    row_cells<R>(a.vox, x0, y, z, lo, hi);
    b = e = 0;
    if (hi > lo) {
        b = a.start[lo];
        e = a.start[hi];
    }
}

// setup: the static neighbourhood table (t2d_internal.h NBR_STRIDE).  One thread per 32-cell word of the row index.
template <typename R> __global__ void __launch_bounds__(256) k_build_nbr(DevVox<R> vx, int2* nbr)
{
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nwords = (size_t)vx.ncz * vx.ncy * vx.nwx;
    if (w >= nwords) return;
    const uint2 e = vx.words[w];
    unsigned bits = e.x;
    if (!bits) return;
    const size_t row = w / vx.nwx;
    const int wx = (int)(w % vx.nwx), y = (int)(row % vx.ncy), z = (int)(row / vx.ncy);
    int idx = (int)e.y;
    while (bits) {
        const int x0 = wx * 32 + __ffs(bits) - 1;
        bits &= bits - 1;
        int2* out = nbr + (size_t)idx * NBR_STRIDE;
#pragma unroll
        for (int m = 0; m < 9; ++m) {
            int lo, hi;
            row_cells<R>(vx, x0, y + (m % 3) - 1, z + (m / 3) - 1, lo, hi);
            out[m] = make_int2(lo, hi);
        }
        out[9] = make_int2(0, 0);
        ++idx;
    }
}

// number of resident slots: a host constant, or (slab mode) a device-side count that changes every step
template <typename R> __device__ __forceinline__ int resident_count(const StepArgs<R>& a)
{
    return a.comm.on ? a.comm.state->n : a.N;
}

// bucket key of a particle: nearest-vertex id (table criterion) or compact 3-D cell (Euclidean criterion)
template <typename R> __device__ __forceinline__ uint32_t bucket_key(const StepArgs<R>& a, const Pos3<R>& X, int vid, BlockCounters& bc)
{
    if (a.mode == T2D_NEIGH_TABLE) return (uint32_t)vid;
    int c[3];
    cell_coords<R>(a.vox, X, c);
    int idx = vox_index<R>(a.vox, c[0], c[1], c[2]);
    if (idx < 0) {   // not in the static index (cannot happen for points on the mesh): overflow bucket, searched by everyone
        bc.cell_fb++;
        idx = a.vox.M;
    }
    return (uint32_t)idx;
}

// ---------------------------------------------------------------------------------------------------
// setup: voxelise the mesh surface into the sparse cell index.  One warp per face; lanes stride over the
// cells of the face's (grown) bounding box and mark those whose centre is within `reach` of the triangle
// (reach = half a cell diagonal + tolerance, so every cell the triangle touches is marked).
// occ: one 32-bit occupancy word per 32 cells of a row.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double point_triangle_dist2_3d(const double p[3], const double a[3], const double b[3], const double c[3])
{
    double ab[3], ac[3], ap[3];
    for (int k = 0; k < 3; ++k) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = p[k] - a[k]; }
    auto dot = [](const double* u, const double* v) { return u[0] * v[0] + u[1] * v[1] + u[2] * v[2]; };
    auto d2to = [&](double v, double w) {
        double s = 0;
        for (int k = 0; k < 3; ++k) { double e = a[k] + ab[k] * v + ac[k] * w - p[k]; s += e * e; }
        return s;
    };
    double d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) return d2to(0, 0);
    double bp[3] = {p[0] - b[0], p[1] - b[1], p[2] - b[2]};
    double d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0 && d4 <= d3) return d2to(1, 0);
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) return d2to(d1 / (d1 - d3), 0);
    double cp[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
    double d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) return d2to(0, 1);
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) return d2to(0, d2 / (d2 - d6));
    double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        return d2to(1 - w, w);
    }
    double den = 1.0 / (va + vb + vc);
    return d2to(vb * den, vc * den);
}

template <typename R>
__global__ void __launch_bounds__(256) k_voxelize(DevMesh<R> m, double ox, double oy, double oz, double cs, double reach,
                                                  int ncx, int ncy, int ncz, int nwx, unsigned* occ)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < m.F; f += warps) {
        const int4 tv = m.tri_vid[f];
        const Pos3<R> A = m.x3d[tv.x], B = m.x3d[tv.y], C = m.x3d[tv.z];
        const double a[3] = {(double)A.x, (double)A.y, (double)A.z}, b[3] = {(double)B.x, (double)B.y, (double)B.z},
                     c[3] = {(double)C.x, (double)C.y, (double)C.z};
        const double org[3] = {ox, oy, oz};
        const int nc[3] = {ncx, ncy, ncz};
        int lo[3], hi[3];
        for (int k = 0; k < 3; ++k) {
            double mn = fmin(a[k], fmin(b[k], c[k])) - reach, mx = fmax(a[k], fmax(b[k], c[k])) + reach;
            lo[k] = max(0, (int)floor((mn - org[k]) / cs));
            hi[k] = min(nc[k] - 1, (int)floor((mx - org[k]) / cs));
        }
        const int ex = hi[0] - lo[0] + 1, ey = hi[1] - lo[1] + 1, ez = hi[2] - lo[2] + 1;
        if (ex <= 0 || ey <= 0 || ez <= 0) continue;
        const long long total = (long long)ex * ey * ez;
        for (long long q = lane; q < total; q += 32) {
            const int cx = lo[0] + (int)(q % ex), cy = lo[1] + (int)((q / ex) % ey), cz = lo[2] + (int)(q / ((long long)ex * ey));
            const double p[3] = {ox + (cx + 0.5) * cs, oy + (cy + 0.5) * cs, oz + (cz + 0.5) * cs};
            if (point_triangle_dist2_3d(p, a, b, c) <= reach * reach)
                atomicOr(&occ[((size_t)cz * ncy + cy) * nwx + (cx >> 5)], 1u << (cx & 31));
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K1: bucket key + arrival rank + histogram of the resident state (after uploads)
// ---------------------------------------------------------------------------------------------------
template <typename R> __global__ void __launch_bounds__(256) k_bin(StepArgs<R> a)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    BlockCounters bc;
    if (i < resident_count<R>(a)) {
        uint32_t key = bucket_key<R>(a, a.cur.pos[i], a.cur.aux[i].x, bc);
        a.key[i] = key;
        a.rank[i] = (uint32_t)atomicAdd(&a.count[key], 1);
    }
    flush_counters(bc, a.counters);
}

// ---------------------------------------------------------------------------------------------------
// K2: scatter cur -> alt in bucket order
// ---------------------------------------------------------------------------------------------------
template <typename R> __global__ void __launch_bounds__(256) k_scatter(StepArgs<R> a)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = a.comm.on ? a.comm.state->n_res : a.N;
    if (a.comm.on && i == 0) a.comm.state->n = a.start[a.M];   // residents after this sort (nobody reads n in this kernel)
    if (i >= n) return;
    const uint32_t key = a.key[i];
    if (key == KEY_DROP) return;   // slab mode: halo copy of the previous step, or a particle that left
    const int s = a.start[key] + (int)a.rank[i];
    const Pos3<R> P = a.cur.pos[i];
    a.alt.pos[s] = P;
    const Real2<R> U = a.cur.uv[i];
    a.alt.uv[s] = U;
    a.alt.aux[s] = a.cur.aux[i];
    a.alt.rdot[s] = a.cur.rdot[i];
    a.alt.color[s] = a.cur.color[i];
    if (a.alt.cs) {   // fp32 Euclid path: (cos, sin) of the heading for the neighbour sums of the next step, made here once per
                      // particle from the host-built table instead of being carried through the step kernel and the sort
        unsigned long long fb = 0;
        a.alt.cs[s] = trig_lookup(a.trig_d, (int)P.w, fb);
        if (fb) atomicAdd(&a.counters->trig_fallbacks, fb);
    }
    if (a.alt.rec) {   // fp32 fast path: the 32-byte record the neighbour pass of the next step reads (t2d_internal.h)
        const int n = (int)P.w;
        const int slot = (unsigned)n <= 360u ? n : 362;   // index into the shared-memory trig table of k_step_fast2, 362 = not in it
        a.alt.rec[2 * (size_t)s] = make_float4((float)P.x, (float)P.y, (float)P.z, __int_as_float(slot));
        a.alt.rec[2 * (size_t)s + 1] = make_float4((float)U.x, (float)U.y, __int_as_float((int)key), __int_as_float(n));
    }
}

// K2, lean pipeline of the fp32 fast path (single context): the step kernel left the new state as records (alt.rec) + aux,
// r_dot and colour in pre-sort order.  Only what the next step reads is sorted: record (32 B) + aux (16 B), plus the
// 4-byte source index through which k_expand finds r_dot / colour when a caller asks for them.
static __global__ void __launch_bounds__(256) k_scatter_lean(StepArgs<float> a)
{
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    const float4 r0 = a.cur.rec[2 * (size_t)i], r1 = a.cur.rec[2 * (size_t)i + 1];
    const int key = __float_as_int(r1.z);
    const int s = a.start[key] + (int)a.rank[i];
    a.alt.rec[2 * (size_t)s] = r0;
    a.alt.rec[2 * (size_t)s + 1] = r1;
    const int4 ax = a.cur.aux[i];
    a.alt.aux[s] = ax;
    a.src[s] = i;
    if (a.inv) a.inv[ax.w] = s;   // host-buffer path: caller index -> slot, so that the export can gather instead of scatter
}
// The same in slab mode: the step kernel and the unpack kernels leave pos / uv / key in the pre-sort arrays (the slab
// classification needs them), halo copies of the previous step and leavers carry KEY_DROP, and the resident count lives on
// the device.  Still only record + aux + source index are written in bucket order.
// COMM = false: the same source format outside slab mode (right after an upload: pos / uv / key from k_ingest + k_bin).
template <bool COMM> static __global__ void __launch_bounds__(256) k_scatter_lean_posuv(StepArgs<float> a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = COMM ? a.comm.state->n_res : a.N;
    if (COMM && i == 0) a.comm.state->n = a.start[a.M];   // residents after this sort (nobody reads n in this kernel)
    if (i >= n) return;
    const uint32_t key = a.key[i];
    if (COMM && key == KEY_DROP) return;
    const int s = a.start[key] + (int)a.rank[i];
    const Pos3<float> P = a.cur.pos[i];
    const Real2<float> U = a.cur.uv[i];
    const int h = (int)P.w;
    a.alt.rec[2 * (size_t)s] = make_float4(P.x, P.y, P.z, __int_as_float((unsigned)h <= 360u ? h : 362));
    a.alt.rec[2 * (size_t)s + 1] = make_float4(U.x, U.y, __int_as_float((int)key), __int_as_float(h));
    const int4 ax = a.cur.aux[i];
    a.alt.aux[s] = ax;
    a.src[s] = i;
    if (!COMM && a.inv) a.inv[ax.w] = s;
}
// the full sorted state from the lean one: pos / uv from the record, r_dot / colour through the source index
// (a.alt = the pre-sort side that the last step kernel wrote)
static __global__ void __launch_bounds__(256) k_expand(StepArgs<float> a)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= (a.comm.on ? a.comm.state->n : a.N)) return;
    const float4 r0 = a.cur.rec[2 * (size_t)s], r1 = a.cur.rec[2 * (size_t)s + 1];
    const Pos3<float> P = {r0.x, r0.y, r0.z, (float)__float_as_int(r1.w)};
    const Real2<float> U = {r1.x, r1.y};
    a.cur.pos[s] = P;
    a.cur.uv[s] = U;
    const int i = a.src[s];
    a.cur.rdot[s] = a.alt.rdot[i];
    a.cur.color[s] = a.alt.color[i];
}

// ---------------------------------------------------------------------------------------------------
// heading after alignment: OrientationHelper::calculate_average_n_within_distance lines 62-70 +
// mean_unit_circle_vector_angle_degrees (OrientationHelper.cpp:84-116).  mx,my = sum of the neighbours'
// unit vectors in double on BOTH precision paths, so that the integer heading of the fp32 fast path equals
// the fp64 one whenever the neighbour sets agree (the truncation to int makes the last ulp matter).
// ---------------------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ int heading_from_sum(const StepArgs<R>& a, double mx, double my, uint32_t id, BlockCounters& bc)
{
    double z = T2D_DADD(T2D_DMUL(mx, mx), T2D_DMUL(my, my));   // Eigen normalize(): z = squaredNorm(); if (z > 0) v /= sqrt(z)
    if (z > 0.0) {
        double sq = sqrt(z);
        mx = mx / sq;
        my = my / sq;
    }
    bool tie;
    double angle_degrees = mean_angle_degrees_cr(mx, my, a.cr, &tie);
    if (tie) bc.ties_trunc++;
    int avg = (int)angle_degrees;
    if (a.eta360 != 0.0) avg = (int)T2D_DADD((double)avg, noise_deg(a.eta360, a.seed, a.step, id));
    return avg;
}

// speed and velocity: Locomotion::simulate_flight lines 71-81 (Locomotion.cpp)
template <typename R>
__device__ __forceinline__ Real2<R> velocity_from_force(const StepArgs<R>& a, int heading, R fx, R fy, BlockCounters& bc)
{
    R absF = rsqrt_exact<R>(fx * fx + fy * fy);   // F_track.rowwise().norm()
    double2 t = trig_lookup(a.trig_d, heading, bc.trig_fb);   // angles_to_unit_vectors(n) with the OLD heading
    absF = absF + a.v0;
    Real2<R> rd = {(R)t.x * absF, (R)t.y * absF};
    return rd;
}

// ---------------------------------------------------------------------------------------------------
// K5 core: UV point location + lift.  The reference takes the arg-min of (2-D point-triangle distance,
// face index) over ALL faces (CellHelper.cpp:106-117).  The containing face has distance ~1e-17, so the
// arg-min lies among the faces whose (slightly grown) bounding box covers the point: the uniform grid
// cell lists exactly those, in ascending face id; the same distance function decides between them.
// ---------------------------------------------------------------------------------------------------
template <typename R> __device__ __forceinline__ int locate_face(const DevMesh<R>& m, R px, R py, BlockCounters& bc)
{
    const int G = m.G;
    int gi = (int)dev_floor<R>(px * (R)G), gj = (int)dev_floor<R>(py * (R)G);
    gi = gi < 0 ? 0 : (gi > G - 1 ? G - 1 : gi);
    gj = gj < 0 ? 0 : (gj > G - 1 ? G - 1 : gj);
    const int cell = gj * G + gi;
    int best = -1;
    R bd = 0;
    const int qs = m.gstart[cell], qe = m.gstart[cell + 1];
    for (int q = qs; q < qe; ++q) {
        int f = m.gfaces[q];
        TriUV<R> t = m.tri[f];
        R d = point_triangle_distance<R>(px, py, t.ax, t.ay, t.bx, t.by, t.cx, t.cy);
        if (best < 0 || d < bd) {
            bd = d;
            best = f;
        }
        if (bd == R(0)) break;   // ascending face id: nothing later can beat (0, f)
    }
    const R cover_eps = (sizeof(R) == 8) ? R(1e-9) : R(1e-5);
    if (best < 0 || !(bd <= cover_eps)) {   // not covered by the cell list (or NaN): scan all faces like the reference
        bc.loc_fb++;
        best = 0;
        TriUV<R> t0 = m.tri[0];
        bd = point_triangle_distance<R>(px, py, t0.ax, t0.ay, t0.bx, t0.by, t0.cx, t0.cy);
        for (int f = 1; f < m.F; ++f) {
            TriUV<R> t = m.tri[f];
            R d = point_triangle_distance<R>(px, py, t.ax, t.ay, t.bx, t.by, t.cx, t.cy);
            if (d < bd) {
                bd = d;
                best = f;
            }
        }
    }
    return best;
}

// fp32 fast path only: the face of the previous step still contains the point with a barycentric margin far
// above fp32 noise -> it is the arg-min (every other face is at least that far away); skip the grid scan.
__device__ __forceinline__ bool hint_contains(const TriUV<float>& t, float px, float py)
{
    const float d = (t.bx - t.ax) * (t.cy - t.ay) - (t.by - t.ay) * (t.cx - t.ax);
    const float la = ((t.bx - px) * (t.cy - py) - (t.by - py) * (t.cx - px)) / d;
    const float lb = ((t.cx - px) * (t.ay - py) - (t.cy - py) * (t.ax - px)) / d;
    const float lc = 1.0f - la - lb;
    const float m = 1e-3f;
    return la >= m && lb >= m && lc >= m;
}

template <typename R>
__device__ __forceinline__ void project_point(const DevMesh<R>& m, R px, R py, int hint, int& face, int& vid, Pos3<R>& X,
                                              BlockCounters& bc)
{
    int f = -1;
    TriUV<R> t;
    if constexpr (sizeof(R) == 4) {
        if (hint >= 0) {
            t = m.tri[hint];
            if (hint_contains(t, px, py)) f = hint;
        }
    }
    if (f < 0) {
        f = locate_face<R>(m, px, py, bc);
        t = m.tri[f];
    }
    int4 tv = m.tri_vid[f];
    Pos3<R> A = m.x3d[tv.x], B = m.x3d[tv.y], C = m.x3d[tv.z];
    R Av[3] = {A.x, A.y, A.z}, Bv[3] = {B.x, B.y, B.z}, Cv[3] = {C.x, C.y, C.z}, Xv[3];
    int which = lift_to_3d<R>(px, py, t.ax, t.ay, t.bx, t.by, t.cx, t.cy, Av, Bv, Cv, Xv, m.lift_mode == T2D_LIFT_BARYCENTRIC);
    face = f;
    vid = which == 0 ? tv.x : (which == 1 ? tv.y : tv.z);
    X.x = Xv[0];
    X.y = Xv[1];
    X.z = Xv[2];
}

template <typename R> __device__ __forceinline__ bool dev_finite(R v) { return isfinite(v); }

// stages 4b-5 for one particle: seam re-entry, validation flags, projection.  in/out p, n; out face, vid, X
template <typename R>
__device__ __forceinline__ void wrap_and_project(const StepArgs<R>& a, Real2<R> old, Real2<R>& p, int& n, int hint, int& face,
                                                 int& vid, Pos3<R>& X, BlockCounters& bc)
{
    int wraps = 0;
    bool cap = seam_reentry<R>(old.x, old.y, p.x, p.y, n, wraps);
    bc.wraps += wraps;
    if (cap) {
        bc.caps++;
        bc.fault |= T2D_FAULT_WRAP_CAP;
    }
    if (!inside_square<R>(p.x, p.y)) {   // Validation::error_lost_particles
        bc.lost++;
        bc.fault |= T2D_FAULT_LOST;
    }
    if (!dev_finite<R>(p.x) || !dev_finite<R>(p.y)) {   // Validation::error_invalid_values
        bc.nonfinite++;
        bc.fault |= T2D_FAULT_NONFINITE;
    }
    project_point<R>(a.mesh, p.x, p.y, wraps == 0 ? hint : -1, face, vid, X, bc);
    X.w = (R)n;
}

// ---------------------------------------------------------------------------------------------------
// slab mode: message layout and the per-particle classification used by the step kernels
// ---------------------------------------------------------------------------------------------------
template <typename R> __device__ __forceinline__ CommHeader* msg_header(unsigned char* m) { return reinterpret_cast<CommHeader*>(m); }
template <typename R> __device__ __forceinline__ MigRec<R>* msg_mig(unsigned char* m) { return reinterpret_cast<MigRec<R>*>(m + 16); }
template <typename R> __device__ __forceinline__ GhostRec<R>* msg_ghost(unsigned char* m, int mig_cap)
{
    return reinterpret_cast<GhostRec<R>*>(m + 16 + (size_t)mig_cap * sizeof(MigRec<R>));
}

// one OWNED particle (slot i of `st`, values passed in registers): where does its new x put it?
template <typename R>
__device__ __forceinline__ void slab_classify(const StepArgs<R>& a, const ParticleArrays<R>& st, int i, const Pos3<R>& P,
                                              const Real2<R>& uvp, const Real2<R>& rd, int4 ax, int color, BlockCounters& bc)
{
    const DevComm<R>& cm = a.comm;
    uint32_t key = KEY_DROP;
    const R x = P.x;
    bool keep = true;
    if (x < cm.lo2 + cm.halo || x >= cm.hi2 - cm.halo) {
        // landed beyond the adjacent slab, or inside it but within r_max of ITS other cut (seam re-entry / a very
        // fast particle): the far channel — every rank sees it, so the owner adopts it and whoever has it in its
        // halo strip takes a copy; no halo copy here (slabs are at least 4 r_max wide)
        int dest = 0;
        for (int k = 0; k < cm.world - 1; ++k) dest += (x >= cm.cuts[k]) ? 1 : 0;
        const int slot = atomicAdd(&msg_header<R>(cm.far_send)->n_mig, 1);
        if (slot < FAR_CAP) {
            MigRec<R> r;
            r.pos = P;
            r.uv = uvp;
            r.rdot = rd;
            r.aux = make_int4(ax.x, ax.y, ax.z, 0);
            r.color = color;
            r.pad[0] = dest;
            r.pad[1] = r.pad[2] = 0;
            msg_mig<R>(cm.far_send)[slot] = r;
        } else {
            bc.fault |= T2D_FAULT_COMM_OVERFLOW;
        }
        keep = false;
    } else if (x < cm.lo || x >= cm.hi) {
        const int dir = x < cm.lo ? 0 : 1;
        const int slot = atomicAdd(&msg_header<R>(cm.send[dir])->n_mig, 1);
        if (slot < cm.mig_cap) {
            MigRec<R> r;
            r.pos = P;
            r.uv = uvp;
            r.rdot = rd;
            r.aux = make_int4(ax.x, ax.y, ax.z, 0);
            r.color = color;
            r.pad[0] = r.pad[1] = r.pad[2] = 0;
            msg_mig<R>(cm.send[dir])[slot] = r;
        } else {
            bc.fault |= T2D_FAULT_COMM_OVERFLOW;
        }
        keep = dir == 0 ? (x >= cm.lo - cm.halo) : (x < cm.hi + cm.halo);
        if (keep) {   // still within r_max of the cut: our own particles need it as a neighbour — keep a halo copy
            ax.w = ORIGIN_GHOST;
            st.aux[i] = ax;
        }
    } else {
#pragma unroll
        for (int dir = 0; dir < 2; ++dir) {
            const bool has = dir == 0 ? cm.rank > 0 : cm.rank < cm.world - 1;
            const bool near = dir == 0 ? (x < cm.lo + cm.halo) : (x >= cm.hi - cm.halo);
            if (has && near) {
                const int slot = atomicAdd(&msg_header<R>(cm.send[dir])->n_ghost, 1);
                if (slot < cm.ghost_cap) {
                    GhostRec<R> g;
                    g.pos = P;
                    g.uv = uvp;
                    g.id = ax.z;
                    g.pad = 0;
                    msg_ghost<R>(cm.send[dir], cm.mig_cap)[slot] = g;
                } else {
                    bc.fault |= T2D_FAULT_COMM_OVERFLOW;
                }
            }
        }
    }
    if (keep) {
        key = bucket_key<R>(a, P, ax.x, bc);
        a.rank[i] = (uint32_t)atomicAdd(&a.count[key], 1);
    }
    a.key[i] = key;
}

// ---------------------------------------------------------------------------------------------------
// K3-K5 fused (Euclidean criterion).  d_ij = ||X_i - X_j|| on the 3-D positions of the previous projection.
// Cell edge = rmax: a neighbour within rmax lies in the 3 x 3 x 3 cells around the particle's own = 3 x 3 rows of
// 3 x-adjacent cells; a row's 3 cells are ONE contiguous slot range of the sorted state (row_range), so a particle
// walks at most 9 ranges (3.2 non-empty on average on a surface) + the overflow bucket, normally empty.
// One thread per particle; the new state goes to `alt` (other particles still read `cur`), together with
// the particle's next bucket key, arrival rank and the histogram for the counting sort that follows.
//
//   k_step_euclid_exact  fp64 parity path: the in-range neighbours are gathered, sorted by global id and summed in
//                        that order (the reference sums in ascending j), so forces are bit-identical; rows longer
//                        than KMAX are ordered by repeated selection (still exact, counted).
//   k_step_euclid_fast   the round-1 fp32 kernel, kept behind T2D_STEP=legacy for A/B runs: predicates on squared
//                        distances, one rsqrt per in-range pair (d = 0 -> 0.001 as a select), sums in visiting order,
//                        (cos, sin) of the neighbours' headings from the per-particle `cs` array, point location per
//                        thread (previous-face hint, else the grid cell's faces).  The benchmarked fp32 kernel is
//                        k_step_fast2 (step_fast2.cuh); both share fast_epilogue() below.
// ---------------------------------------------------------------------------------------------------
#ifndef T2D_EUCLID_KMAX
#define T2D_EUCLID_KMAX 1024
#endif
// in-range neighbours ordered in the per-thread list (local memory; only the entries in use are touched); longer rows: repeated
// selection, O(row x candidates).  192 in round 1: inside the clumps the lift produces, rows of 200-600 are common after a few
// tens of steps, and the selection path then took ~45 % of the kernel's stall samples at 6 of 32 lanes (ncu source page)
constexpr int EUCLID_KMAX = T2D_EUCLID_KMAX;
#ifndef T2D_STEP_THREADS
#define T2D_STEP_THREADS 128
#endif
constexpr int STEP_THREADS = T2D_STEP_THREADS;
constexpr int NRANGE = 10;   // 3 x 3 rows + the overflow bucket

template <typename R, bool MOVING> __global__ void __launch_bounds__(STEP_THREADS, 4) k_step_euclid_exact(StepArgs<R> a)
{
    __shared__ int s_beg[NRANGE][STEP_THREADS];
    __shared__ int s_end[NRANGE][STEP_THREADS];
    const int tid = threadIdx.x;
    const int i = blockIdx.x * STEP_THREADS + tid;
    BlockCounters bc;
    const bool resident = i < resident_count<R>(a);
    const int4 ai = resident ? a.cur.aux[i] : make_int4(0, 0, 0, ORIGIN_DEAD);
    if (resident && ai.w < 0) {   // slab mode: halo copies are read by others, never advanced; they leave at the next sort
        if (MOVING) {
            a.alt.aux[i] = make_int4(0, -1, ai.z, ORIGIN_DEAD);
            a.key[i] = KEY_DROP;
        }
    } else if (resident) {
        const Pos3<R> Pi = a.cur.pos[i];
        const Real2<R> ui = a.cur.uv[i];
        const int heading = (int)Pi.w;
        int nr = 0;   // non-empty candidate ranges of this particle
        {
            int c[3];
            cell_coords<R>(a.vox, Pi, c);
#pragma unroll
            for (int m = 0; m < 9; ++m) {
                int rb, re;
                row_range<R>(a, c[0], c[1] + (m % 3) - 1, c[2] + (m / 3) - 1, rb, re);
                if (re > rb) {
                    s_beg[nr][tid] = rb;
                    s_end[nr][tid] = re;
                    nr++;
                }
            }
            const int ob = a.start[a.vox.M], oe = a.start[a.vox.M + 1];   // overflow bucket: normally empty
            if (oe > ob) {
                s_beg[nr][tid] = ob;
                s_end[nr][tid] = oe;
                nr++;
            }
        }
        const R rmax = a.two_sigma > a.color_r ? a.two_sigma : a.color_r;
        const R rmax2 = rmax * rmax * R(1.0001);
        R fx = 0, fy = 0;
        double mx = 0, my = 0;
        int color = 0;
        int npairs = 0;

        unsigned long long list[EUCLID_KMAX];
        int cnt = 0;
        bool overflow = false;
        // pass 1: every candidate once: colour, cutoff ties, list of in-range neighbours
        {
            int m = 0, j = nr > 0 ? s_beg[0][tid] : 0, e = nr > 0 ? s_end[0][tid] : 0;
            for (;;) {
                while (j >= e && m < nr - 1) {
                    ++m;
                    j = s_beg[m][tid];
                    e = s_end[m][tid];
                }
                if (j >= e) break;
                const Pos3<R> Pj = a.cur.pos[j];
                const R dx = Pi.x - Pj.x, dy = Pi.y - Pj.y, dz = Pi.z - Pj.z;
                const R d2 = dx * dx + dy * dy + dz * dz;
                if (d2 <= rmax2) {
                    const R d = (j == i) ? R(0) : rsqrt_exact<R>(d2);
                    if (d != R(0) && d <= a.color_r) color++;   // _2DTissue::count_particle_neighbors
                    if (d == a.two_sigma) bc.ties_cut++;
                    if (d < a.two_sigma) {
                        if (cnt < EUCLID_KMAX) list[cnt] = ((unsigned long long)(uint32_t)a.cur.aux[j].z << 32) | (unsigned)j;
                        cnt++;
                    }
                }
                ++j;
            }
        }
        overflow = cnt > EUCLID_KMAX;
        if ((unsigned long long)cnt > bc.max_row) bc.max_row = cnt;
        if (!overflow) {
            // Shell sort by (id, slot), Ciura gaps: ~k^1.3 moves instead of k^2/4 — rows of 50-200 neighbours are
            // common inside the clumps the lift produces (ncu: the plain insertion sort was 31 % of the kernel)
            const int gaps[8] = {701, 301, 132, 57, 23, 10, 4, 1};
#pragma unroll 1
            for (int gi = 0; gi < 8; ++gi) {
                const int gap = gaps[gi];
                if (gap >= cnt) continue;
                for (int p = gap; p < cnt; ++p) {
                    unsigned long long kx = list[p];
                    int q = p - gap;
                    while (q >= 0 && list[q] > kx) {
                        list[q + gap] = list[q];
                        q -= gap;
                    }
                    list[q + gap] = kx;
                }
            }
        } else {
            bc.order_fb++;   // long row: ordered by repeated selection instead of the register list (still exact)
        }
        // pass 2: accumulate in ascending global id, the reference's summation order
        unsigned long long last = 0;
        bool first = true;
        for (int p = 0; p < cnt; ++p) {
            int jj;
            if (!overflow) {
                jj = (int)(unsigned)list[p];
            } else {   // smallest (id, slot) key above `last` among the in-range candidates
                unsigned long long best = ~0ull;
                int m = 0, j = nr > 0 ? s_beg[0][tid] : 0, e = nr > 0 ? s_end[0][tid] : 0;
                for (;;) {
                    while (j >= e && m < nr - 1) {
                        ++m;
                        j = s_beg[m][tid];
                        e = s_end[m][tid];
                    }
                    if (j >= e) break;
                    const Pos3<R> Pj = a.cur.pos[j];
                    const R dx = Pi.x - Pj.x, dy = Pi.y - Pj.y, dz = Pi.z - Pj.z;
                    const R d2 = dx * dx + dy * dy + dz * dz;
                    if (d2 <= rmax2) {
                        const R d = (j == i) ? R(0) : rsqrt_exact<R>(d2);
                        if (d < a.two_sigma) {
                            unsigned long long key = ((unsigned long long)(uint32_t)a.cur.aux[j].z << 32) | (unsigned)j;
                            if ((first || key > last) && key < best) best = key;
                        }
                    }
                    ++j;
                }
                last = best;
                first = false;
                jj = (int)(unsigned)best;
            }
            const Pos3<R> Pj = a.cur.pos[jj];
            const R dx = Pi.x - Pj.x, dy = Pi.y - Pj.y, dz = Pi.z - Pj.z;
            const R d = (jj == i) ? R(0) : rsqrt_exact<R>(dx * dx + dy * dy + dz * dz);
            const double2 t = trig_lookup(a.trig_d, (int)Pj.w, bc.trig_fb);
            mx += t.x;
            my += t.y;
            if (jj != i) {
                npairs++;
                R dd = d;
                if (dd == R(0)) dd += R(0.001);   // ForceHelper.cpp:59-62
                const R Fij = pair_fij<R>(a.k, a.two_sigma, dd);
                const Real2<R> uj = a.cur.uv[jj];
                fx += Fij * ((ui.x - uj.x) / dd);
                fy += Fij * ((ui.y - uj.y) / dd);
            }
        }
        bc.pairs += (unsigned long long)npairs;

        const Real2<R> rd = velocity_from_force<R>(a, heading, fx, fy, bc);
        int n_new = heading_from_sum<R>(a, mx, my, (uint32_t)ai.z, bc);
        if (MOVING) {
            Real2<R> p = {ui.x + rd.x * a.step_size, ui.y + rd.y * a.step_size};   // Locomotion.cpp:84
            int face, vid;
            Pos3<R> X;
            wrap_and_project<R>(a, ui, p, n_new, ai.y, face, vid, X, bc);
            a.alt.pos[i] = X;
            a.alt.uv[i] = p;
            a.alt.aux[i] = make_int4(vid, face, ai.z, ai.w);
            a.alt.rdot[i] = rd;
            a.alt.color[i] = color;
            if (!a.comm.on) {
                const uint32_t key = bucket_key<R>(a, X, vid, bc);
                a.key[i] = key;
                a.rank[i] = (uint32_t)atomicAdd(&a.count[key], 1);
            } else {   // slab mode: stays / migrates / halo copy, messages, key
                slab_classify<R>(a, a.alt, i, X, p, rd, make_int4(vid, face, ai.z, ai.w), color, bc);
            }
        } else {   // t2d_forces: report without moving
            Real2<R> Fv = {fx, fy};
            a.F[i] = Fv;
            a.new_heading[i] = n_new;
            a.alt.color[i] = color;
        }
    }
    flush_counters(bc, a.counters);
}

// ---- fp32 fast path -----------------------------------------------------------------------------------
// first face of the point's grid cell (ascending id) that contains it: the reference's arg-min of (distance, face id)
// for every point that is not within rounding of an edge; -1 if none does (the caller then takes the arg-min itself)
__device__ __forceinline__ int locate_face_contains(const DevMesh<float>& m, float px, float py)
{
    const int G = m.G;
    int gi = (int)floorf(px * (float)G), gj = (int)floorf(py * (float)G);
    gi = gi < 0 ? 0 : (gi > G - 1 ? G - 1 : gi);
    gj = gj < 0 ? 0 : (gj > G - 1 ? G - 1 : gj);
    const int cell = gj * G + gi;
    const int qs = __ldg(&m.gstart[cell]), qe = __ldg(&m.gstart[cell + 1]);
    for (int q = qs; q < qe; ++q) {
        const int f = __ldg(&m.gfaces[q]);
        const TriUV<float> t = m.tri[f];
        const float ex = t.ax - px, ey = t.ay - py, fx = t.bx - px, fy = t.by - py, gx = t.cx - px, gy = t.cy - py;
        const float wa = fx * gy - fy * gx, wb = gx * ey - gy * ex, wc = ex * fy - ey * fx;   // 2 * signed sub-areas
        if ((wa >= 0.0f && wb >= 0.0f && wc >= 0.0f) || (wa <= 0.0f && wb <= 0.0f && wc <= 0.0f)) return f;
    }
    return -1;
}

struct FastSmem {
    int rb[NRANGE][STEP_THREADS], rl[NRANGE][STEP_THREADS];   // candidate ranges (begin, length), longest first; column = thread
};

// accumulators of one particle's neighbour sums
struct PairAcc {
    float fx = 0.0f, fy = 0.0f;
    double mx = 0.0, my = 0.0;
};

// one in-range pair.  F_ij / d = -k (2 sigma - d) / (2 sigma d) = (-k) / d + k / (2 sigma): one FMA on 1/d
// (ForceHelper.cpp:84-104).  d = 0 -> d := 0.001 (ForceHelper.cpp:59-62): the particle itself (ui - uj = 0, no force)
// or one whose 3-D position coincides with it in fp32 while its uv does not — common inside the dense clumps the
// lift produces, so the rule matters: without it 1/d is unbounded and a single pair throws both particles off the chart.
__device__ __forceinline__ void pair_term(const double2* __restrict__ cs, const Real2<float>* __restrict__ uv, int j, float d2,
                                          const Real2<float>& ui, float g1, float g0, PairAcc& acc)
{
    const double2 t = cs[j];
    const Real2<float> uj = uv[j];
    acc.mx += t.x;
    acc.my += t.y;
    const float g = fmaf(d2 == 0.0f ? 1000.0f : rsqrtf(d2), g1, g0);
    acc.fx = fmaf(g, ui.x - uj.x, acc.fx);
    acc.fy = fmaf(g, ui.y - uj.y, acc.fy);
}

// ---------------------------------------------------------------------------------------------------
// per-particle tail of the fp32 fast path, shared by k_step_fast2 (step_fast2.cuh) and the legacy k_step_euclid_fast: speed and velocity
// (Locomotion.cpp:71-81), heading after alignment (+ noise), Euler step (Locomotion.cpp:84), seam re-entry, validation,
// UV point location (previous-face hint, else first containing face of the grid cell, else the distance arg-min),
// lift, next bucket key / slab classification.  `own` = (cos, sin) of the particle's OLD heading.
// ---------------------------------------------------------------------------------------------------
// rare paths of the fast epilogue, out of line (code size: the hot loop shares the instruction cache with them)
// (arguments and results by value: a reference into the caller's registers or into the kernel's parameter block would
// force them into local memory)
struct SeamResult { float px, py; int n, wraps, cap; };
static __device__ __noinline__ SeamResult seam_reentry_f32(float oldx, float oldy, float px, float py, int n)
{
    SeamResult r;
    int wraps = 0;
    r.cap = seam_reentry<float>(oldx, oldy, px, py, n, wraps) ? 1 : 0;
    r.px = px;
    r.py = py;
    r.n = n;
    r.wraps = wraps;
    return r;
}
// returns the face; bit 31 set = the cell list did not cover the point and all faces were scanned (counted by the caller)
static __device__ __noinline__ int locate_face_argmin_f32(int V, int F, int G, const TriUV<float>* tri, const int* gstart,
                                                          const int* gfaces, float px, float py)
{
    DevMesh<float> m;
    m.V = V;
    m.F = F;
    m.G = G;
    m.tri = tri;
    m.gstart = gstart;
    m.gfaces = gfaces;
    BlockCounters lc;
    const int f = locate_face<float>(m, px, py, lc);
    return lc.loc_fb ? (f | (int)0x80000000) : f;
}

template <bool MOVING>
__device__ __forceinline__ void fast_epilogue(const StepArgs<float>& a, int i, const int4 ai, const Real2<float> ui, const double2 own,
                                              const PairAcc& acc, int color, int hits, unsigned& npairs, unsigned& nties,
                                              int old_cell = -1, float ox = 0.0f, float oy = 0.0f, float oz = 0.0f)
{
    typedef float R;
    const float fx = acc.fx, fy = acc.fy;
    npairs = hits > 0 ? (unsigned)(hits - 1) : 0u;   // without itself
    const float absF = sqrtf(fx * fx + fy * fy) + a.v0;
    Real2<R> rd = {(float)own.x * absF, (float)own.y * absF};
    int n_new;
    {   // the sums are doubles on this path too (see heading_from_sum)
        BlockCounters hc;
        n_new = heading_from_sum<R>(a, acc.mx, acc.my, (uint32_t)ai.z, hc);
        nties = (unsigned)hc.ties_trunc;
    }
    if (!MOVING) {   // t2d_forces: report without moving (colour into the scratch side of the double buffer)
        Real2<R> Fv = {fx, fy};
        a.F[i] = Fv;
        a.new_heading[i] = n_new;
        a.alt.color[i] = color;
        return;
    }
    Real2<R> p = {ui.x + rd.x * a.step_size, ui.y + rd.y * a.step_size};   // Locomotion.cpp:84
    Real2<R> old = ui;
    int wraps = 0;
    bool cap = false;
    if (!inside_square<R>(p.x, p.y)) {   // ~1 particle in 10^3 per step
        const SeamResult sr = seam_reentry_f32(old.x, old.y, p.x, p.y, n_new);
        p.x = sr.px;
        p.y = sr.py;
        n_new = sr.n;
        wraps = sr.wraps;
        cap = sr.cap != 0;
    }
    unsigned fault = 0;
    if (wraps) atomicAdd(&a.counters->wraps, (unsigned long long)wraps);
    if (cap) {
        atomicAdd(&a.counters->wrap_cap_hits, 1ull);
        fault |= T2D_FAULT_WRAP_CAP;
    }
    if (!inside_square<R>(p.x, p.y)) {   // Validation::error_lost_particles
        atomicAdd(&a.counters->lost, 1ull);
        fault |= T2D_FAULT_LOST;
    }
    if (!isfinite(p.x) || !isfinite(p.y)) {   // Validation::error_invalid_values
        atomicAdd(&a.counters->nonfinite, 1ull);
        fault |= T2D_FAULT_NONFINITE;
    }
    if (fault) atomicOr(&a.counters->fault, fault);
    const int hint = wraps == 0 ? ai.y : -1;
    int f = hint;
    bool need_locate = true;
    if (hint >= 0) need_locate = !hint_contains(a.mesh.tri[hint], p.x, p.y);
    if (need_locate) {
        // the particle left its previous face (about 1 in 7 per step): first face of its grid cell that contains it
        f = locate_face_contains(a.mesh, p.x, p.y);
        if (f < 0) {   // within rounding of an edge (or outside every listed face): the reference's arg-min over distances
            f = locate_face_argmin_f32(a.mesh.V, a.mesh.F, a.mesh.G, a.mesh.tri, a.mesh.gstart, a.mesh.gfaces, p.x, p.y);
            if (f < 0) {
                f &= 0x7fffffff;
                atomicAdd(&a.counters->locate_fallbacks, 1ull);
            }
        }
    }
    const TriUV<R> t = a.mesh.tri[f];
    const int4 tv = a.mesh.tri_vid[f];
    const Pos3<R> A = a.mesh.x3d[tv.x], B = a.mesh.x3d[tv.y], C = a.mesh.x3d[tv.z];
    const R Av[3] = {A.x, A.y, A.z}, Bv[3] = {B.x, B.y, B.z}, Cv[3] = {C.x, C.y, C.z};
    R Xv[3];
    const int which = lift_to_3d<R>(p.x, p.y, t.ax, t.ay, t.bx, t.by, t.cx, t.cy, Av, Bv, Cv, Xv,
                                    a.mesh.lift_mode == T2D_LIFT_BARYCENTRIC);
    const int vid = which == 0 ? tv.x : (which == 1 ? tv.y : tv.z);
    const Pos3<R> X = {Xv[0], Xv[1], Xv[2], (R)n_new};
    a.alt.aux[i] = make_int4(vid, f, ai.z, ai.w);
    a.alt.rdot[i] = rd;
    a.alt.color[i] = color;
    if (a.comm.on) {   // slab mode: stays / migrates / halo copy, messages, key
        a.alt.pos[i] = X;
        a.alt.uv[i] = p;
        BlockCounters sbc;
        slab_classify<R>(a, a.alt, i, X, p, rd, make_int4(vid, f, ai.z, ai.w), color, sbc);
        if (sbc.fault) atomicOr(&a.counters->fault, sbc.fault);
        if (sbc.cell_fb) atomicAdd(&a.counters->cell_fallbacks, sbc.cell_fb);
    } else {
        int c[3];
        cell_coords<R>(a.vox, X, c);
        int idx = old_cell;
        bool same = false;
        if ((unsigned)old_cell < (unsigned)a.vox.M) {   // a particle moves ~1 % of a cell per step: same cell coordinates as before
            const Pos3<R> O = {ox, oy, oz, 0.0f};       // -> same compact cell, without the row-word lookup (a scattered 8-byte
            int c0[3];                                  // read in a table of several hundred MB: one DRAM sector per particle)
            cell_coords<R>(a.vox, O, c0);
            same = c[0] == c0[0] && c[1] == c0[1] && c[2] == c0[2];
        }
        if (!same) {
            idx = vox_index<R>(a.vox, c[0], c[1], c[2]);
            if (idx < 0) {   // not in the static index (cannot happen for points on the mesh): overflow bucket
                atomicAdd(&a.counters->cell_fallbacks, 1ull);
                idx = a.vox.M;
            }
        }
        if (a.lean) {   // lean pipeline (api.cu): the new state leaves as the 32-byte record the sort and the next step read
            a.alt.rec[2 * (size_t)i] = make_float4(X.x, X.y, X.z, __int_as_float((unsigned)n_new <= 360u ? n_new : 362));
            a.alt.rec[2 * (size_t)i + 1] = make_float4(p.x, p.y, __int_as_float(idx), __int_as_float(n_new));
        } else {
            a.alt.pos[i] = X;
            a.alt.uv[i] = p;
            a.key[i] = (uint32_t)idx;
        }
        a.rank[i] = (uint32_t)atomicAdd(&a.count[idx], 1);
    }
}

#ifndef T2D_UNROLL
#define T2D_UNROLL 4
#endif
constexpr int FAST_UNROLL = T2D_UNROLL;   // candidates per trip of the inner loop (loads of a trip are issued together)
#ifndef T2D_FAST_MIN_BLOCKS
#define T2D_FAST_MIN_BLOCKS 8
#endif

template <bool MOVING> __global__ void __launch_bounds__(STEP_THREADS, T2D_FAST_MIN_BLOCKS) k_step_euclid_fast(StepArgs<float> a)
{
    typedef float R;
    __shared__ FastSmem sm;
    const int tid = threadIdx.x;
    const int i = blockIdx.x * STEP_THREADS + tid;
    const bool resident = i < resident_count<R>(a);
    const int4 ai = resident ? a.cur.aux[i] : make_int4(0, 0, 0, ORIGIN_DEAD);
    const bool live = resident && ai.w >= 0;   // slab mode: halo copies are read by others, never advanced
    if (MOVING && resident && !live) {
        a.alt.aux[i] = make_int4(0, -1, ai.z, ORIGIN_DEAD);
        a.key[i] = KEY_DROP;
    }

    unsigned npairs = 0, nties = 0;
    Real2<R> ui = {0.0f, 0.0f};
    int color = 0;

    const float r2s = a.two_sigma * a.two_sigma, r2c = a.color_r * a.color_r;
    const float g1 = -a.k, g0 = a.k / a.two_sigma;
    const Pos3<R>* __restrict__ pos = a.cur.pos;
    const double2* __restrict__ cs = a.cur.cs;
    const Real2<R>* __restrict__ uv = a.cur.uv;
    const unsigned r2c_bits = __float_as_uint(r2c);
    const unsigned tie_s_lo = __float_as_uint(r2s) - 9u, tie_c_lo = r2c_bits - 9u;   // within 8 ulps of a cutoff
    const bool count_ties = a.count_ties != 0;
    unsigned ncut = 0;
    unsigned long long trig_fb = 0;
    PairAcc acc;
    int hits = 0;

    if (live) {
        // candidate ranges: 3 x 3 rows of 3 x-adjacent cells (+ the overflow bucket, normally empty), sorted by
        // length: every lane walks its longest range first, so the lanes of a warp finish their m-th range at about
        // the same time (measured on the bench workload: 67 instead of 92 warp iterations per particle row)
        const Pos3<R> Pi = a.cur.pos[i];
        ui = a.cur.uv[i];
        int rb[9], rl[9];
        int c[3];
        cell_coords<R>(a.vox, Pi, c);
#pragma unroll
        for (int m = 0; m < 9; ++m) {
            int e;
            row_range<R>(a, c[0], c[1] + (m % 3) - 1, c[2] + (m / 3) - 1, rb[m], e);
            rl[m] = e - rb[m];
        }
#define T2D_CSWAP(x, y)                                \
    if (rl[x] < rl[y]) {                               \
        int t_ = rl[x]; rl[x] = rl[y]; rl[y] = t_;     \
        t_ = rb[x]; rb[x] = rb[y]; rb[y] = t_;         \
    }
        // 9-input sorting network (25 compare-exchanges)
        T2D_CSWAP(0, 1) T2D_CSWAP(3, 4) T2D_CSWAP(6, 7) T2D_CSWAP(1, 2) T2D_CSWAP(4, 5) T2D_CSWAP(7, 8)
        T2D_CSWAP(0, 1) T2D_CSWAP(3, 4) T2D_CSWAP(6, 7) T2D_CSWAP(0, 3) T2D_CSWAP(3, 6) T2D_CSWAP(0, 3)
        T2D_CSWAP(1, 4) T2D_CSWAP(4, 7) T2D_CSWAP(1, 4) T2D_CSWAP(2, 5) T2D_CSWAP(5, 8) T2D_CSWAP(2, 5)
        T2D_CSWAP(1, 3) T2D_CSWAP(5, 7) T2D_CSWAP(2, 6) T2D_CSWAP(4, 6) T2D_CSWAP(2, 4) T2D_CSWAP(2, 3)
        T2D_CSWAP(5, 6)
#undef T2D_CSWAP
        int nr = 0;
#pragma unroll
        for (int m = 0; m < 9; ++m) {
            sm.rb[m][tid] = rb[m];
            sm.rl[m][tid] = rl[m];
            nr += rl[m] > 0 ? 1 : 0;
#ifdef T2D_PREFETCH
            if (m > 0 && m < T2D_PREFETCH && rl[m] > 0) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pos + rb[m]));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(cs + rb[m]));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(uv + rb[m]));
            }
#endif
        }
        {
            const int ob = a.start[a.vox.M], ol = a.start[a.vox.M + 1] - ob;
            if (ol > 0) {
                sm.rb[nr][tid] = ob;
                sm.rl[nr][tid] = ol;
                nr++;
            }
        }
#pragma unroll 1
        for (int m = 0; m < nr; ++m) {
            const int jb = sm.rb[m][tid], len = sm.rl[m][tid];
            const Pos3<R>* q = pos + jb;
#pragma unroll FAST_UNROLL
            for (int t = 0; t < len; ++t) {
                const Pos3<R> Pj = q[t];
                const float dx = Pi.x - Pj.x, dy = Pi.y - Pj.y, dz = Pi.z - Pj.z;
                const float d2 = dx * dx + dy * dy + dz * dz;
                // _2DTissue::count_particle_neighbors: 0 != d <= 2.4 sigma, as ONE unsigned compare on the bit patterns
                // (d2 >= 0: bits(d2) - 1 wraps to 0xffffffff for d2 == 0 and keeps the order of positive floats)
                // (as predicated add: the compiler's select form costs one more instruction per candidate)
                asm("{\n\t.reg .pred p;\n\t.reg .u32 t;\n\tadd.u32 t, %1, -1;\n\tsetp.lt.u32 p, t, %2;\n\t@p add.s32 %0, %0, 1;\n\t}"
                    : "+r"(color)
                    : "r"(__float_as_uint(d2)), "r"(r2c_bits));
                if (count_ties) {   // near-cutoff candidates: the "logged ties" of the parity bar (see step_fast2.cuh)
                    const unsigned bm1 = __float_as_uint(d2) - 1u;
                    if ((bm1 - tie_s_lo) <= 16u || (bm1 - tie_c_lo) <= 16u) ncut++;
                }
                if (d2 < r2s) {
                    pair_term(cs, uv, jb + t, d2, ui, g1, g0, acc);
                    hits++;
                }
            }
        }
    }

    __syncwarp();   // reconverge: lanes leave the candidate loops at different times, the tail below is the same for all
    if (live) {
        const double2 own = cs[i];
        fast_epilogue<MOVING>(a, i, ai, ui, own, acc, color, hits, npairs, nties);
    }
    // diagnostic counters: one atomic per warp
    npairs = __reduce_add_sync(0xffffffffu, npairs);
    nties = __reduce_add_sync(0xffffffffu, nties);
    ncut = __reduce_add_sync(0xffffffffu, ncut);
    const unsigned fbw = __reduce_add_sync(0xffffffffu, (unsigned)trig_fb);
    if ((tid & 31) == 0) {
        if (npairs) atomicAdd(&a.counters->pairs_in_range, (unsigned long long)npairs);
        if (nties) atomicAdd(&a.counters->ties_trunc, (unsigned long long)nties);
        if (ncut) atomicAdd(&a.counters->ties_cutoff, (unsigned long long)ncut);
        if (fbw) atomicAdd(&a.counters->trig_fallbacks, (unsigned long long)fbw);
    }
}

// ---------------------------------------------------------------------------------------------------
// K3 (vertex-distance-table criterion).  dist_length(i,j) = D(v_i, v_j) (Locomotion.cpp:94-111), so all
// particles of bucket v share one neighbour set: the particles of the buckets u listed in CSR row v.
// One CTA per bucket (dynamic queue).  The row's particles are staged in shared memory (id, uv, cos n,
// sin n, row entry), sorted by global id, and every particle of the bucket walks the staged list in that
// order — the reference's ascending-j summation.  Rows longer than CAP are processed in unsorted tiles
// (counted as order fallbacks; still within 1e-9 of the reference).
// Writes uv_new / new_heading (consumed by k_wrap_project) and rdot / colour into the resident state.
// ---------------------------------------------------------------------------------------------------
template <typename R> struct TableSmem {
    static constexpr int CAP = (sizeof(R) == 8) ? 2048 : 4096;   // staged neighbours per tile
    static constexpr int ECAP = 1024;                            // distinct row entries per tile
    // per staged element: key 8, cos 8, sin 8, ux R, uy R, id 4, entry 2  (+2 pad so that the total is a multiple of 16)
    static constexpr size_t ELEM_BYTES = (size_t)CAP * (8 + 8 + 8 + 2 * sizeof(R) + 4 + 2 + 2);
    static constexpr size_t BYTES = ELEM_BYTES + (size_t)ECAP * 2 * sizeof(R);
};

template <typename R, bool EXACT, int THREADS> __global__ void __launch_bounds__(THREADS) k_neigh_table(StepArgs<R> a)
{
    constexpr int CAP = TableSmem<R>::CAP;
    constexpr int ECAP = TableSmem<R>::ECAP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(smem_raw);   // (id << 32) | staging index
    double* s_c = reinterpret_cast<double*>(s_key + CAP);
    double* s_s = s_c + CAP;
    R* s_ux = reinterpret_cast<R*>(s_s + CAP);
    R* s_uy = s_ux + CAP;
    uint32_t* s_id = reinterpret_cast<uint32_t*>(s_uy + CAP);
    unsigned short* s_ent = reinterpret_cast<unsigned short*>(s_id + CAP);
    static_assert(TableSmem<R>::ELEM_BYTES % 16 == 0, "per-entry arrays must start 16-byte aligned");
    R* s_edd = reinterpret_cast<R*>(smem_raw + TableSmem<R>::ELEM_BYTES);   // per row entry: d (0 -> 0.001) and F_ij
    R* s_efij = s_edd + ECAP;
    __shared__ int s_v;
    __shared__ int s_red[2];

    BlockCounters bc;
    const int tid = threadIdx.x;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_v = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const int v = s_v;
        if (v >= a.csr.V) break;
        const int pb = a.start[v], pe = a.start[v + 1];
        if (pb == pe) continue;
        const int rb = a.csr.start[v], re = a.csr.start[v + 1];

        // bucket-wide colour count and in-range list length
        if (tid == 0) s_red[0] = s_red[1] = 0;
        __syncthreads();
        {
            int col = 0, kr = 0;
            for (int e = rb + tid; e < re; e += THREADS) {
                int u = a.csr.col[e];
                double d = a.csr.d[e];
                int cu = a.start[u + 1] - a.start[u];
                if (d != 0.0 && d <= a.color_r_d) col += cu;
                if (d < a.two_sigma_d) kr += cu;
                if (d == a.two_sigma_d) bc.ties_cut += (unsigned long long)cu * (unsigned long long)(pe - pb);
            }
            if (col) atomicAdd(&s_red[0], col);
            if (kr) atomicAdd(&s_red[1], kr);
        }
        __syncthreads();
        const int color_bucket = s_red[0];
        const int krange = s_red[1];
        if (tid == 0 && (unsigned long long)krange > bc.max_row) bc.max_row = krange;

        // tile builder: stages list elements [t0, t0 + fill) of the concatenation (row order, slot order);
        // a tile closes after CAP elements or ECAP row entries.  Returns fill.
        auto build_tile = [&](int t0) -> int {
            __syncthreads();
            int fill = 0;     // staged so far (uniform)
            int seen = 0;     // list elements passed so far (uniform)
            int ne = 0;       // row entries used by this tile (uniform)
            for (int e = rb; e < re; ++e) {
                double d = a.csr.d[e];
                if (!(d < a.two_sigma_d)) continue;
                int u = a.csr.col[e];
                int sb = a.start[u], se = a.start[u + 1];
                int len = se - sb;
                if (len == 0) continue;
                int lo = seen, hi = seen + len;
                seen = hi;
                if (hi <= t0) continue;
                if (fill == CAP || ne == ECAP) break;
                int from = (lo < t0) ? (t0 - lo) : 0;
                int to = len;
                if (fill + (to - from) > CAP) to = from + (CAP - fill);
                if (tid == 0) {
                    R dd = (R)d;
                    if (dd == R(0)) dd += R(0.001);   // ForceHelper.cpp:59-62
                    s_edd[ne] = dd;
                    s_efij[ne] = pair_fij<R>(a.k, a.two_sigma, dd);
                }
                for (int q = from + tid; q < to; q += THREADS) {
                    int slot = sb + q;
                    int pos = fill + (q - from);
                    uint32_t idj = (uint32_t)a.cur.aux[slot].z;
                    s_key[pos] = ((unsigned long long)idj << 32) | (unsigned)pos;
                    Real2<R> uj = a.cur.uv[slot];
                    double2 t = trig_lookup(a.trig_d, (int)a.cur.pos[slot].w, bc.trig_fb);
                    s_ux[pos] = uj.x;
                    s_uy[pos] = uj.y;
                    s_c[pos] = t.x;
                    s_s[pos] = t.y;
                    s_id[pos] = idj;
                    s_ent[pos] = (unsigned short)ne;
                }
                fill += (to - from);
                ne++;
            }
            __syncthreads();
            return fill;
        };

        // bitonic sort of the staged keys by global id
        auto sort_tile = [&](int fill) {
            int n = 1;
            while (n < fill) n <<= 1;
            for (int t = fill + tid; t < n; t += THREADS) s_key[t] = ~0ull;
            __syncthreads();
            for (int kk = 2; kk <= n; kk <<= 1) {
                for (int j = kk >> 1; j > 0; j >>= 1) {
                    for (int t = tid; t < n; t += THREADS) {
                        int ixj = t ^ j;
                        if (ixj > t) {
                            unsigned long long x = s_key[t], y = s_key[ixj];
                            bool asc = ((t & kk) == 0);
                            if ((x > y) == asc) {
                                s_key[t] = y;
                                s_key[ixj] = x;
                            }
                        }
                    }
                    __syncthreads();
                }
            }
        };

        // walk the staged tile for one particle; order = s_key order when sorted, staging order otherwise
        auto walk_tile = [&](int fill, bool sorted, uint32_t my_id, R uix, R uiy, R& fx, R& fy, double& mx, double& my) {
            for (int t = 0; t < fill; ++t) {
                int p = sorted ? (int)(unsigned)s_key[t] : t;
                mx += s_c[p];
                my += s_s[p];
                if (s_id[p] != my_id) {
                    int en = s_ent[p];
                    R dd = s_edd[en];
                    R Fij = s_efij[en];
                    fx += Fij * ((uix - s_ux[p]) / dd);
                    fy += Fij * ((uiy - s_uy[p]) / dd);
                }
            }
        };

        auto finish = [&](int slot, Real2<R> ui, int heading, uint32_t my_id, R fx, R fy, double mx, double my) {
            bc.pairs += (unsigned long long)(krange - 1);
            const Real2<R> rd = velocity_from_force<R>(a, heading, fx, fy, bc);
            a.cur.rdot[slot] = rd;
            a.cur.color[slot] = color_bucket;   // the table's diagonal is 0 (checked at upload): self never counts
            Real2<R> un = {ui.x + rd.x * a.step_size, ui.y + rd.y * a.step_size};
            a.uv_new[slot] = un;
            if (a.write_F) {
                Real2<R> Fv = {fx, fy};
                a.F[slot] = Fv;
            }
            a.new_heading[slot] = heading_from_sum<R>(a, mx, my, my_id, bc);
        };

        const int fill0 = build_tile(0);
        const bool single_tile = (fill0 == krange);
        if (!single_tile && tid == 0) bc.order_fb += (unsigned long long)(pe - pb);

        if (single_tile) {
            const int fill = fill0;
            bool sorted = false;
            if (EXACT) {
                sort_tile(fill);
                sorted = true;
            }
            for (int p0 = pb; p0 < pe; p0 += THREADS) {
                int slot = p0 + tid;
                if (slot < pe) {
                    Real2<R> ui = a.cur.uv[slot];
                    int heading = (int)a.cur.pos[slot].w;
                    uint32_t my_id = (uint32_t)a.cur.aux[slot].z;
                    R fx = 0, fy = 0;
                    double mx = 0, my = 0;
                    walk_tile(fill, sorted, my_id, ui.x, ui.y, fx, fy, mx, my);
                    finish(slot, ui, heading, my_id, fx, fy, mx, my);
                }
            }
        } else {
            for (int p0 = pb; p0 < pe; p0 += THREADS) {
                int slot = p0 + tid;
                bool act = slot < pe;
                Real2<R> ui = {R(0), R(0)};
                int heading = 0;
                uint32_t my_id = 0xffffffffu;
                if (act) {
                    ui = a.cur.uv[slot];
                    heading = (int)a.cur.pos[slot].w;
                    my_id = (uint32_t)a.cur.aux[slot].z;
                }
                R fx = 0, fy = 0;
                double mx = 0, my = 0;
                for (int t0 = 0; t0 < krange;) {
                    int fill = build_tile(t0);
                    if (act) walk_tile(fill, false, my_id, ui.x, ui.y, fx, fy, mx, my);
                    if (fill == 0) break;
                    t0 += fill;
                }
                if (act) finish(slot, ui, heading, my_id, fx, fy, mx, my);
            }
        }
    }
    flush_counters(bc, a.counters);
}

// ---------------------------------------------------------------------------------------------------
// K3 for the table criterion, fp32 fast path: one WARP per bucket, no shared memory, no barriers.
// All particles of bucket v share one neighbour set (see above), so the lanes of the warp are the bucket's particles
// (32 at a time) and the neighbour list is walked in lockstep: every load of a neighbour's (uv, heading, id) is one
// uniform request, every lane adds the same unit vector and its own force term.  The fast path sums in staging
// order anyway (the CTA kernel above sorts by id only on the exact path), so the arithmetic is that kernel's
// walk_tile(sorted = false) — but a bucket of 13 particles (config 3 on the refined chart: 75 k buckets) no longer
// occupies a 256-thread CTA and a dozen barriers: measured 2.3 ms -> see DESIGN.md for 1 M particles.
// ---------------------------------------------------------------------------------------------------
template <typename R> __global__ void __launch_bounds__(128) k_neigh_table_warp(StepArgs<R> a)
{
    const int lane = threadIdx.x & 31;
    BlockCounters bc;
    int next = 0;
    if (lane == 0) next = atomicAdd(a.work_counter, 1);
    for (;;) {
        const int v = __shfl_sync(0xffffffffu, next, 0);
        if (v >= a.csr.V) break;
        if (lane == 0) next = atomicAdd(a.work_counter, 1);
        const int pb = a.start[v], pe = a.start[v + 1];
        if (pb == pe) continue;
        const int rb = a.csr.start[v], re = a.csr.start[v + 1];
        int col = 0, kr = 0;
        for (int e = rb + lane; e < re; e += 32) {
            const int u = a.csr.col[e];
            const double d = a.csr.d[e];
            const int cu = a.start[u + 1] - a.start[u];
            if (d != 0.0 && d <= a.color_r_d) col += cu;
            if (d < a.two_sigma_d) kr += cu;
            if (d == a.two_sigma_d) bc.ties_cut += (unsigned long long)cu * (unsigned long long)(pe - pb);
        }
        const int color_bucket = __reduce_add_sync(0xffffffffu, col);
        const int krange = __reduce_add_sync(0xffffffffu, kr);
        if (lane == 0 && (unsigned long long)krange > bc.max_row) bc.max_row = krange;
        for (int p0 = pb; p0 < pe; p0 += 32) {
            const int slot = p0 + lane;
            const bool act = slot < pe;
            Real2<R> ui = {R(0), R(0)};
            int heading = 0;
            uint32_t my_id = 0xffffffffu;
            if (act) {
                ui = a.cur.uv[slot];
                heading = (int)a.cur.pos[slot].w;
                my_id = (uint32_t)a.cur.aux[slot].z;
            }
            R fx = 0, fy = 0;
            double mx = 0, my = 0;
            for (int e = rb; e < re; ++e) {
                const double d = a.csr.d[e];
                if (!(d < a.two_sigma_d)) continue;
                const int u = a.csr.col[e];
                const int sb = a.start[u], se = a.start[u + 1];
                R dd = (R)d;
                if (dd == R(0)) dd += R(0.001);   // ForceHelper.cpp:59-62
                const R Fij = pair_fij<R>(a.k, a.two_sigma, dd);
                for (int q0 = sb; q0 < se; q0 += 32) {   // 32 neighbours per round: one coalesced load each, then broadcast
                    const int q = q0 + lane;
                    Real2<R> uq = {R(0), R(0)};
                    double2 tq = make_double2(0.0, 0.0);
                    uint32_t idq = 0;
                    if (q < se) {
                        uq = a.cur.uv[q];
                        tq = trig_lookup(a.trig_d, (int)a.cur.pos[q].w, bc.trig_fb);
                        idq = (uint32_t)a.cur.aux[q].z;
                    }
                    const int cnt = min(32, se - q0);
                    for (int t = 0; t < cnt; ++t) {
                        const R ujx = __shfl_sync(0xffffffffu, uq.x, t), ujy = __shfl_sync(0xffffffffu, uq.y, t);
                        const double c = __shfl_sync(0xffffffffu, tq.x, t), sn = __shfl_sync(0xffffffffu, tq.y, t);
                        const uint32_t idj = __shfl_sync(0xffffffffu, idq, t);
                        mx += c;
                        my += sn;
                        if (idj != my_id) {
                            fx += Fij * ((ui.x - ujx) / dd);
                            fy += Fij * ((ui.y - ujy) / dd);
                        }
                    }
                }
            }
            if (act) {
                bc.pairs += (unsigned long long)(krange - 1);
                const Real2<R> rd = velocity_from_force<R>(a, heading, fx, fy, bc);
                a.cur.rdot[slot] = rd;
                a.cur.color[slot] = color_bucket;   // the table's diagonal is 0 (checked at upload): self never counts
                Real2<R> un = {ui.x + rd.x * a.step_size, ui.y + rd.y * a.step_size};
                a.uv_new[slot] = un;
                if (a.write_F) {
                    Real2<R> Fv = {fx, fy};
                    a.F[slot] = Fv;
                }
                a.new_heading[slot] = heading_from_sum<R>(a, mx, my, my_id, bc);
            }
        }
    }
    flush_counters(bc, a.counters);
}

// ---------------------------------------------------------------------------------------------------
// slab mode.  After the step kernel wrote the new state: classify every owned particle by
// its new 3-D x — stays / migrates to slab-1 or slab+1 — pack migrants (full state) and halo copies (what the
// neighbour search reads) into the two fixed-capacity messages, and emit key + rank + histogram for what
// remains resident.  A migrant still within r_max of the cut stays behind as a halo copy, so no second
// exchange round is needed.
// ---------------------------------------------------------------------------------------------------
// stand-alone pass over the resident state (the first exchange after an upload; during stepping the step kernels
// classify each particle themselves, right after computing its new position)
template <typename R> __global__ void __launch_bounds__(256) k_comm_pack(StepArgs<R> a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    BlockCounters bc;
    if (i < a.comm.state->n) {
        const int4 ax = a.cur.aux[i];
        if (ax.w >= 0)
            slab_classify<R>(a, a.cur, i, a.cur.pos[i], a.cur.uv[i], a.cur.rdot[i], ax, a.cur.color[i], bc);
        else
            a.key[i] = KEY_DROP;   // halo copy of the previous step
    }
    flush_counters(bc, a.counters);
}

// append what the two neighbours sent behind the resident slots: migrants become owned particles, halo records
// become halo copies; each gets key + rank + histogram entry.  One thread per record index and role.
template <typename R> __global__ void __launch_bounds__(256) k_comm_unpack(StepArgs<R> a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    BlockCounters bc;
    const DevComm<R>& cm = a.comm;
    int cnt[4] = {0, 0, 0, 0};   // migrants from rank-1, halo from rank-1, migrants from rank+1, halo from rank+1
    if (cm.rank > 0) {
        const CommHeader h = *reinterpret_cast<const CommHeader*>(cm.recv[0]);
        cnt[0] = min(h.n_mig, cm.mig_cap);
        cnt[1] = min(h.n_ghost, cm.ghost_cap);
    }
    if (cm.rank < cm.world - 1) {
        const CommHeader h = *reinterpret_cast<const CommHeader*>(cm.recv[1]);
        cnt[2] = min(h.n_mig, cm.mig_cap);
        cnt[3] = min(h.n_ghost, cm.ghost_cap);
    }
    const int base = cm.state->n;
    int off[4];
    off[0] = base;
    off[1] = off[0] + cnt[0];
    off[2] = off[1] + cnt[1];
    off[3] = off[2] + cnt[2];
    const int total = off[3] + cnt[3];
    if (t == 0) {
        cm.state->n_res = min(total, cm.capacity);
        if (total > cm.capacity) bc.fault |= T2D_FAULT_COMM_OVERFLOW;
    }
#pragma unroll
    for (int role = 0; role < 4; ++role) {
        if (t >= cnt[role]) continue;
        const int slot = off[role] + t;
        if (slot >= cm.capacity) continue;
        unsigned char* m = const_cast<unsigned char*>(cm.recv[role >> 1]);
        Pos3<R> P;
        int vid = 0;
        if ((role & 1) == 0) {
            const MigRec<R> r = msg_mig<R>(m)[t];
            P = r.pos;
            vid = r.aux.x;
            a.cur.pos[slot] = r.pos;
            a.cur.uv[slot] = r.uv;
            a.cur.rdot[slot] = r.rdot;
            a.cur.aux[slot] = r.aux;
            a.cur.color[slot] = r.color;
        } else {
            const GhostRec<R> g = msg_ghost<R>(m, cm.mig_cap)[t];
            P = g.pos;
            Real2<R> z = {R(0), R(0)};
            a.cur.pos[slot] = g.pos;
            a.cur.uv[slot] = g.uv;
            a.cur.rdot[slot] = z;
            a.cur.aux[slot] = make_int4(0, -1, g.id, ORIGIN_GHOST);
            a.cur.color[slot] = 0;
        }
        const uint32_t key = bucket_key<R>(a, P, vid, bc);
        a.key[slot] = key;
        a.rank[slot] = (uint32_t)atomicAdd(&a.count[key], 1);
    }
    flush_counters(bc, a.counters);
}

// the far channel: every rank sees every rank's far migrants; the destination slab adopts the particle, a rank whose halo
// zone it landed in takes a halo copy.  Appended behind what k_comm_unpack appended (atomic cursor: the order inside a
// bucket is arrival order everywhere in this library; the exact path sums by global id).
template <typename R> __global__ void __launch_bounds__(256) k_comm_unpack_far(StepArgs<R> a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x, src = blockIdx.y;
    BlockCounters bc;
    const DevComm<R>& cm = a.comm;
    if (src != cm.rank) {
        unsigned char* m = const_cast<unsigned char*>(cm.far_recv) + (size_t)src * cm.far_bytes;
        const int n = min(msg_header<R>(m)->n_mig, FAR_CAP);
        if (t < n) {
            const MigRec<R> r = msg_mig<R>(m)[t];
            const R x = r.pos.x;
            const bool mine = r.pad[0] == cm.rank;
            const bool halo = !mine && x >= cm.lo - cm.halo && x < cm.hi + cm.halo;
            if (mine || halo) {
                const int slot = atomicAdd(&cm.state->n_res, 1);
                if (slot >= cm.capacity) {
                    atomicSub(&cm.state->n_res, 1);
                    bc.fault |= T2D_FAULT_COMM_OVERFLOW;
                } else {
                    a.cur.pos[slot] = r.pos;
                    a.cur.uv[slot] = r.uv;
                    if (mine) {
                        a.cur.rdot[slot] = r.rdot;
                        a.cur.aux[slot] = r.aux;
                        a.cur.color[slot] = r.color;
                    } else {
                        Real2<R> z = {R(0), R(0)};
                        a.cur.rdot[slot] = z;
                        a.cur.aux[slot] = make_int4(0, -1, r.aux.z, ORIGIN_GHOST);
                        a.cur.color[slot] = 0;
                    }
                    const uint32_t key = bucket_key<R>(a, r.pos, r.aux.x, bc);
                    a.key[slot] = key;
                    a.rank[slot] = (uint32_t)atomicAdd(&a.count[key], 1);
                }
            }
        }
    }
    flush_counters(bc, a.counters);
}

// K4+K5 (table mode): seam re-entry, projection, validation (Validation.cpp:40-72), in place, + next key
template <typename R> __global__ void __launch_bounds__(128) k_wrap_project(StepArgs<R> a)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    BlockCounters bc;
    if (i < a.N) {
        const Real2<R> old = a.cur.uv[i];
        Real2<R> p = a.uv_new[i];
        int n = a.new_heading[i];
        int4 ax = a.cur.aux[i];
        int face, vid;
        Pos3<R> X;
        wrap_and_project<R>(a, old, p, n, ax.y, face, vid, X, bc);
        a.cur.uv[i] = p;
        a.cur.pos[i] = X;
        ax.x = vid;
        ax.y = face;
        a.cur.aux[i] = ax;
        const uint32_t key = bucket_key<R>(a, X, vid, bc);
        a.key[i] = key;
        a.rank[i] = (uint32_t)atomicAdd(&a.count[key], 1);
    }
    flush_counters(bc, a.counters);
}

// projection of the resident uv without a step (CellHelper::get_r3d at _2DTissue::start)
template <typename R> __global__ void __launch_bounds__(128) k_project_only(StepArgs<R> a)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    BlockCounters bc;
    if (i < a.N) {
        Real2<R> p = a.cur.uv[i];
        int face, vid;
        Pos3<R> X;
        int4 ax = a.cur.aux[i];
        // ax.y = -1 after an upload; the host-buffer path of fp32 contexts passes the face of the previous call as a hint
        // (accepted only when the point is well inside it, otherwise the full search runs: same result either way)
        project_point<R>(a.mesh, p.x, p.y, ax.y, face, vid, X, bc);
        X.w = a.cur.pos[i].w;
        a.cur.pos[i] = X;
        ax.x = vid;
        ax.y = face;
        a.cur.aux[i] = ax;
    }
    flush_counters(bc, a.counters);
}

// EuclideanTiling alone, on caller-provided arrays
template <typename R>
__global__ void __launch_bounds__(128) k_tiling_only(StepArgs<R> a, Real2<R>* uv_old, Real2<R>* uv, int* heading, int N)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    BlockCounters bc;
    if (i < N) {
        Real2<R> old = uv_old[i], p = uv[i];
        int n = heading[i];
        int wraps = 0;
        bool cap = seam_reentry<R>(old.x, old.y, p.x, p.y, n, wraps);
        bc.wraps += wraps;
        if (cap) {
            bc.caps++;
            bc.fault |= T2D_FAULT_WRAP_CAP;
        }
        uv_old[i] = old;
        uv[i] = p;
        heading[i] = n;
    }
    flush_counters(bc, a.counters);
}

// LinearAlgebra::angles_to_unit_vectors: out = N cos then N sin
template <typename R> __global__ void __launch_bounds__(256) k_unit_vectors(StepArgs<R> a, const int* heading, R* out, int N)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    BlockCounters bc;
    if (i < N) {
        double2 t = trig_lookup(a.trig_d, heading[i], bc.trig_fb);
        out[i] = (R)t.x;
        out[N + i] = (R)t.y;
    }
    flush_counters(bc, a.counters);
}

// ---------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------
inline int div_up(int a, int b) { return (a + b - 1) / b; }

template <typename R>
void Launch<R>::voxelize(const DevMesh<R>& m, const double org[3], double cs, double reach, const int nc[3], int nwx,
                         unsigned* occ, cudaStream_t s)
{
    int grid = div_up(m.F * 32, 256);
    if (grid > 148 * 16) grid = 148 * 16;
    k_voxelize<R><<<grid, 256, 0, s>>>(m, org[0], org[1], org[2], cs, reach, nc[0], nc[1], nc[2], nwx, occ);
}
// slab mode: the resident count lives on the device and changes every step; grids cover the context's capacity
template <typename R> static int launch_extent(const StepArgs<R>& a) { return a.comm.on ? a.comm.capacity : a.N; }

template <typename R> void Launch<R>::bin(const StepArgs<R>& a, cudaStream_t s)
{
    const int n = launch_extent<R>(a);
    if (n > 0) k_bin<R><<<div_up(n, 256), 256, 0, s>>>(a);
}
template <typename R> void Launch<R>::scatter(const StepArgs<R>& a, cudaStream_t s)
{
    const int n = launch_extent<R>(a);
    if (n > 0) k_scatter<R><<<div_up(n, 256), 256, 0, s>>>(a);
}
template <typename R> void Launch<R>::scatter_lean(const StepArgs<R>& a, cudaStream_t s)
{
    if constexpr (sizeof(R) == 4) {
        const int n = launch_extent<R>(a);
        if (n <= 0) return;
        if (a.comm.on)
            k_scatter_lean_posuv<true><<<div_up(n, 256), 256, 0, s>>>(a);
        else if (a.lean)
            launch_pdl(a.pdl != 0, k_scatter_lean, (unsigned)div_up(n, 256), 256u, s, a);   // source = records written by k_step_fast2
        else
            k_scatter_lean_posuv<false><<<div_up(n, 256), 256, 0, s>>>(a);   // source = pos / uv / key (after an upload)
    }
}
template <typename R> void Launch<R>::expand(const StepArgs<R>& a, cudaStream_t s)
{
    if constexpr (sizeof(R) == 4) {
        const int n = launch_extent<R>(a);
        if (n > 0) k_expand<<<div_up(n, 256), 256, 0, s>>>(a);
    }
}
template <typename R> void Launch<R>::step_euclid(const StepArgs<R>& a, bool moving, cudaStream_t s)
{
    const int n = launch_extent<R>(a);
    if (n <= 0) return;
    const int grid = div_up(n, STEP_THREADS);
    if constexpr (sizeof(R) == 4) {
        if (moving)
            k_step_euclid_fast<true><<<grid, STEP_THREADS, 0, s>>>(a);
        else
            k_step_euclid_fast<false><<<grid, STEP_THREADS, 0, s>>>(a);
    } else {
        if (moving)
            k_step_euclid_exact<R, true><<<grid, STEP_THREADS, 0, s>>>(a);
        else
            k_step_euclid_exact<R, false><<<grid, STEP_THREADS, 0, s>>>(a);
    }
}
template <typename R> void Launch<R>::comm_pack(const StepArgs<R>& a, cudaStream_t s)
{
    k_comm_pack<R><<<div_up(a.comm.capacity, 256), 256, 0, s>>>(a);
}
template <typename R> void Launch<R>::comm_unpack(const StepArgs<R>& a, cudaStream_t s)
{
    const int n = a.comm.mig_cap > a.comm.ghost_cap ? a.comm.mig_cap : a.comm.ghost_cap;
    k_comm_unpack<R><<<div_up(n, 256), 256, 0, s>>>(a);
}
template <typename R> void Launch<R>::comm_unpack_far(const StepArgs<R>& a, cudaStream_t s)
{
    if (a.comm.world > 1) k_comm_unpack_far<R><<<dim3(div_up(FAR_CAP, 256), a.comm.world), 256, 0, s>>>(a);
}
template <typename R> size_t Launch<R>::comm_far_bytes() { return 16 + (size_t)FAR_CAP * sizeof(MigRec<R>); }
template <typename R> size_t Launch<R>::comm_message_bytes(int mig_cap, int ghost_cap)
{
    return 16 + (size_t)mig_cap * sizeof(MigRec<R>) + (size_t)ghost_cap * sizeof(GhostRec<R>);
}
template <typename R> void Launch<R>::neigh_table(const StepArgs<R>& a, cudaStream_t s, int sm_count)
{
    if (a.N <= 0) return;
    if constexpr (sizeof(R) == 4) {   // fast path: warp per bucket (the exact path needs the CTA kernel's sort by id)
        cudaMemsetAsync(a.work_counter, 0, sizeof(int), s);
        int grid = sm_count * 8;
        if (grid > div_up(a.csr.V, 4)) grid = div_up(a.csr.V, 4);
        k_neigh_table_warp<R><<<grid, 128, 0, s>>>(a);
        return;
    }
    constexpr int THREADS = 256;
    auto kern = k_neigh_table<R, sizeof(R) == 8, THREADS>;
    size_t smem = TableSmem<R>::BYTES;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    cudaMemsetAsync(a.work_counter, 0, sizeof(int), s);
    int grid = sm_count * 2;
    if (grid > a.csr.V) grid = a.csr.V;
    kern<<<grid, THREADS, smem, s>>>(a);
}
template <typename R> void Launch<R>::wrap_project(const StepArgs<R>& a, cudaStream_t s)
{
    if (a.N > 0) k_wrap_project<R><<<div_up(a.N, 128), 128, 0, s>>>(a);
}
template <typename R> void Launch<R>::project_only(const StepArgs<R>& a, cudaStream_t s)
{
    if (a.N > 0) k_project_only<R><<<div_up(a.N, 128), 128, 0, s>>>(a);
}
template <typename R>
void Launch<R>::tiling_only(const StepArgs<R>& a, Real2<R>* uv_old, Real2<R>* uv, int* heading, int N, cudaStream_t s)
{
    if (N > 0) k_tiling_only<R><<<div_up(N, 128), 128, 0, s>>>(a, uv_old, uv, heading, N);
}
template <typename R> void Launch<R>::unit_vectors(const StepArgs<R>& a, const int* heading, R* out, int N, cudaStream_t s)
{
    if (N > 0) k_unit_vectors<R><<<div_up(N, 256), 256, 0, s>>>(a, heading, out, N);
}

}  // namespace t2d
