// kernels.cuh — the step's CUDA kernels, templated on the arithmetic type.
//
//   k_count_keys      bucket key per particle (nearest-vertex id | hashed 3-D cell) + arrival rank (K1)
//   k_reorder         counting-sort scatter of the SoA state into bucket order (K2)
//   k_neigh_table     stage 2-5 for the vertex-distance-table criterion: one CTA per bucket, neighbour
//                     buckets from the thresholded CSR row, shared-memory staging, exact ascending-id sums (K3)
//   k_neigh_euclid    stage 2-5 for the Euclidean criterion: hashed cell list, 2x2x2 octant stencil (K3)
//   k_wrap_project    seam re-entry + UV point location + 3-D lift + validation flags (K4+K5)
//
// No tensor cores anywhere: the work is gather/scatter + O(10^2) flop per particle.
#pragma once
#include "t2d_internal.h"

namespace t2d {

// ---------------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------------
template <typename R> struct TrigLookup;
template <> struct TrigLookup<double> {
    static __device__ __forceinline__ void get(const StepArgs<double>& a, int n, double& c, double& s, unsigned& fb)
    {
        if (n >= TRIG_MIN && n <= TRIG_MAX) {
            double2 t = __ldg(&a.trig_d[n - TRIG_MIN]);
            c = t.x;
            s = t.y;
        } else {   // outside the host-built table: CUDA libm (may differ from glibc in the last ulp) — counted
            double r = (double)n * DEG_TO_RAD_D;
            c = cos(r);
            s = sin(r);
            fb++;
        }
    }
};
template <> struct TrigLookup<float> {
    static __device__ __forceinline__ void get(const StepArgs<float>& a, int n, float& c, float& s, unsigned& fb)
    {
        if (n >= TRIG_MIN && n <= TRIG_MAX) {
            float2 t = __ldg(&a.trig_f[n - TRIG_MIN]);
            c = t.x;
            s = t.y;
        } else {
            double r = (double)n * DEG_TO_RAD_D;
            c = (float)cos(r);
            s = (float)sin(r);
            fb++;
        }
    }
};

__device__ __forceinline__ uint32_t cell_hash(int cx, int cy, int cz, uint32_t mask)
{
    uint32_t h = (uint32_t)cx * 73856093u ^ (uint32_t)cy * 19349663u ^ (uint32_t)cz * 83492791u;
    h ^= h >> 15;
    h *= 0x2c1b3c6du;
    h ^= h >> 12;
    return h & mask;
}

template <typename R> __device__ __forceinline__ R dev_floor(R v);
template <> __device__ __forceinline__ double dev_floor<double>(double v) { return floor(v); }
template <> __device__ __forceinline__ float dev_floor<float>(float v) { return floorf(v); }

// block-wide accumulation of diagnostic counters: one global atomic per block per counter
struct BlockCounters {
    unsigned long long pairs = 0, ties_cut = 0, ties_trunc = 0, wraps = 0, caps = 0, order_fb = 0, trig_fb = 0,
                       loc_fb = 0, max_row = 0, lost = 0, nonfinite = 0;
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

// every thread of the block must call this (once, at the end of the kernel)
__device__ __forceinline__ void flush_counters(const BlockCounters& c, DevCounters* g)
{
    unsigned f = c.fault;
    for (int o = 16; o > 0; o >>= 1) f |= __shfl_down_sync(0xffffffffu, f, o);
    unsigned long long v[10] = {c.pairs, c.ties_cut, c.ties_trunc, c.wraps, c.caps,
                                c.order_fb, c.trig_fb, c.loc_fb, c.lost, c.nonfinite};
#pragma unroll
    for (int q = 0; q < 10; ++q) v[q] = warp_sum(v[q]);
    unsigned long long mr = warp_max(c.max_row);
    if ((threadIdx.x & 31) == 0) {
        unsigned long long* dst[10] = {&g->pairs_in_range, &g->ties_cutoff, &g->ties_trunc, &g->wraps, &g->wrap_cap_hits,
                                       &g->order_fallbacks, &g->trig_fallbacks, &g->locate_fallbacks, &g->lost, &g->nonfinite};
#pragma unroll
        for (int q = 0; q < 10; ++q)
            if (v[q]) atomicAdd(dst[q], v[q]);
        if (mr) atomicMax(&g->max_row, mr);
        if (f) atomicOr(&g->fault, f);
    }
}

// ---------------------------------------------------------------------------------------------------
// K1: bucket key + arrival rank
// ---------------------------------------------------------------------------------------------------
template <typename R> __device__ __forceinline__ void cell_coords(const StepArgs<R>& a, const Pos3<R>& X, int c[3], int side[3])
{
    R q[3] = {(X.x - a.mesh.eucl_origin[0]) * a.inv_cell, (X.y - a.mesh.eucl_origin[1]) * a.inv_cell,
              (X.z - a.mesh.eucl_origin[2]) * a.inv_cell};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        R fl = dev_floor<R>(q[k]);
        c[k] = (int)fl;
        side[k] = (q[k] - fl < R(0.5)) ? -1 : 1;
    }
}

template <typename R> __global__ void __launch_bounds__(256) k_count_keys(StepArgs<R> a)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    uint32_t key;
    if (a.mode == T2D_NEIGH_TABLE) {
        key = (uint32_t)a.cur.hv[i].y;
    } else {
        int c[3], side[3];
        cell_coords<R>(a, a.cur.X[i], c, side);
        key = cell_hash(c[0], c[1], c[2], a.hash_mask);
    }
    a.key[i] = key;
    a.rank[i] = (uint32_t)atomicAdd(&a.count[key], 1);
}

// ---------------------------------------------------------------------------------------------------
// K2: scatter into bucket order
// ---------------------------------------------------------------------------------------------------
template <typename R> __global__ void __launch_bounds__(256) k_reorder(StepArgs<R> a)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    int s = a.start[a.key[i]] + (int)a.rank[i];
    a.alt.uv[s] = a.cur.uv[i];
    a.alt.hv[s] = a.cur.hv[i];
    a.alt.X[s] = a.cur.X[i];
    a.alt.face[s] = a.cur.face[i];
    a.alt.id[s] = a.cur.id[i];
    if (a.cur.origin != a.cur.id) a.alt.origin[s] = a.cur.origin[i];
}

// ---------------------------------------------------------------------------------------------------
// shared tail of K3: speed, Euler step, new heading.
//   Locomotion::simulate_flight lines 71-84 (Locomotion.cpp) + OrientationHelper.cpp:67-70 (+ noise)
// ---------------------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ void finish_particle(const StepArgs<R>& a, int slot, Real2<R> ui, int heading, uint32_t id, R fx,
                                                R fy, R mx, R my, int color, BlockCounters& bc)
{
    R absF = rsqrt_exact<R>(fx * fx + fy * fy);   // F_track.rowwise().norm()
    R c, s;
    unsigned tf = 0;
    TrigLookup<R>::get(a, heading, c, s, tf);     // angles_to_unit_vectors(n) with the OLD heading
    bc.trig_fb += tf;
    absF = absF + a.v0;
    R rx = c * absF, ry = s * absF;
    Real2<R> rd = {rx, ry};
    a.rdot[slot] = rd;
    Real2<R> un = {ui.x + rx * a.step_size, ui.y + ry * a.step_size};
    a.uv_new[slot] = un;
    if (a.write_F) {
        Real2<R> Fv = {fx, fy};
        a.F[slot] = Fv;
    }
    a.color[slot] = color;

    // mean_unit_circle_vector_angle_degrees, OrientationHelper.cpp:84-116 (Eigen normalize(): z>0 guard)
    double angle_degrees;
    if (sizeof(R) == 8) {
        double dmx = (double)mx, dmy = (double)my;
        double z = dmx * dmx + dmy * dmy;
        if (z > 0.0) {
            double sq = sqrt(z);
            dmx = dmx / sq;
            dmy = dmy / sq;
        }
        angle_degrees = atan2(dmy, dmx) * RAD_TO_DEG_D;
        if (angle_degrees < 0) angle_degrees += 360.0;
        if (fabs(angle_degrees - rint(angle_degrees)) < 1e-9) bc.ties_trunc++;
    } else {
        angle_degrees = (double)atan2f((float)my, (float)mx) * RAD_TO_DEG_D;
        if (angle_degrees < 0) angle_degrees += 360.0;
    }
    int avg = (int)angle_degrees;
    if (a.eta360 != 0.0) avg = (int)((double)avg + noise_deg(a.eta360, a.seed, a.step, id));
    a.new_heading[slot] = avg;
}

// ---------------------------------------------------------------------------------------------------
// K3 (Euclidean criterion).  d_ij = ||X_i - X_j|| on the 3-D positions of the previous projection.
// Hashed cell list with cell edge 2*rmax: a neighbour within rmax lies in the particle's own cell or the
// adjacent one on the nearer side per axis -> 8 buckets.  EXACT: in-range neighbours are gathered, sorted by
// global id and summed in that order (the reference sums in ascending j), so forces and headings are
// bit-identical; rows longer than KMAX fall back to unordered sums (counted).
// ---------------------------------------------------------------------------------------------------
constexpr int EUCLID_KMAX = 48;

template <typename R, bool COLLECT>
__device__ __forceinline__ void euclid_visit(const StepArgs<R>& a, int i, const Pos3<R>& Xi, const Real2<R>& ui,
                                             const uint32_t keys[8], R& fx, R& fy, R& mx, R& my, int& color,
                                             unsigned long long* list, int& cnt, bool& overflow, BlockCounters& bc)
{
    const R rmax = a.two_sigma > a.color_r ? a.two_sigma : a.color_r;
    const R rmax2 = rmax * rmax * R(1.0001);
#pragma unroll 1
    for (int m = 0; m < 8; ++m) {
        bool dup = false;
        for (int p = 0; p < m; ++p) dup |= (keys[p] == keys[m]);
        if (dup) continue;
        int s = a.start[keys[m]], e = a.start[keys[m] + 1];
        for (int j = s; j < e; ++j) {
            Pos3<R> Xj = a.cur.X[j];
            R dx = Xi.x - Xj.x, dy = Xi.y - Xj.y, dz = Xi.z - Xj.z;
            R d2 = dx * dx + dy * dy + dz * dz;
            if (d2 > rmax2) continue;
            R d = (j == i) ? R(0) : rsqrt_exact<R>(d2);
            if (COLLECT) {   // first pass: colour + tie statistics are order independent
                if (d != R(0) && d <= a.color_r) color++;
                if (d == a.two_sigma) bc.ties_cut++;
            }
            if (!(d < a.two_sigma)) continue;
            if (COLLECT) {
                if (cnt < EUCLID_KMAX)
                    list[cnt++] = ((unsigned long long)a.cur.id[j] << 32) | (unsigned)j;
                else
                    overflow = true;
            } else {
                R c, sn;
                unsigned tf = 0;
                TrigLookup<R>::get(a, a.cur.hv[j].x, c, sn, tf);
                bc.trig_fb += tf;
                mx += c;
                my += sn;
                if (j != i) {
                    bc.pairs++;
                    R dd = d;
                    if (dd == R(0)) dd += R(0.001);
                    R Fij = pair_fij<R>(a.k, a.two_sigma, dd);
                    Real2<R> uj = a.cur.uv[j];
                    fx += Fij * ((ui.x - uj.x) / dd);
                    fy += Fij * ((ui.y - uj.y) / dd);
                }
            }
        }
    }
}

template <typename R, bool EXACT> __global__ void __launch_bounds__(128) k_neigh_euclid(StepArgs<R> a)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    BlockCounters bc;
    if (i < a.N) {
        Pos3<R> Xi = a.cur.X[i];
        Real2<R> ui = a.cur.uv[i];
        int2 hvi = a.cur.hv[i];
        int c[3], side[3];
        cell_coords<R>(a, Xi, c, side);
        uint32_t keys[8];
#pragma unroll
        for (int m = 0; m < 8; ++m)
            keys[m] = cell_hash(c[0] + ((m & 1) ? side[0] : 0), c[1] + ((m & 2) ? side[1] : 0),
                                c[2] + ((m & 4) ? side[2] : 0), a.hash_mask);
        R fx = 0, fy = 0, mx = 0, my = 0;
        int color = 0, cnt = 0;
        bool overflow = false;
        if (EXACT) {
            unsigned long long list[EUCLID_KMAX];
            euclid_visit<R, true>(a, i, Xi, ui, keys, fx, fy, mx, my, color, list, cnt, overflow, bc);
            if (!overflow) {
                for (int p = 1; p < cnt; ++p) {   // insertion sort by (id, slot)
                    unsigned long long kx = list[p];
                    int q = p - 1;
                    while (q >= 0 && list[q] > kx) {
                        list[q + 1] = list[q];
                        --q;
                    }
                    list[q + 1] = kx;
                }
                for (int p = 0; p < cnt; ++p) {
                    int j = (int)(unsigned)list[p];
                    R cj, sj;
                    unsigned tf = 0;
                    TrigLookup<R>::get(a, a.cur.hv[j].x, cj, sj, tf);
                    bc.trig_fb += tf;
                    mx += cj;
                    my += sj;
                    if (j != i) {
                        bc.pairs++;
                        Pos3<R> Xj = a.cur.X[j];
                        R dx = Xi.x - Xj.x, dy = Xi.y - Xj.y, dz = Xi.z - Xj.z;
                        R dd = rsqrt_exact<R>(dx * dx + dy * dy + dz * dz);
                        if (dd == R(0)) dd += R(0.001);
                        R Fij = pair_fij<R>(a.k, a.two_sigma, dd);
                        Real2<R> uj = a.cur.uv[j];
                        fx += Fij * ((ui.x - uj.x) / dd);
                        fy += Fij * ((ui.y - uj.y) / dd);
                    }
                }
                if ((unsigned long long)cnt > bc.max_row) bc.max_row = cnt;
            } else {
                bc.order_fb++;
                int dummy_color = 0, dummy_cnt = 0;
                bool dummy_of = false;
                euclid_visit<R, false>(a, i, Xi, ui, keys, fx, fy, mx, my, dummy_color, nullptr, dummy_cnt, dummy_of, bc);
            }
        } else {
            // fast path: statistics and sums in one unordered pass
            const R rmax = a.two_sigma > a.color_r ? a.two_sigma : a.color_r;
            const R rmax2 = rmax * rmax * R(1.0001);
#pragma unroll 1
            for (int m = 0; m < 8; ++m) {
                bool dup = false;
                for (int p = 0; p < m; ++p) dup |= (keys[p] == keys[m]);
                if (dup) continue;
                int s = a.start[keys[m]], e = a.start[keys[m] + 1];
                for (int j = s; j < e; ++j) {
                    Pos3<R> Xj = a.cur.X[j];
                    R dx = Xi.x - Xj.x, dy = Xi.y - Xj.y, dz = Xi.z - Xj.z;
                    R d2 = dx * dx + dy * dy + dz * dz;
                    if (d2 > rmax2) continue;
                    R d = (j == i) ? R(0) : rsqrt_exact<R>(d2);
                    if (d != R(0) && d <= a.color_r) color++;
                    if (!(d < a.two_sigma)) continue;
                    R cj, sj;
                    unsigned tf = 0;
                    TrigLookup<R>::get(a, a.cur.hv[j].x, cj, sj, tf);
                    mx += cj;
                    my += sj;
                    if (j != i) {
                        bc.pairs++;
                        R dd = d;
                        if (dd == R(0)) dd += R(0.001);
                        R Fij = pair_fij<R>(a.k, a.two_sigma, dd);
                        Real2<R> uj = a.cur.uv[j];
                        fx += Fij * ((ui.x - uj.x) / dd);
                        fy += Fij * ((ui.y - uj.y) / dd);
                    }
                }
            }
        }
        finish_particle<R>(a, i, ui, hvi.x, a.cur.id[i], fx, fy, mx, my, color, bc);
    }
    flush_counters(bc, a.counters);
}

// ---------------------------------------------------------------------------------------------------
// K3 (vertex-distance-table criterion).  dist_length(i,j) = D(v_i, v_j) (Locomotion.cpp:94-111), so all
// particles of bucket v share one neighbour set: the particles of the buckets u listed in CSR row v.
// One CTA per bucket (dynamic queue).  The row's particles are staged in shared memory (id, uv, cos n,
// sin n, row entry), sorted by global id, and every particle of the bucket walks the staged list in that
// order — the reference's ascending-j summation.  Rows longer than CAP are processed in unsorted tiles
// (counted as order fallbacks; still within 1e-9 of the reference).
// ---------------------------------------------------------------------------------------------------
template <typename R> struct TableSmem {
    static constexpr int CAP = (sizeof(R) == 8) ? 2048 : 4096;   // staged neighbours per tile
    static constexpr int ECAP = 1024;                            // distinct row entries per tile
    static constexpr size_t ELEM_BYTES = (size_t)CAP * (8 + 4 * sizeof(R) + 4 + 2);   // multiple of 16
    static constexpr size_t BYTES = ELEM_BYTES + (size_t)ECAP * 2 * sizeof(R);
};

template <typename R, bool EXACT, int THREADS> __global__ void __launch_bounds__(THREADS) k_neigh_table(StepArgs<R> a)
{
    constexpr int CAP = TableSmem<R>::CAP;
    constexpr int ECAP = TableSmem<R>::ECAP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* s_key = reinterpret_cast<unsigned long long*>(smem_raw);   // (id << 32) | staging index
    R* s_ux = reinterpret_cast<R*>(s_key + CAP);
    R* s_uy = s_ux + CAP;
    R* s_c = s_uy + CAP;
    R* s_s = s_c + CAP;
    uint32_t* s_id = reinterpret_cast<uint32_t*>(s_s + CAP);
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
                    uint32_t idj = a.cur.id[slot];
                    s_key[pos] = ((unsigned long long)idj << 32) | (unsigned)pos;
                    Real2<R> uj = a.cur.uv[slot];
                    R cj, sj;
                    unsigned tf = 0;
                    TrigLookup<R>::get(a, a.cur.hv[slot].x, cj, sj, tf);
                    bc.trig_fb += tf;
                    s_ux[pos] = uj.x;
                    s_uy[pos] = uj.y;
                    s_c[pos] = cj;
                    s_s[pos] = sj;
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
        auto walk_tile = [&](int fill, bool sorted, uint32_t my_id, R uix, R uiy, R& fx, R& fy, R& mx, R& my) {
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

        const int self_color = 0;   // the diagonal of the table is 0 (checked at upload), so self never counts
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
                    int2 hvi = a.cur.hv[slot];
                    uint32_t my_id = a.cur.id[slot];
                    R fx = 0, fy = 0, mx = 0, my = 0;
                    walk_tile(fill, sorted, my_id, ui.x, ui.y, fx, fy, mx, my);
                    bc.pairs += (unsigned long long)(krange - 1);
                    finish_particle<R>(a, slot, ui, hvi.x, my_id, fx, fy, mx, my, color_bucket - self_color, bc);
                }
            }
        } else {
            for (int p0 = pb; p0 < pe; p0 += THREADS) {
                int slot = p0 + tid;
                bool act = slot < pe;
                Real2<R> ui = {R(0), R(0)};
                int2 hvi = {0, 0};
                uint32_t my_id = 0xffffffffu;
                if (act) {
                    ui = a.cur.uv[slot];
                    hvi = a.cur.hv[slot];
                    my_id = a.cur.id[slot];
                }
                R fx = 0, fy = 0, mx = 0, my = 0;
                for (int t0 = 0; t0 < krange;) {
                    int fill = build_tile(t0);
                    if (act) walk_tile(fill, false, my_id, ui.x, ui.y, fx, fy, mx, my);
                    if (fill == 0) break;
                    t0 += fill;
                }
                if (act) {
                    bc.pairs += (unsigned long long)(krange - 1);
                    finish_particle<R>(a, slot, ui, hvi.x, my_id, fx, fy, mx, my, color_bucket - self_color, bc);
                }
            }
        }
    }
    flush_counters(bc, a.counters);
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

template <typename R> __device__ __forceinline__ void project_point(const DevMesh<R>& m, R px, R py, int& face, int& vid,
                                                                    Pos3<R>& X, BlockCounters& bc)
{
    int f = locate_face<R>(m, px, py, bc);
    TriUV<R> t = m.tri[f];
    int4 tv = m.tri_vid[f];
    Pos3<R> A = m.x3d[tv.x], B = m.x3d[tv.y], C = m.x3d[tv.z];
    R Av[3] = {A.x, A.y, A.z}, Bv[3] = {B.x, B.y, B.z}, Cv[3] = {C.x, C.y, C.z}, Xv[3];
    int which = lift_to_3d<R>(px, py, t.ax, t.ay, t.bx, t.by, t.cx, t.cy, Av, Bv, Cv, Xv);
    face = f;
    vid = which == 0 ? tv.x : (which == 1 ? tv.y : tv.z);
    X.x = Xv[0];
    X.y = Xv[1];
    X.z = Xv[2];
    X.w = R(0);
}

template <typename R> __device__ __forceinline__ bool dev_finite(R v) { return isfinite(v); }

// K4+K5: seam re-entry, projection, validation (Validation.cpp:40-72)
template <typename R> __global__ void __launch_bounds__(128) k_wrap_project(StepArgs<R> a)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    BlockCounters bc;
    if (i < a.N) {
        Real2<R> old = a.cur.uv[i];
        Real2<R> p = a.uv_new[i];
        int n = a.new_heading[i];
        int wraps = 0;
        bool cap = seam_reentry<R>(old.x, old.y, p.x, p.y, n, wraps);
        bc.wraps += wraps;
        if (cap) {
            bc.caps++;
            bc.fault |= T2D_FAULT_WRAP_CAP;
        }
        if (!inside_square<R>(p.x, p.y)) {
            bc.lost++;
            bc.fault |= T2D_FAULT_LOST;
        }
        if (!dev_finite<R>(p.x) || !dev_finite<R>(p.y)) {
            bc.nonfinite++;
            bc.fault |= T2D_FAULT_NONFINITE;
        }
        int face, vid;
        Pos3<R> X;
        project_point<R>(a.mesh, p.x, p.y, face, vid, X, bc);
        a.cur.uv[i] = p;
        int2 hv = {n, vid};
        a.cur.hv[i] = hv;
        a.cur.face[i] = face;
        a.cur.X[i] = X;
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
        project_point<R>(a.mesh, p.x, p.y, face, vid, X, bc);
        int2 hv = a.cur.hv[i];
        hv.y = vid;
        a.cur.hv[i] = hv;
        a.cur.face[i] = face;
        a.cur.X[i] = X;
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
        R c, s;
        unsigned tf = 0;
        TrigLookup<R>::get(a, heading[i], c, s, tf);
        bc.trig_fb += tf;
        out[i] = c;
        out[N + i] = s;
    }
    flush_counters(bc, a.counters);
}

// ---------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------
inline int div_up(int a, int b) { return (a + b - 1) / b; }

template <typename R> void Launch<R>::count_keys(const StepArgs<R>& a, cudaStream_t s)
{
    if (a.N > 0) k_count_keys<R><<<div_up(a.N, 256), 256, 0, s>>>(a);
}
template <typename R> void Launch<R>::reorder(const StepArgs<R>& a, cudaStream_t s)
{
    if (a.N > 0) k_reorder<R><<<div_up(a.N, 256), 256, 0, s>>>(a);
}
template <typename R> void Launch<R>::neigh_euclid(const StepArgs<R>& a, cudaStream_t s)
{
    if (a.N > 0) k_neigh_euclid<R, sizeof(R) == 8><<<div_up(a.N, 128), 128, 0, s>>>(a);
}
template <typename R> void Launch<R>::neigh_table(const StepArgs<R>& a, cudaStream_t s, int sm_count)
{
    if (a.N <= 0) return;
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
