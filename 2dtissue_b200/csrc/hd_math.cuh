// hd_math.cuh — per-particle arithmetic of the 2DTissue step as __host__ __device__ templates.
//
// Real = double is the parity path: it is compiled with --fmad=false (device) / -ffp-contract=off (host
// self-test) and every expression keeps the reference's operation order, so results are bit-identical to
// the reference's Eigen code built for baseline x86-64.  Real = float is the fast path (FMA allowed).
// Each function cites the reference lines it replaces (paths under /root/reference).
#pragma once
#include <math.h>
#include <stdint.h>

#include "cr_tables.h"

#if defined(__CUDACC__)
#define T2D_HD __host__ __device__ __forceinline__
#else
#define T2D_HD inline
#endif

// products and sums that must round separately even in a translation unit compiled with FMA contraction
#if defined(__CUDA_ARCH__)
#define T2D_DMUL(a, b) __dmul_rn((a), (b))
#define T2D_DADD(a, b) __dadd_rn((a), (b))
#else
#define T2D_DMUL(a, b) ((a) * (b))
#define T2D_DADD(a, b) ((a) + (b))
#endif

namespace t2d {

constexpr int WRAP_CAP = 4096;            // seam re-entry rounds before T2D_FAULT_WRAP_CAP (the reference loops until inside)
constexpr int TRIG_MIN = -3600;            // host-built cos/sin table covers integer degrees [TRIG_MIN, TRIG_MAX]
constexpr int TRIG_MAX = 1079;
constexpr int TRIG_N = TRIG_MAX - TRIG_MIN + 1;
constexpr double DEG_TO_RAD_D = 3.14159265358979323846 / 180.0;  // M_PI / 180.0, OrientationHelper.h:26
constexpr double RAD_TO_DEG_D = 180.0 / 3.14159265358979323846;

template <typename R> struct Vec2 { R x, y; };

template <typename R> T2D_HD R rsqrt_exact(R v);
template <> T2D_HD double rsqrt_exact<double>(double v) { return sqrt(v); }
template <> T2D_HD float rsqrt_exact<float>(float v) { return sqrtf(v); }
template <typename R> T2D_HD R rabs(R v) { return v < R(0) ? -v : v; }
template <typename R> T2D_HD R rmin(R a, R b) { return b < a ? b : a; }   // std::min
template <typename R> T2D_HD R rmax(R a, R b) { return a < b ? b : a; }   // std::max

// SurfaceParametrization::check_point_in_polygon on the square border == closed unit square
// (MeshCartographyLib SurfaceParametrization.cpp:45-75; equivalence pinned in tests/golden/inside_pins.npz)
template <typename R> T2D_HD bool inside_square(R x, R y) { return x >= R(0) && x <= R(1) && y >= R(0) && y <= R(1); }

// ---------------------------------------------------------------------------------------------------
// CellHelper::pointSegmentDistance / pointTriangleDistance, CellHelper.cpp:162-222, all z == 0
// ---------------------------------------------------------------------------------------------------
template <typename R> T2D_HD R point_segment_distance(R px, R py, R ax, R ay, R bx, R by)
{
    R abx = bx - ax, aby = by - ay;
    R t = (abx * (px - ax) + aby * (py - ay)) / (abx * abx + aby * aby);
    t = (t < R(0)) ? R(0) : ((R(1) < t) ? R(1) : t);
    R ex = (ax + abx * t) - px, ey = (ay + aby * t) - py;
    return rsqrt_exact<R>(ex * ex + ey * ey);
}

template <typename R> T2D_HD R point_triangle_distance(R px, R py, R ax, R ay, R bx, R by, R cx, R cy)
{
    R abx = bx - ax, aby = by - ay;
    R acx = cx - ax, acy = cy - ay;
    R apx = px - ax, apy = py - ay;
    R bpx = px - bx, bpy = py - by;
    R cpx = px - cx, cpy = py - cy;

    R d_ab_ap = abx * apx + aby * apy;
    R d_ac_ap = acx * apx + acy * apy;
    R d_ab_bp = abx * bpx + aby * bpy;
    R d_ac_bp = acx * bpx + acy * bpy;
    R d_ab_cp = abx * cpx + aby * cpy;
    R d_ac_cp = acx * cpx + acy * cpy;

    if (d_ab_ap <= R(0) && d_ac_ap <= R(0)) return rsqrt_exact<R>(apx * apx + apy * apy);
    if (d_ab_bp >= R(0) && d_ac_bp <= d_ab_bp) return rsqrt_exact<R>(bpx * bpx + bpy * bpy);
    if (d_ac_cp >= R(0) && d_ab_cp <= d_ac_cp) return rsqrt_exact<R>(cpx * cpx + cpy * cpy);

    R vc = d_ab_ap * d_ac_bp - d_ab_bp * d_ac_ap;
    if (vc <= R(0) && d_ab_ap >= R(0) && d_ab_bp <= R(0)) return point_segment_distance<R>(px, py, ax, ay, bx, by);
    R vb = d_ab_cp * d_ac_ap - d_ab_ap * d_ac_cp;
    if (vb <= R(0) && d_ac_ap >= R(0) && d_ac_cp <= R(0)) return point_segment_distance<R>(px, py, ax, ay, cx, cy);
    R va = d_ab_bp * d_ac_cp - d_ab_cp * d_ac_bp;
    if (va <= R(0) && (d_ac_bp - d_ab_bp) >= R(0) && (d_ab_cp - d_ac_cp) >= R(0))
        return point_segment_distance<R>(px, py, bx, by, cx, cy);

    R denom = R(1) / (va + vb + vc);
    R v = vb * denom;
    R w = vc * denom;
    R ex = ((ax + abx * v) + acx * w) - px;
    R ey = ((ay + aby * v) + acy * w) - py;
    return rsqrt_exact<R>(ex * ex + ey * ey);
}

// ---------------------------------------------------------------------------------------------------
// UV -> 3-D lift after the arg-min, CellHelper::calculate_barycentric_3D_coord, CellHelper.cpp:119-159.
// ua/ub/uc: UV corners; A/B/C: 3-D corners; returns which corner (0,1,2) is nearest to the lifted point.
// ---------------------------------------------------------------------------------------------------
// barycentric = false: the reference's weights (normalised UV distances to the corners); true: T2D_LIFT_BARYCENTRIC,
// w_a = [(b-p) x (c-p)] / [(b-a) x (c-a)], w_b = [(c-p) x (a-p)] / same, w_c = 1 - w_a - w_b (same order in the oracle)
template <typename R>
T2D_HD int lift_to_3d(R px, R py, R uax, R uay, R ubx, R uby, R ucx, R ucy, const R* A, const R* B, const R* C, R* X,
                      bool barycentric = false)
{
    R dax = px - uax, day = py - uay;
    R dbx = px - ubx, dby = py - uby;
    R dcx = px - ucx, dcy = py - ucy;
    R w_a, w_b, w_c;
    if (barycentric) {
        R den = (ubx - uax) * (ucy - uay) - (uby - uay) * (ucx - uax);
        w_a = (dbx * dcy - dby * dcx) / den;
        w_b = (dcx * day - dcy * dax) / den;
        w_c = (R(1) - w_a) - w_b;
    } else {
        w_a = rsqrt_exact<R>(dax * dax + day * day);
        w_b = rsqrt_exact<R>(dbx * dbx + dby * dby);
        w_c = rsqrt_exact<R>(dcx * dcx + dcy * dcy);
        R sum_weights = w_a + w_b + w_c;
        w_a = w_a / sum_weights;
        w_b = w_b / sum_weights;
        w_c = w_c / sum_weights;
    }
    R da2 = 0, db2 = 0, dc2 = 0;
    R e[3][3];
    for (int k = 0; k < 3; ++k) {
        X[k] = (w_a * A[k] + w_b * B[k]) + w_c * C[k];
        e[0][k] = X[k] - A[k];
        e[1][k] = X[k] - B[k];
        e[2][k] = X[k] - C[k];
    }
    da2 = (e[0][0] * e[0][0] + e[0][1] * e[0][1]) + e[0][2] * e[0][2];
    db2 = (e[1][0] * e[1][0] + e[1][1] * e[1][1]) + e[1][2] * e[1][2];
    dc2 = (e[2][0] * e[2][0] + e[2][1] * e[2][1]) + e[2][2] * e[2][2];
    R dist_a = rsqrt_exact<R>(da2), dist_b = rsqrt_exact<R>(db2), dist_c = rsqrt_exact<R>(dc2);
    R m = dist_a;
    if (dist_b < m) m = dist_b;
    if (dist_c < m) m = dist_c;
    if (m == dist_a) return 0;
    if (m == dist_b) return 1;
    return 2;
}

// ---------------------------------------------------------------------------------------------------
// Seam re-entry: EuclideanTiling::diagonal_seam_edges_square_border + processPoints +
// check_border_crossings + intersection_point + is_point_on_segment, EuclideanTiling.cpp:31-208,
// with Tessellation's borders = the 4 sides of the unit square as 2-point segments in the order
// left, right, up, down (EuclideanTiling.cpp:79).  The reference's serial restart loop is independent per
// particle, so it is evaluated per particle.
// ---------------------------------------------------------------------------------------------------
template <typename R> T2D_HD bool is_point_on_segment(R px, R py, R ax, R ay, R bx, R by)
{
    if (px < rmin(ax, bx) || px > rmax(ax, bx) || py < rmin(ay, by) || py > rmax(ay, by)) return false;
    R crossProduct = (px - ax) * (by - ay) - (py - ay) * (bx - ax);
    return rabs(crossProduct) < R(1e-9);
}

template <typename R> T2D_HD bool border_intersection(R ax, R ay, R bx, R by, R cx, R cy, R dx, R dy, R* ox, R* oy)
{
    if (is_point_on_segment<R>(ax, ay, cx, cy, dx, dy)) return false;
    R det = (bx - ax) * (dy - cy) - (by - ay) * (dx - cx);
    if (rabs(det) < R(1e-9)) return false;
    R t = ((cx - ax) * (dy - cy) - (cy - ay) * (dx - cx)) / det;
    R s = ((cx - ax) * (by - ay) - (cy - ay) * (bx - ax)) / det;
    if (t >= R(0) && t <= R(1) && s >= R(0) && s <= R(1)) {
        *ox = ax + t * (bx - ax);
        *oy = ay + t * (by - ay);
        return true;
    }
    return false;
}

// returns 0 left, 1 right, 2 up, 3 down, 4 no intersection; (xo, yo) = exit point (snapped) or the start
template <typename R> T2D_HD int check_border_crossings(R sx, R sy, R ex, R ey, R* xo, R* yo)
{
    const R B[4][4] = {{0, 0, 0, 1}, {1, 0, 1, 1}, {0, 1, 1, 1}, {0, 0, 1, 0}};
    for (int b = 0; b < 4; ++b) {
        R x, y;
        if (border_intersection<R>(sx, sy, ex, ey, B[b][0], B[b][1], B[b][2], B[b][3], &x, &y)) {
            if (rabs(x) < R(1e-3)) x = R(0);   // BORDER_THRESHOLD, EuclideanTiling.h
            if (rabs(y) < R(1e-3)) y = R(0);
            *xo = x;
            *yo = y;
            return b;
        }
    }
    *xo = sx;
    *yo = sy;
    return 4;
}

// in/out: (oldx, oldy) = r_UV_old, (px, py) = r_UV, n = heading.  Returns true if WRAP_CAP was hit.
template <typename R> T2D_HD bool seam_reentry(R& oldx, R& oldy, R& px, R& py, int& n, int& wraps)
{
    for (int it = 0; it < WRAP_CAP; ++it) {
        if (inside_square<R>(px, py)) return false;
        R ex, ey;
        int border = check_border_crossings<R>(oldx, oldy, px, py, &ex, &ey);
        double n_double = (double)n;
        R nx, ny;
        if (border == 0) { nx = py; ny = -px; n_double -= 90.0; }
        else if (border == 1) { nx = py; ny = R(2) - px; n_double -= 90.0; }
        else if (border == 2) { nx = R(2) - py; ny = px; n_double -= 270.0; }
        else { nx = -py; ny = px; n_double -= 270.0; }
        n = (int)n_double;
        ++wraps;
        px = nx;
        py = ny;
        if (inside_square<R>(nx, ny)) return false;
        oldx = ey;   // entry_point = (exit[1], exit[0])
        oldy = ex;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10, counter = (id, step_lo, step_hi, 0), key = (seed_lo, seed_hi)  (same in oracle/t2d_oracle.c)
// ---------------------------------------------------------------------------------------------------
T2D_HD double philox_uniform(uint64_t seed, uint64_t step, uint32_t id)
{
    uint32_t c0 = id, c1 = (uint32_t)step, c2 = (uint32_t)(step >> 32), c3 = 0u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    uint64_t bits = ((uint64_t)c0 << 32) | c1;
    return (double)(bits >> 11) * 0x1.0p-53;
}

// eta_i = (eta*360) * (u - 0.5); the two products are rounded separately on every path (no FMA possible:
// (u - 0.5) is a separate subtraction and the result is added to an integer afterwards by the caller)
T2D_HD double noise_deg(double eta360, uint64_t seed, uint64_t step, uint32_t id)
{
    double u = philox_uniform(seed, step, id);
    return T2D_DMUL(eta360, T2D_DADD(u, -0.5));
}

// ---------------------------------------------------------------------------------------------------
// Mean angle in degrees exactly as OrientationHelper::mean_unit_circle_vector_angle_degrees produces it
// (OrientationHelper.cpp:102-116):  a = atan2(my, mx) * RAD_TO_DEG;  if (a < 0) a += 360.
// The caller truncates a to int, and for aligned or isolated particles the true angle IS an integer degree,
// so the last ulp of atan2 decides the heading.  glibc's atan2 is correctly rounded on these inputs (checked
// exhaustively over the tie family in tests/test_hd_math_cpu.py); CUDA's is only 2-ulp.  So atan2 is rebuilt
// here correctly rounded, without calling atan2 in double at all:
//     theta = m*pi/180 + eps,  m = nearest integer degree (from a float atan2f),
//     tan(eps) = (y*C - x*S) / (x*C + y*S)  with C,S = cos/sin(m deg) as double-doubles (cr_tables.h),
// the numerator evaluated with error-free products (fma) so that eps is good to ~1e-32 absolute, then
// theta = fl(p_hi + fl(p_lo + eps)).  All later operations are plain IEEE double ops, identical on both sides.
// `tab` = kCrTable (host) or its device copy.  *tie is set when |a - rint(a)| < 1e-9 (diagnostic counter).
// ---------------------------------------------------------------------------------------------------
T2D_HD double mean_angle_degrees_cr(double mx, double my, const CrEntry* tab, bool* tie)
{
    double theta;
    if (mx == 0.0 && my == 0.0) {
        theta = atan2(my, mx);   // IEEE special cases (+-0, +-pi)
    } else {
        const double ay = fabs(my);
        float af = atan2f((float)ay, (float)mx) * 57.29577951308232f;
        int m = (int)rintf(af);
        m = m < 0 ? 0 : (m > 180 ? 180 : m);
        const CrEntry e = tab[m];
        const double p1 = T2D_DMUL(ay, e.c_hi), e1 = fma(ay, e.c_hi, -p1);
        const double p2 = T2D_DMUL(mx, e.s_hi), e2 = fma(mx, e.s_hi, -p2);
        const double hi = T2D_DADD(p1, -p2);                     // exact (Sterbenz) wherever it matters
        const double bb = T2D_DADD(hi, -p1);
        const double er = T2D_DADD(T2D_DADD(p1, -T2D_DADD(hi, -bb)), T2D_DADD(-p2, -bb));   // two-sum remainder elsewhere
        const double lo = T2D_DADD(T2D_DADD(T2D_DADD(T2D_DADD(e1, -e2), T2D_DMUL(ay, e.c_lo)), -T2D_DMUL(mx, e.s_lo)), er);
        const double num = T2D_DADD(hi, lo);
        const double den = T2D_DADD(T2D_DMUL(mx, e.c_hi), T2D_DMUL(ay, e.s_hi));
        const double tt = num / den;                             // tan(eps), |eps| < 0.01
        const double t2 = T2D_DMUL(tt, tt);
        const double ser = T2D_DADD(1.0 / 3.0, -T2D_DMUL(t2, 0.2));
        const double eps = T2D_DADD(tt, -T2D_DMUL(T2D_DMUL(tt, t2), ser));
        const double s = T2D_DADD(e.p_lo, eps);
        theta = T2D_DADD(e.p_hi, s);
        if (signbit(my)) theta = -theta;
    }
    double a = T2D_DMUL(theta, RAD_TO_DEG_D);
    if (a < 0) a = T2D_DADD(a, 360.0);
    if (tie) *tie = fabs(a - rint(a)) < 1e-9;
    return a;
}

// ---------------------------------------------------------------------------------------------------
// Pair force magnitude, ForceHelper::repulsive_adhesion_motion, ForceHelper.cpp:84-104 (the adhesion branch
// is unreachable from calculate_forces_between_particles because of the `dist >= 2σ -> continue` above it)
// ---------------------------------------------------------------------------------------------------
template <typename R> T2D_HD R pair_fij(R k, R two_sigma, R dist) { return (-k * (two_sigma - dist)) / two_sigma; }

}  // namespace t2d
