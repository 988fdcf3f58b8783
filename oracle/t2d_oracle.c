/* oracle/t2d_oracle.c — TEST INFRASTRUCTURE, not product code (see t2d_oracle.h).
 *
 * CPU restatement of the reference's per-timestep particle update.  Every function cites the reference
 * lines it follows (paths under /root/reference).  Compiled with -ffp-contract=off so that no FMA is
 * formed: the reference is built for baseline x86-64 (no FMA) and the arithmetic must round identically.
 *
 * Parity status: PINNED — against the reference's own test vectors (tests/test_oracle_kat.py) and
 * against outputs of the compiled reference (tests/golden/ npz, tests/test_oracle_vs_reference.py).
 */
#include "t2d_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define DEG_TO_RAD (M_PI / 180.0) /* OrientationHelper.h:26, LinearAlgebra.h */
#define RAD_TO_DEG (180.0 / M_PI)
#define WRAP_CAP 4096

struct t2do_ctx {
    int V, F;
    double *uv, *x3d; /* row-major V*2, V*3 */
    int* faces;       /* F*3 */
    /* table */
    int tableV;
    double* Df64;
    uint8_t* Du8;
    /* UV grid for point location */
    int G;
    int *gstart, *gfaces;
    int lift_mode;            /* 0: the reference's lift (CellHelper.cpp:133-146); 1: true barycentric weights (t2d.h T2D_LIFT_BARYCENTRIC) */
    double* angle_out;        /* optional: receives every particle's mean angle in degrees (before truncation) */
    const double* eta_inject; /* optional per-particle noise (degrees) replacing the Philox draw; borrowed */
    /* table-mode CSR cache */
    double csr_rmax;
    int* csr_start;
    int* csr_col;
    double* csr_d;
};

/* ------------------------------------------------------------------------------------------------ */
/* small pure functions                                                                              */
/* ------------------------------------------------------------------------------------------------ */

/* SurfaceParametrization::check_point_in_polygon (MeshCartographyLib SurfaceParametrization.cpp:45-75) on
 * the square border polygon == closed unit square (pinned by MCL tests/test_SurfaceParametrization.cpp:59-123
 * and by tests/test_oracle_vs_reference.py against the compiled polygon test). */
int t2do_inside(double x, double y) { return (x >= 0.0 && x <= 1.0 && y >= 0.0 && y <= 1.0) ? 1 : 0; }

/* LinearAlgebra::angles_to_unit_vectors, LinearAlgebra.cpp:25-46 */
void t2do_angles_to_unit_vectors(int N, const int* n, double* out)
{
    for (int i = 0; i < N; ++i) {
        double angle_degrees = n[i];
        double angle_radians = angle_degrees * DEG_TO_RAD;
        out[i] = cos(angle_radians);
        out[N + i] = sin(angle_radians);
    }
}

/* OrientationHelper::mean_unit_circle_vector_angle_degrees, OrientationHelper.cpp:84-116 */
static double mean_angle_from_sum(double mx, double my)
{
    /* Eigen normalize(): z = squaredNorm(); if (z > 0) v /= sqrt(z) */
    double z = mx * mx + my * my;
    if (z > 0.0) {
        double s = sqrt(z);
        mx = mx / s;
        my = my / s;
    }
    double angle_radians = atan2(my, mx);
    double angle_degrees = angle_radians * RAD_TO_DEG;
    if (angle_degrees < 0)
        angle_degrees += 360.0;
    return angle_degrees;
}

double t2do_mean_angle_deg(const double* angles, int count, int* empty)
{
    if (empty)
        *empty = (count == 0);
    if (count == 0)
        return NAN; /* reference throws std::invalid_argument, OrientationHelper.cpp:86-89 */
    double mx = 0.0, my = 0.0;
    for (int i = 0; i < count; ++i) {
        double r = angles[i] * DEG_TO_RAD;
        mx += cos(r);
        my += sin(r);
    }
    return mean_angle_from_sum(mx, my);
}

/* ForceHelper::repulsive_adhesion_motion, ForceHelper.cpp:84-104 */
void t2do_repulsive_adhesion(double k, double sigma, double dist, double r_adh, double k_adh, double dvx, double dvy,
                             double* out)
{
    double Fij_rep = 0, Fij_adh = 0;
    if (dist < 2 * sigma)
        Fij_rep = (-k * (2 * sigma - dist)) / (2 * sigma);
    if (dist >= 2 * sigma && dist <= r_adh)
        Fij_adh = (k_adh * (2 * sigma - dist)) / (2 * sigma - r_adh);
    double Fij = Fij_rep + Fij_adh;
    out[0] = Fij * (dvx / dist);
    out[1] = Fij * (dvy / dist);
}

/* Locomotion::get_dist_vect, Locomotion.cpp:143-162: diff(i,j) = r_i - r_j */
void t2do_get_dist_vect(int N, const double* r, double* dx, double* dy)
{
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            dx[(size_t)i * N + j] = r[i] - r[j];
            dy[(size_t)i * N + j] = r[N + i] - r[N + j];
        }
}

/* Locomotion::transform_into_symmetric_matrix, Locomotion.cpp:127-138 */
void t2do_symmetrize_min(int N, double* A)
{
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j) {
            double a = A[(size_t)i * N + j], b = A[(size_t)j * N + i];
            double m = (b < a) ? b : a; /* std::min(a, b) */
            A[(size_t)i * N + j] = A[(size_t)j * N + i] = m;
        }
}

/* Philox4x32-10 (Salmon et al., SC'11), counter = (id, step_lo, step_hi, 0), key = (seed_lo, seed_hi). */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0, c[1] = n1, c[2] = n2, c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

double t2do_philox_uniform(uint64_t seed, uint64_t step, uint32_t id)
{
    uint32_t c[4] = {id, (uint32_t)step, (uint32_t)(step >> 32), 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint64_t bits = ((uint64_t)c[0] << 32) | c[1];
    return (double)(bits >> 11) * 0x1.0p-53;
}

/* eta_i = (eta*360) * (u - 0.5) degrees (SURVEY §8b RNG row) */
double t2do_noise_deg(double eta, uint64_t seed, uint64_t step, uint32_t id)
{
    double u = t2do_philox_uniform(seed, step, id);
    return (eta * 360.0) * (u - 0.5);
}

/* ------------------------------------------------------------------------------------------------ */
/* CellHelper: point-triangle distance and the UV -> 3-D lift                                        */
/* ------------------------------------------------------------------------------------------------ */

/* CellHelper::pointSegmentDistance, CellHelper.cpp:216-222, with z == 0 everywhere */
static double point_segment_distance(double px, double py, double ax, double ay, double bx, double by)
{
    double abx = bx - ax, aby = by - ay;
    double t = (abx * (px - ax) + aby * (py - ay)) / (abx * abx + aby * aby);
    t = (t < 0.0) ? 0.0 : ((1.0 < t) ? 1.0 : t); /* std::clamp */
    double ex = (ax + abx * t) - px, ey = (ay + aby * t) - py;
    return sqrt(ex * ex + ey * ey);
}

/* CellHelper::pointTriangleDistance, CellHelper.cpp:162-214 (Ericson regions), z == 0 everywhere: the
 * reference's over-read z of p is forced to 0 by the harness, and all UV vertices have z = 0. */
static double point_triangle_distance(double px, double py, double ax, double ay, double bx, double by, double cx,
                                      double cy)
{
    double abx = bx - ax, aby = by - ay;
    double acx = cx - ax, acy = cy - ay;
    double apx = px - ax, apy = py - ay;
    double bpx = px - bx, bpy = py - by;
    double cpx = px - cx, cpy = py - cy;

    double d_ab_ap = abx * apx + aby * apy;
    double d_ac_ap = acx * apx + acy * apy;
    double d_ab_bp = abx * bpx + aby * bpy;
    double d_ac_bp = acx * bpx + acy * bpy;
    double d_ab_cp = abx * cpx + aby * cpy;
    double d_ac_cp = acx * cpx + acy * cpy;

    if (d_ab_ap <= 0.0 && d_ac_ap <= 0.0)
        return sqrt(apx * apx + apy * apy);
    if (d_ab_bp >= 0.0 && d_ac_bp <= d_ab_bp)
        return sqrt(bpx * bpx + bpy * bpy);
    if (d_ac_cp >= 0.0 && d_ab_cp <= d_ac_cp)
        return sqrt(cpx * cpx + cpy * cpy);

    double vc = d_ab_ap * d_ac_bp - d_ab_bp * d_ac_ap;
    if (vc <= 0.0 && d_ab_ap >= 0.0 && d_ab_bp <= 0.0)
        return point_segment_distance(px, py, ax, ay, bx, by);

    double vb = d_ab_cp * d_ac_ap - d_ab_ap * d_ac_cp;
    if (vb <= 0.0 && d_ac_ap >= 0.0 && d_ac_cp <= 0.0)
        return point_segment_distance(px, py, ax, ay, cx, cy);

    double va = d_ab_bp * d_ac_cp - d_ab_cp * d_ac_bp;
    if (va <= 0.0 && (d_ac_bp - d_ab_bp) >= 0.0 && (d_ab_cp - d_ac_cp) >= 0.0)
        return point_segment_distance(px, py, bx, by, cx, cy);

    double denom = 1.0 / (va + vb + vc);
    double v = vb * denom;
    double w = vc * denom;
    double ex = ((ax + abx * v) + acx * w) - px;
    double ey = ((ay + aby * v) + acy * w) - py;
    return sqrt(ex * ex + ey * ey);
}

double t2do_point_triangle_distance(const double* p, const double* a, const double* b, const double* c)
{
    return point_triangle_distance(p[0], p[1], a[0], a[1], b[0], b[1], c[0], c[1]);
}

static double face_distance(const t2do_ctx* c, int f, double px, double py)
{
    const int* fv = c->faces + 3 * f;
    const double *a = c->uv + 2 * fv[0], *b = c->uv + 2 * fv[1], *cc = c->uv + 2 * fv[2];
    return point_triangle_distance(px, py, a[0], a[1], b[0], b[1], cc[0], cc[1]);
}

/* the part of CellHelper::calculate_barycentric_3D_coord after the arg-min, CellHelper.cpp:119-159 */
static void lift_to_3d(const t2do_ctx* c, int f, double px, double py, double* X, int* vid)
{
    const int* fv = c->faces + 3 * f;
    const double *ua = c->uv + 2 * fv[0], *ub = c->uv + 2 * fv[1], *uc = c->uv + 2 * fv[2];
    const double *a = c->x3d + 3 * fv[0], *b = c->x3d + 3 * fv[1], *cc = c->x3d + 3 * fv[2];
    double dax = px - ua[0], day = py - ua[1];
    double dbx = px - ub[0], dby = py - ub[1];
    double dcx = px - uc[0], dcy = py - uc[1];
    double w_a, w_b, w_c;
    if (c->lift_mode == 1) { /* extension: barycentric coordinates of p in the UV triangle */
        double den = (ub[0] - ua[0]) * (uc[1] - ua[1]) - (ub[1] - ua[1]) * (uc[0] - ua[0]);
        w_a = (dbx * dcy - dby * dcx) / den;
        w_b = (dcx * day - dcy * dax) / den;
        w_c = (1.0 - w_a) - w_b;
    } else {
        w_a = sqrt(dax * dax + day * day);
        w_b = sqrt(dbx * dbx + dby * dby);
        w_c = sqrt(dcx * dcx + dcy * dcy);
        double sum_weights = w_a + w_b + w_c;
        w_a /= sum_weights;
        w_b /= sum_weights;
        w_c /= sum_weights;
    }
    for (int k = 0; k < 3; ++k)
        X[k] = (w_a * a[k] + w_b * b[k]) + w_c * cc[k];
    double ea[3], eb[3], ec[3];
    for (int k = 0; k < 3; ++k) {
        ea[k] = X[k] - a[k];
        eb[k] = X[k] - b[k];
        ec[k] = X[k] - cc[k];
    }
    double dist_a = sqrt((ea[0] * ea[0] + ea[1] * ea[1]) + ea[2] * ea[2]);
    double dist_b = sqrt((eb[0] * eb[0] + eb[1] * eb[1]) + eb[2] * eb[2]);
    double dist_c = sqrt((ec[0] * ec[0] + ec[1] * ec[1]) + ec[2] * ec[2]);
    double m = dist_a;
    if (dist_b < m)
        m = dist_b;
    if (dist_c < m)
        m = dist_c; /* std::min({a,b,c}) */
    if (m == dist_a)
        *vid = fv[0];
    else if (m == dist_b)
        *vid = fv[1];
    else
        *vid = fv[2];
}

/* arg-min over ALL faces of (distance, face index): CellHelper.cpp:106-117 */
static int locate_brute(const t2do_ctx* c, double px, double py)
{
    int best = 0;
    double bd = face_distance(c, 0, px, py);
    for (int f = 1; f < c->F; ++f) {
        double d = face_distance(c, f, px, py);
        if (d < bd) { /* std::pair compare: smaller distance, then smaller index (ascending scan keeps it) */
            bd = d;
            best = f;
        }
    }
    return best;
}

/* uniform UV grid: cell -> ascending list of faces whose bbox (grown by 1e-7) touches the cell */
static void build_uv_grid(t2do_ctx* c)
{
    int G = (int)ceil(sqrt((double)c->F) * 2.0);
    if (G < 8)
        G = 8;
    if (G > 2048)
        G = 2048;
    c->G = G;
    const double eps = 1e-7;
    int* cnt = (int*)calloc((size_t)G * G + 1, sizeof(int));
    for (int pass = 0; pass < 2; ++pass) {
        for (int f = 0; f < c->F; ++f) {
            const int* fv = c->faces + 3 * f;
            double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
            for (int k = 0; k < 3; ++k) {
                double x = c->uv[2 * fv[k]], y = c->uv[2 * fv[k] + 1];
                if (x < x0) x0 = x;
                if (x > x1) x1 = x;
                if (y < y0) y0 = y;
                if (y > y1) y1 = y;
            }
            int i0 = (int)floor((x0 - eps) * G), i1 = (int)floor((x1 + eps) * G);
            int j0 = (int)floor((y0 - eps) * G), j1 = (int)floor((y1 + eps) * G);
            if (i0 < 0) i0 = 0;
            if (j0 < 0) j0 = 0;
            if (i1 > G - 1) i1 = G - 1;
            if (j1 > G - 1) j1 = G - 1;
            for (int j = j0; j <= j1; ++j)
                for (int i = i0; i <= i1; ++i) {
                    int cell = j * G + i;
                    if (pass == 0)
                        cnt[cell + 1]++;
                    else
                        c->gfaces[c->gstart[cell] + cnt[cell]++] = f;
                }
        }
        if (pass == 0) {
            c->gstart = (int*)malloc(((size_t)G * G + 1) * sizeof(int));
            c->gstart[0] = 0;
            for (int i = 0; i < G * G; ++i)
                c->gstart[i + 1] = c->gstart[i] + cnt[i + 1];
            c->gfaces = (int*)malloc((size_t)(c->gstart[G * G] > 0 ? c->gstart[G * G] : 1) * sizeof(int));
            memset(cnt, 0, ((size_t)G * G + 1) * sizeof(int));
        }
    }
    free(cnt);
}

static int locate_grid(const t2do_ctx* c, double px, double py, int* fallback)
{
    int G = c->G;
    int i = (int)floor(px * G), j = (int)floor(py * G);
    if (i < 0) i = 0;
    if (j < 0) j = 0;
    if (i > G - 1) i = G - 1;
    if (j > G - 1) j = G - 1;
    int cell = j * G + i;
    int best = -1;
    double bd = 0;
    for (int q = c->gstart[cell]; q < c->gstart[cell + 1]; ++q) {
        int f = c->gfaces[q];
        double d = face_distance(c, f, px, py);
        if (best < 0 || d < bd) {
            bd = d;
            best = f;
        }
    }
    if (best < 0 || !(bd <= 1e-9)) { /* not covered (or NaN): do what the reference does */
        if (fallback)
            *fallback = 1;
        return locate_brute(c, px, py);
    }
    return best;
}

int t2do_get_r3d(t2do_ctx* c, int N, const double* uv, double* r3d, int* vid, int* face, int brute, t2do_stats* st)
{
    int64_t fb = 0;
#pragma omp parallel for schedule(static) reduction(+ : fb)
    for (int i = 0; i < N; ++i) {
        double px = uv[i], py = uv[N + i];
        int fall = 0;
        int f = brute ? locate_brute(c, px, py) : locate_grid(c, px, py, &fall);
        fb += fall;
        double X[3];
        int v;
        lift_to_3d(c, f, px, py, X, &v);
        r3d[i] = X[0];
        r3d[N + i] = X[1];
        r3d[2 * N + i] = X[2];
        vid[i] = v;
        if (face)
            face[i] = f;
    }
    if (st)
        st->locate_fallbacks += fb;
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* EuclideanTiling: seam re-entry                                                                    */
/* ------------------------------------------------------------------------------------------------ */

/* EuclideanTiling::is_point_on_segment, EuclideanTiling.cpp:156-168 */
static int is_point_on_segment(double px, double py, double ax, double ay, double bx, double by)
{
    if (px < fmin(ax, bx) || px > fmax(ax, bx) || py < fmin(ay, by) || py > fmax(ay, by))
        return 0;
    double crossProduct = (px - ax) * (by - ay) - (py - ay) * (bx - ax);
    return fabs(crossProduct) < 1e-9;
}

/* EuclideanTiling::intersection_point for a 2-point border C-D, EuclideanTiling.cpp:170-208 */
static int intersection_point(double ax, double ay, double bx, double by, double cx, double cy, double dx, double dy,
                              double* ox, double* oy)
{
    if (is_point_on_segment(ax, ay, cx, cy, dx, dy))
        return 0;
    double det = (bx - ax) * (dy - cy) - (by - ay) * (dx - cx);
    if (fabs(det) < 1e-9)
        return 0;
    double t = ((cx - ax) * (dy - cy) - (cy - ay) * (dx - cx)) / det;
    double s = ((cx - ax) * (by - ay) - (cy - ay) * (bx - ax)) / det;
    if (t >= 0 && t <= 1 && s >= 0 && s <= 1) {
        *ox = ax + t * (bx - ax);
        *oy = ay + t * (by - ay);
        return 1;
    }
    return 0;
}

/* borders in the order of EuclideanTiling.cpp:79: left, right, up, down (2-point unit-square sides, the
 * harness's population of Tessellation::*_border) */
static const double BORDERS[4][4] = {{0, 0, 0, 1}, {1, 0, 1, 1}, {0, 1, 1, 1}, {0, 0, 1, 0}};

/* EuclideanTiling::check_border_crossings, EuclideanTiling.cpp:71-101: returns border id 0..3 or 4 = none */
static int check_border_crossings(double sx, double sy, double ex, double ey, double* xo, double* yo)
{
    for (int b = 0; b < 4; ++b) {
        double x, y;
        if (intersection_point(sx, sy, ex, ey, BORDERS[b][0], BORDERS[b][1], BORDERS[b][2], BORDERS[b][3], &x, &y)) {
            if (fabs(x) < 1e-3)
                x = 0.0;
            if (fabs(y) < 1e-3)
                y = 0.0;
            *xo = x;
            *yo = y;
            return b;
        }
    }
    *xo = sx;
    *yo = sy;
    return 4;
}

/* per-particle form of diagonal_seam_edges_square_border + processPoints, EuclideanTiling.cpp:31-69,112-154.
 * The reference's serial do/while restarts from particle 0 whenever a particle is still outside; already
 * settled particles are untouched by a restart, so the loop is independent per particle. */
static int tiling_one(double* oldx, double* oldy, double* px, double* py, int* n, int* wraps)
{
    for (int it = 0; it < WRAP_CAP; ++it) {
        double ax = *oldx, ay = *oldy, qx = *px, qy = *py;
        double n_double = *n;
        if (t2do_inside(qx, qy))
            return 0; /* new_point = point_outside; n unchanged */
        double ex, ey;
        int border = check_border_crossings(ax, ay, qx, qy, &ex, &ey);
        double entry_x = ey, entry_y = ex; /* entry_point = (exit[1], exit[0]) */
        double nx, ny;
        if (border == 0) {
            nx = qy, ny = -qx;
            n_double -= 90.0;
        } else if (border == 1) {
            nx = qy, ny = 2 - qx;
            n_double -= 90.0;
        } else if (border == 2) {
            nx = 2 - qy, ny = qx;
            n_double -= 270.0;
        } else {
            nx = -qy, ny = qx;
            n_double -= 270.0;
        }
        *n = (int)n_double;
        (*wraps)++;
        *px = nx;
        *py = ny;
        if (t2do_inside(nx, ny))
            return 0;
        *oldx = entry_x;
        *oldy = entry_y;
    }
    return 1; /* cap hit (the reference would loop forever) */
}

int t2do_tiling(t2do_ctx* c, int N, double* uv_old, double* uv, int* n, t2do_stats* st)
{
    (void)c;
    int64_t wraps = 0, caps = 0;
#pragma omp parallel for schedule(static) reduction(+ : wraps, caps)
    for (int i = 0; i < N; ++i) {
        int w = 0;
        caps += tiling_one(&uv_old[i], &uv_old[N + i], &uv[i], &uv[N + i], &n[i], &w);
        wraps += w;
    }
    if (st) {
        st->wraps += wraps;
        st->wrap_cap_hits += caps;
    }
    return caps ? 4 : 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* context, table                                                                                    */
/* ------------------------------------------------------------------------------------------------ */

t2do_ctx* t2do_create(int V, int F, const double* uv, const double* x3d, const int* faces)
{
    t2do_ctx* c = (t2do_ctx*)calloc(1, sizeof(t2do_ctx));
    c->V = V;
    c->F = F;
    c->uv = (double*)malloc(sizeof(double) * 2 * V);
    c->x3d = (double*)malloc(sizeof(double) * 3 * V);
    c->faces = (int*)malloc(sizeof(int) * 3 * F);
    memcpy(c->uv, uv, sizeof(double) * 2 * V);
    memcpy(c->x3d, x3d, sizeof(double) * 3 * V);
    memcpy(c->faces, faces, sizeof(int) * 3 * F);
    build_uv_grid(c);
    c->csr_rmax = -1;
    return c;
}

static void free_csr(t2do_ctx* c)
{
    free(c->csr_start);
    free(c->csr_col);
    free(c->csr_d);
    c->csr_start = c->csr_col = NULL;
    c->csr_d = NULL;
    c->csr_rmax = -1;
}

void t2do_destroy(t2do_ctx* c)
{
    if (!c)
        return;
    free(c->uv);
    free(c->x3d);
    free(c->faces);
    free(c->Df64);
    free(c->Du8);
    free(c->gstart);
    free(c->gfaces);
    free_csr(c);
    free(c);
}

int t2do_set_table_f64(t2do_ctx* c, int V, const double* D)
{
    free(c->Df64);
    free(c->Du8);
    c->Du8 = NULL;
    c->Df64 = (double*)malloc(sizeof(double) * (size_t)V * V);
    memcpy(c->Df64, D, sizeof(double) * (size_t)V * V);
    c->tableV = V;
    free_csr(c);
    return 0;
}

int t2do_set_table_u8(t2do_ctx* c, int V, const uint8_t* D)
{
    free(c->Df64);
    free(c->Du8);
    c->Df64 = NULL;
    c->Du8 = (uint8_t*)malloc((size_t)V * V);
    memcpy(c->Du8, D, (size_t)V * V);
    c->tableV = V;
    free_csr(c);
    return 0;
}

const uint8_t* t2do_table_u8(t2do_ctx* c) { return c->Du8; }

void t2do_inject_noise(t2do_ctx* c, const double* eta_deg) { c->eta_inject = eta_deg; }
void t2do_set_lift_mode(t2do_ctx* c, int mode) { c->lift_mode = mode; }

void t2do_set_angle_out(t2do_ctx* c, double* buf) { c->angle_out = buf; }

static inline double table_at(const t2do_ctx* c, int a, int b)
{
    size_t k = (size_t)a * c->tableV + b;
    return c->Df64 ? c->Df64[k] : (double)c->Du8[k];
}

/* DijkstraDistanceHelper::calculate_edge_count_distance (MeshCartographyLib DijkstraDistanceHelper.cpp:82-111):
 * Dijkstra with unit edge weights over the halfedge graph of <stem>_open.off == BFS hop count over the
 * undirected edge graph of the faces.  Values saturate at 255 (ellipsoid max is 64). */
int t2do_build_hop_table(t2do_ctx* c)
{
    const int V = c->V, F = c->F;
    /* adjacency CSR from face edges (duplicates are harmless for BFS) */
    int* deg = (int*)calloc((size_t)V + 1, sizeof(int));
    for (int f = 0; f < F; ++f)
        for (int k = 0; k < 3; ++k) {
            deg[c->faces[3 * f + k] + 1] += 2;
        }
    int* start = (int*)malloc(((size_t)V + 1) * sizeof(int));
    start[0] = 0;
    for (int v = 0; v < V; ++v)
        start[v + 1] = start[v] + deg[v + 1];
    int* fill = (int*)calloc((size_t)V, sizeof(int));
    int* adj = (int*)malloc((size_t)start[V] * sizeof(int));
    for (int f = 0; f < F; ++f)
        for (int k = 0; k < 3; ++k) {
            int a = c->faces[3 * f + k], b = c->faces[3 * f + (k + 1) % 3];
            adj[start[a] + fill[a]++] = b;
            adj[start[b] + fill[b]++] = a;
        }
    free(c->Df64);
    free(c->Du8);
    c->Df64 = NULL;
    c->Du8 = (uint8_t*)malloc((size_t)V * V);
    c->tableV = V;
    free_csr(c);
#pragma omp parallel
    {
        int* queue = (int*)malloc((size_t)V * sizeof(int));
        int* dist = (int*)malloc((size_t)V * sizeof(int));
#pragma omp for schedule(dynamic, 16)
        for (int s = 0; s < V; ++s) {
            for (int v = 0; v < V; ++v)
                dist[v] = -1;
            int head = 0, tail = 0;
            queue[tail++] = s;
            dist[s] = 0;
            while (head < tail) {
                int u = queue[head++];
                for (int q = start[u]; q < start[u + 1]; ++q) {
                    int w = adj[q];
                    if (dist[w] < 0) {
                        dist[w] = dist[u] + 1;
                        queue[tail++] = w;
                    }
                }
            }
            uint8_t* row = c->Du8 + (size_t)s * V;
            for (int v = 0; v < V; ++v)
                row[v] = (dist[v] < 0 || dist[v] > 255) ? 255 : (uint8_t)dist[v];
        }
        free(queue);
        free(dist);
    }
    free(deg);
    free(start);
    free(fill);
    free(adj);
    return 0;
}

/* per-vertex CSR of table entries that can matter: d < 2 sigma  or  d <= color_factor * sigma */
static void ensure_csr(t2do_ctx* c, double two_sigma, double color_r)
{
    double rmax = two_sigma > color_r ? two_sigma : color_r;
    if (c->csr_start && c->csr_rmax == rmax)
        return;
    free_csr(c);
    const int V = c->tableV;
    c->csr_start = (int*)malloc(((size_t)V + 1) * sizeof(int));
    size_t total = 0;
    for (int v = 0; v < V; ++v) {
        c->csr_start[v] = (int)total;
        for (int u = 0; u < V; ++u) {
            /* min(D[v][u], D[u][v]): what Locomotion.cpp:110 (symmetrise by min) leaves in dist_length */
            double d = table_at(c, v, u), dT = table_at(c, u, v);
            if (dT < d)
                d = dT;
            if (d < two_sigma || d <= color_r)
                total++;
        }
    }
    c->csr_start[V] = (int)total;
    c->csr_col = (int*)malloc((total ? total : 1) * sizeof(int));
    c->csr_d = (double*)malloc((total ? total : 1) * sizeof(double));
    size_t q = 0;
    for (int v = 0; v < V; ++v)
        for (int u = 0; u < V; ++u) {
            double d = table_at(c, v, u), dT = table_at(c, u, v);
            if (dT < d)
                d = dT;
            if (d < two_sigma || d <= color_r) {
                c->csr_col[q] = u;
                c->csr_d[q] = d;
                q++;
            }
        }
    c->csr_rmax = rmax;
}

/* ------------------------------------------------------------------------------------------------ */
/* the step                                                                                          */
/* ------------------------------------------------------------------------------------------------ */

typedef struct {
    int j;
    double d;
} nb_t;

static int nb_cmp(const void* a, const void* b)
{
    int x = ((const nb_t*)a)->j, y = ((const nb_t*)b)->j;
    return (x > y) - (x < y);
}

/* Everything the reference derives from one row i of dist_length, given that row's in-range entries in
 * ascending j (the reference scans j = 0..N-1):
 *   ForceHelper::calculate_forces_between_particles  ForceHelper.cpp:34-73  (+ repulsive_adhesion_motion :84-104)
 *   OrientationHelper::calculate_average_n_within_distance  OrientationHelper.cpp:29-74
 *   _2DTissue::count_particle_neighbors  2DTissue.cpp:254-268
 * list must contain every j with d_ij < 2 sigma or (d_ij != 0 and d_ij <= color_r), including j == i (d = 0). */
static void row_physics(int i, const nb_t* list, int cnt, const double* uv, int N, const int* n_old, double k,
                        double sigma, double color_r, double* Fx, double* Fy, double* angle_deg, int* color,
                        int64_t* pairs, int64_t* ties_cut)
{
    const double two_sigma = 2 * sigma;
    double fx = 0.0, fy = 0.0, mx = 0.0, my = 0.0;
    int col = 0;
    for (int q = 0; q < cnt; ++q) {
        int j = list[q].j;
        double dist = list[q].d;
        if (dist != 0 && dist <= color_r)
            col += 1;
        if (dist == two_sigma)
            (*ties_cut)++;
        if (!(dist < two_sigma))
            continue;
        /* orientation: includes j == i */
        double r = (double)n_old[j] * DEG_TO_RAD;
        mx += cos(r);
        my += sin(r);
        if (j == i)
            continue;
        (*pairs)++;
        if (dist == 0)
            dist += 0.001;
        double Fij = (-k * (two_sigma - dist)) / (two_sigma);
        double dvx = uv[i] - uv[j], dvy = uv[N + i] - uv[N + j];
        fx += Fij * (dvx / dist);
        fy += Fij * (dvy / dist);
    }
    *Fx = fx;
    *Fy = fy;
    *color = col;
    *angle_deg = mean_angle_from_sum(mx, my);
}

static inline uint32_t cell_hash(int64_t cx, int64_t cy, int64_t cz, uint32_t mask)
{
    uint64_t h = (uint64_t)cx * 0x9E3779B97F4A7C15ull ^ (uint64_t)cy * 0xC2B2AE3D27D4EB4Full ^
                 (uint64_t)cz * 0x165667B19E3779F9ull;
    h ^= h >> 29;
    return (uint32_t)h & mask;
}

int t2do_step(t2do_ctx* c, const t2do_params* P, int N, double* uv, int* n, int* vid, double* r3d, const uint32_t* ids,
              uint64_t step_index, double* rdot, int* color, double* Fout, int* face, t2do_stats* st)
{
    const double sigma = P->sigma, two_sigma = 2 * sigma, color_r = P->color_factor * sigma;
    const double rmax = two_sigma > color_r ? two_sigma : color_r;
    int64_t pairs = 0, ties_cut = 0, ties_trunc = 0;
#ifdef _OPENMP
    if (P->threads > 0)
        omp_set_num_threads(P->threads);
#endif
    if (P->mode == 0 && !c->Df64 && !c->Du8)
        return -2;

    double* Fx = (double*)malloc(sizeof(double) * N);
    double* Fy = (double*)malloc(sizeof(double) * N);
    double* ang = (double*)malloc(sizeof(double) * N);
    double* uv_old = (double*)malloc(sizeof(double) * 2 * N);
    memcpy(uv_old, uv, sizeof(double) * 2 * N); /* r_UV_old = r_UV, 2DTissue.cpp:139 */

    /* ---- stage 1-3: distances (Locomotion.cpp:94-111) -> force, alignment, colour ---- */
    if (P->brute) {
#pragma omp parallel reduction(+ : pairs, ties_cut)
        {
            nb_t* list = (nb_t*)malloc(sizeof(nb_t) * (size_t)N);
#pragma omp for schedule(dynamic, 16)
            for (int i = 0; i < N; ++i) {
                int cnt = 0;
                for (int j = 0; j < N; ++j) {
                    double d;
                    if (i == j)
                        d = 0.0; /* dist_length.diagonal() = 0, Locomotion.cpp:109 */
                    else if (P->mode == 0) {
                        /* symmetrise-by-min (Locomotion.cpp:110) of a gather from a symmetric table */
                        double a = table_at(c, vid[i], vid[j]), b = table_at(c, vid[j], vid[i]);
                        d = (b < a) ? b : a;
                    } else {
                        double dx = r3d[i] - r3d[j], dy = r3d[N + i] - r3d[N + j], dz = r3d[2 * N + i] - r3d[2 * N + j];
                        d = sqrt(dx * dx + dy * dy + dz * dz);
                    }
                    if (d < two_sigma || (d != 0 && d <= color_r)) {
                        list[cnt].j = j;
                        list[cnt].d = d;
                        cnt++;
                    }
                }
                row_physics(i, list, cnt, uv_old, N, n, P->k, sigma, color_r, &Fx[i], &Fy[i], &ang[i], &color[i], &pairs,
                            &ties_cut);
            }
            free(list);
        }
    } else if (P->mode == 0) {
        /* bucket particles by nearest-vertex id; neighbours of bucket v = buckets u with D[v][u] in range */
        ensure_csr(c, two_sigma, color_r);
        const int V = c->tableV;
        int* bstart = (int*)calloc((size_t)V + 1, sizeof(int));
        int* bitems = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
        for (int i = 0; i < N; ++i)
            bstart[vid[i] + 1]++;
        for (int v = 0; v < V; ++v)
            bstart[v + 1] += bstart[v];
        int* fill = (int*)calloc((size_t)V, sizeof(int));
        for (int i = 0; i < N; ++i) /* ascending i inside each bucket */
            bitems[bstart[vid[i]] + fill[vid[i]]++] = i;
        free(fill);
#pragma omp parallel reduction(+ : pairs, ties_cut)
        {
            size_t cap = 1024;
            nb_t* list = (nb_t*)malloc(sizeof(nb_t) * cap);
#pragma omp for schedule(dynamic, 8)
            for (int v = 0; v < V; ++v) {
                if (bstart[v + 1] == bstart[v])
                    continue;
                size_t cnt = 0;
                for (int q = c->csr_start[v]; q < c->csr_start[v + 1]; ++q) {
                    int u = c->csr_col[q];
                    double d = c->csr_d[q];
                    for (int t = bstart[u]; t < bstart[u + 1]; ++t) {
                        if (cnt == cap) {
                            cap *= 2;
                            list = (nb_t*)realloc(list, sizeof(nb_t) * cap);
                        }
                        list[cnt].j = bitems[t];
                        list[cnt].d = d;
                        cnt++;
                    }
                }
                qsort(list, cnt, sizeof(nb_t), nb_cmp);
                for (int t = bstart[v]; t < bstart[v + 1]; ++t) {
                    int i = bitems[t];
                    /* the diagonal is forced to 0 (Locomotion.cpp:109); same-bucket j != i keep D[v][v] */
                    for (size_t q = 0; q < cnt; ++q)
                        if (list[q].j == i) {
                            list[q].d = 0.0;
                            break;
                        }
                    row_physics(i, list, (int)cnt, uv_old, N, n, P->k, sigma, color_r, &Fx[i], &Fy[i], &ang[i], &color[i],
                                &pairs, &ties_cut);
                    for (size_t q = 0; q < cnt; ++q)
                        if (list[q].j == i) {
                            list[q].d = table_at(c, v, v);
                            break;
                        }
                }
            }
            free(list);
        }
        free(bstart);
        free(bitems);
    } else {
        /* hashed 3-D cell list, cell edge >= rmax, 27-cell stencil */
        const double cs = rmax * (1.0 + 1e-9) + 1e-300;
        double mn[3] = {1e300, 1e300, 1e300};
        for (int i = 0; i < N; ++i)
            for (int k = 0; k < 3; ++k)
                if (r3d[k * N + i] < mn[k])
                    mn[k] = r3d[k * N + i];
        uint32_t M = 1;
        while (M < (uint32_t)(2 * (N > 0 ? N : 1)))
            M <<= 1;
        uint32_t mask = M - 1;
        int64_t* cell = (int64_t*)malloc(sizeof(int64_t) * 3 * (size_t)(N > 0 ? N : 1));
        int* hstart = (int*)calloc((size_t)M + 1, sizeof(int));
        int* hitems = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
        uint32_t* key = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(N > 0 ? N : 1));
        for (int i = 0; i < N; ++i) {
            for (int k = 0; k < 3; ++k)
                cell[3 * (size_t)i + k] = (int64_t)floor((r3d[k * N + i] - mn[k]) / cs);
            key[i] = cell_hash(cell[3 * (size_t)i], cell[3 * (size_t)i + 1], cell[3 * (size_t)i + 2], mask);
            hstart[key[i] + 1]++;
        }
        for (uint32_t h = 0; h < M; ++h)
            hstart[h + 1] += hstart[h];
        int* fill = (int*)calloc((size_t)M, sizeof(int));
        for (int i = 0; i < N; ++i)
            hitems[hstart[key[i]] + fill[key[i]]++] = i;
        free(fill);
#pragma omp parallel reduction(+ : pairs, ties_cut)
        {
            size_t cap = 256;
            nb_t* list = (nb_t*)malloc(sizeof(nb_t) * cap);
#pragma omp for schedule(dynamic, 256)
            for (int i = 0; i < N; ++i) {
                size_t cnt = 0;
                const int64_t* ci = cell + 3 * (size_t)i;
                for (int dz = -1; dz <= 1; ++dz)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            int64_t cx = ci[0] + dx, cy = ci[1] + dy, cz = ci[2] + dz;
                            uint32_t h = cell_hash(cx, cy, cz, mask);
                            for (int t = hstart[h]; t < hstart[h + 1]; ++t) {
                                int j = hitems[t];
                                const int64_t* cj = cell + 3 * (size_t)j;
                                if (cj[0] != cx || cj[1] != cy || cj[2] != cz)
                                    continue; /* hash collision: belongs to another cell */
                                double d;
                                if (i == j)
                                    d = 0.0;
                                else {
                                    double ex = r3d[i] - r3d[j], ey = r3d[N + i] - r3d[N + j],
                                           ez = r3d[2 * N + i] - r3d[2 * N + j];
                                    d = sqrt(ex * ex + ey * ey + ez * ez);
                                }
                                if (d < two_sigma || (d != 0 && d <= color_r)) {
                                    if (cnt == cap) {
                                        cap *= 2;
                                        list = (nb_t*)realloc(list, sizeof(nb_t) * cap);
                                    }
                                    list[cnt].j = j;
                                    list[cnt].d = d;
                                    cnt++;
                                }
                            }
                        }
                qsort(list, cnt, sizeof(nb_t), nb_cmp);
                row_physics(i, list, (int)cnt, uv_old, N, n, P->k, sigma, color_r, &Fx[i], &Fy[i], &ang[i], &color[i],
                            &pairs, &ties_cut);
            }
            free(list);
        }
        free(cell);
        free(hstart);
        free(hitems);
        free(key);
    }

    /* ---- stage 3-5: speed, Euler (Locomotion.cpp:71-84), new heading (OrientationHelper.cpp:67-70) ---- */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        double abs_F = sqrt(Fx[i] * Fx[i] + Fy[i] * Fy[i]); /* F_track.rowwise().norm() */
        double angle_radians = (double)n[i] * DEG_TO_RAD;   /* angles_to_unit_vectors(n) with the OLD n */
        double cx = cos(angle_radians), sy = sin(angle_radians);
        abs_F = abs_F + P->v0;
        double rx = cx * abs_F, ry = sy * abs_F;
        rdot[i] = rx;
        rdot[N + i] = ry;
        uv[i] = uv[i] + rx * P->step_size;
        uv[N + i] = uv[N + i] + ry * P->step_size;
        if (Fout) {
            Fout[i] = Fx[i];
            Fout[N + i] = Fy[i];
        }
    }
    for (int i = 0; i < N; ++i) { /* separate loop: n_old must stay intact while rows are evaluated above */
        double a = ang[i];
        if (c->angle_out)
            c->angle_out[i] = a;
        if (fabs(a - nearbyint(a)) < 1e-9)
            ties_trunc++;
        int avg = (int)a; /* avg_n(i) = mean angle: double -> int truncation */
        if (c->eta_inject) {
            avg = (int)((double)avg + c->eta_inject[i]);
        } else if (P->eta != 0.0) {
            uint32_t id = ids ? ids[i] : (uint32_t)i;
            double e = t2do_noise_deg(P->eta, P->seed, step_index, id);
            avg = (int)((double)avg + e); /* avg_n(i) += noise: int += double */
        }
        n[i] = avg;
    }

    /* ---- stage 6: seam re-entry; stage 7: projection; stage 8: checks ---- */
    t2do_stats local;
    memset(&local, 0, sizeof(local));
    int fault = t2do_tiling(c, N, uv_old, uv, n, &local);
    int* ftmp = face ? face : (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    t2do_get_r3d(c, N, uv, r3d, vid, ftmp, P->brute, &local);
    if (!face)
        free(ftmp);
    int64_t lost = 0, bad = 0;
    for (int i = 0; i < N; ++i) {
        if (!t2do_inside(uv[i], uv[N + i]))
            lost++;
        if (!isfinite(uv[i]) || !isfinite(uv[N + i]))
            bad++;
    }
    if (lost)
        fault |= 1;
    if (bad)
        fault |= 2;
    if (st) {
        st->pairs_in_range += pairs;
        st->ties_cutoff += ties_cut;
        st->ties_trunc += ties_trunc;
        st->wraps += local.wraps;
        st->wrap_cap_hits += local.wrap_cap_hits;
        st->locate_fallbacks += local.locate_fallbacks;
        st->lost += lost;
        st->nonfinite += bad;
    }
    free(Fx);
    free(Fy);
    free(ang);
    free(uv_old);
    return fault;
}

/* phi = |sum_i (cos n_i, sin n_i)| / N ; mean speed = <|rdot_i|> */
void t2do_observables(int N, const int* n, const double* rdot, double* out)
{
    double sx = 0, sy = 0, sp = 0;
    for (int i = 0; i < N; ++i) {
        double r = (double)n[i] * DEG_TO_RAD;
        sx += cos(r);
        sy += sin(r);
        sp += sqrt(rdot[i] * rdot[i] + rdot[N + i] * rdot[N + i]);
    }
    out[0] = N ? sqrt(sx * sx + sy * sy) / N : 0.0;
    out[1] = N ? sp / N : 0.0;
}
