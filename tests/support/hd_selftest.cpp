// tests/support/hd_selftest.cpp — TEST SUPPORT, not a product path.
// Compiles the product's __host__ __device__ arithmetic (2dtissue_b200/csrc/hd_math.cuh) for the CPU so that
// the no-GPU test tier can check it against the golden vectors before the kernels ever run on a B200.
// Nothing in the package loads this library.
#include "../../2dtissue_b200/csrc/hd_math.cuh"

extern "C" {
double hd_point_triangle_distance(const double* p, const double* a, const double* b, const double* c)
{
    return t2d::point_triangle_distance<double>(p[0], p[1], a[0], a[1], b[0], b[1], c[0], c[1]);
}
int hd_lift(const double* p, const double* ua, const double* ub, const double* uc, const double* A, const double* B,
            const double* C, double* X)
{
    return t2d::lift_to_3d<double>(p[0], p[1], ua[0], ua[1], ub[0], ub[1], uc[0], uc[1], A, B, C, X);
}
// column-major arrays like the reference; returns number of cap hits
int hd_seam(int N, double* uv_old, double* uv, int* n, long long* wraps_out)
{
    int caps = 0;
    long long wraps = 0;
    for (int i = 0; i < N; ++i) {
        int w = 0;
        caps += t2d::seam_reentry<double>(uv_old[i], uv_old[N + i], uv[i], uv[N + i], n[i], w) ? 1 : 0;
        wraps += w;
    }
    if (wraps_out) *wraps_out = wraps;
    return caps;
}
int hd_seam_f32(int N, float* uv_old, float* uv, int* n)
{
    int caps = 0;
    for (int i = 0; i < N; ++i) {
        int w = 0;
        caps += t2d::seam_reentry<float>(uv_old[i], uv_old[N + i], uv[i], uv[N + i], n[i], w) ? 1 : 0;
    }
    return caps;
}
double hd_philox_uniform(unsigned long long seed, unsigned long long step, unsigned id) { return t2d::philox_uniform(seed, step, id); }
double hd_noise_deg(double eta360, unsigned long long seed, unsigned long long step, unsigned id)
{
    return t2d::noise_deg(eta360, seed, step, id);
}
double hd_pair_fij(double k, double two_sigma, double dist) { return t2d::pair_fij<double>(k, two_sigma, dist); }
// heading of a neighbour-set sum exactly as finish_particle computes it on the fp64 path
double hd_mean_angle_cr(double mx, double my)
{
    double z = mx * mx + my * my;
    if (z > 0.0) {
        double sq = sqrt(z);
        mx = mx / sq;
        my = my / sq;
    }
    bool tie;
    return t2d::mean_angle_degrees_cr(mx, my, t2d::kCrTable, &tie);
}
// glibc pipeline (what the reference executes, OrientationHelper.cpp:102-116) vs the correctly-rounded rebuild.
// family 0: k copies of heading n (aligned flock / isolated particle); family 1: symmetric pair {n-a, n+a} plus
// k copies of n; family 2: pseudo-random sets.  Returns the number of sets whose truncated heading differs;
// *n_sets, *n_ties (|angle - rint| < 1e-9 in the glibc pipeline) and *n_val (angle doubles differ) are filled.
static void hd_sum_norm(const int* h, int cnt, double& mx, double& my)
{
    mx = 0; my = 0;
    for (int q = 0; q < cnt; ++q) {
        double r = (double)h[q] * t2d::DEG_TO_RAD_D;
        mx += cos(r);
        my += sin(r);
    }
    double z = mx * mx + my * my;
    if (z > 0.0) {
        double sq = sqrt(z);
        mx = mx / sq;
        my = my / sq;
    }
}
long long hd_cr_selfcheck(int family, long long* n_sets, long long* n_ties, long long* n_val)
{
    long long bad = 0, sets = 0, ties = 0, val = 0;
    int h[80];
    unsigned long long lcg = 12345;
    auto check = [&](int cnt) {
        double mx, my;
        hd_sum_norm(h, cnt, mx, my);
        double a = atan2(my, mx) * t2d::RAD_TO_DEG_D;
        if (a < 0) a += 360.0;
        bool tie;
        double b = t2d::mean_angle_degrees_cr(mx, my, t2d::kCrTable, &tie);
        sets++;
        if (fabs(a - rint(a)) < 1e-9) ties++;
        if (a != b) val++;
        if ((int)a != (int)b) bad++;
    };
    if (family == 0) {
        for (int n = -1080; n <= 1079; ++n)
            for (int k = 1; k <= 64; ++k) {
                for (int q = 0; q < k; ++q) h[q] = n;
                check(k);
            }
    } else if (family == 1) {
        for (int n = -360; n <= 719; ++n)
            for (int a = 1; a <= 89; a += 4)
                for (int k = 0; k <= 3; ++k) {
                    h[0] = n - a;
                    h[1] = n + a;
                    for (int q = 0; q < k; ++q) h[2 + q] = n;
                    check(2 + k);
                }
    } else {
        for (int it = 0; it < 2000000; ++it) {
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            int cnt = 1 + (int)((lcg >> 33) % 6);
            for (int q = 0; q < cnt; ++q) {
                lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
                h[q] = (int)((lcg >> 33) % 1440) - 720;
            }
            check(cnt);
        }
    }
    *n_sets = sets; *n_ties = ties; *n_val = val;
    return bad;
}
int hd_inside(double x, double y) { return t2d::inside_square<double>(x, y) ? 1 : 0; }
}
