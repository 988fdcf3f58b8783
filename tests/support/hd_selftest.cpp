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
int hd_inside(double x, double y) { return t2d::inside_square<double>(x, y) ? 1 : 0; }
}
