// step_f32.cu — the fp32 fast path (FMA contraction allowed).
#include "kernels.cuh"
#include "step_fast2.cuh"
namespace t2d {
template struct Launch<float>;
}
#define T2D_IO_IMPL
#include "io_kernels.cuh"
namespace t2d {
template struct IoLaunch<float>;
}
#ifdef T2D_F2_TIMELINE
extern "C" __attribute__((visibility("default"))) int t2d_dev_timeline(unsigned long long* out) {
    return (int)cudaMemcpyFromSymbol(out, t2d::g_f2_timeline, sizeof(unsigned long long) * 2 * 8192);
}
extern "C" __attribute__((visibility("default"))) int t2d_dev_rowstat(unsigned* out) {
    return (int)cudaMemcpyFromSymbol(out, t2d::g_f2_rowstat, sizeof(unsigned) * 4 * 131072);
}
#endif
