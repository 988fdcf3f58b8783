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
