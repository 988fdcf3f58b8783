// step_f64.cu — the fp64 parity path.  Compiled with --fmad=false: the reference is built for baseline
// x86-64 (no FMA), so every product and sum must round separately for bit-identical results.
#include "kernels.cuh"
#include "step_fast2.cuh"
namespace t2d {
template struct Launch<double>;
}
#define T2D_IO_IMPL
#include "io_kernels.cuh"
namespace t2d {
template struct IoLaunch<double>;
}
