// io_kernels.cuh — conversion between the reference's host layouts (Eigen column-major doubles, int32) and
// the device SoA state.  Raw host arrays are copied to a device staging area and converted there, so the
// PCIe traffic of t2d_step_host is exactly the reference's own array sizes.
#pragma once
#include "t2d_internal.h"

namespace t2d {

struct HostViewIn {        // device copies of the caller's arrays (any may be null)
    const double* uv;      // [2N]
    const int* heading;    // [N]
    const int* vid;        // [N]
    const double* r3d;     // [3N]
    const uint32_t* ids;   // [N]
};
struct HostViewOut {
    double* uv;
    int* heading;
    int* vid;
    double* r3d;
    double* rdot;
    int* color;
    int* face;
    double* F;
    int* new_heading;
    uint32_t* ids;
};

template <typename R>
__global__ void __launch_bounds__(256) k_ingest(int N, HostViewIn in, ParticleArrays<R> p)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    Real2<R> u = {(R)in.uv[i], (R)in.uv[N + i]};
    p.uv[i] = u;
    Pos3<R> X = {R(0), R(0), R(0), (R)(in.heading ? in.heading[i] : 0)};
    if (in.r3d) {
        X.x = (R)in.r3d[i];
        X.y = (R)in.r3d[N + i];
        X.z = (R)in.r3d[2 * N + i];
    }
    p.pos[i] = X;
    p.aux[i] = make_int4(in.vid ? in.vid[i] : 0, -1, (int)(in.ids ? in.ids[i] : (uint32_t)i), i);
    Real2<R> z = {R(0), R(0)};
    p.rdot[i] = z;
    p.color[i] = 0;
}

// Device-side seeding (replaces CellHelper::init_particle_position, CellHelper.cpp:43-67, whose
// std::random_device / mt19937 stream is neither reproducible nor parallel): particle i draws from the Philox stream
// (seed, draw k, id).  mode 0 = the synthetic inputs of §8d: u, v ~ U(0, 1), heading ~ U{0..359}; mode 1 = the reference's
// scheme: the gravity centre of a random face (drawn with replacement — the reference erases drawn faces from a list it
// searches linearly, which only works while N << F) and a random heading.
template <typename R>
__global__ void __launch_bounds__(256) k_seed(int N, uint64_t seed, int mode, uint32_t first_id, int F, const TriUV<R>* __restrict__ tri,
                                              ParticleArrays<R> p)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint32_t id = first_id + (uint32_t)i;
    Real2<R> u;
    if (mode == 1) {
        int f = (int)(philox_uniform(seed, 0, id) * (double)F);
        f = f < 0 ? 0 : (f > F - 1 ? F - 1 : f);
        const TriUV<R> t = tri[f];
        u.x = (t.ax + t.bx + t.cx) / R(3);   // get_face_gravity_center_coord
        u.y = (t.ay + t.by + t.cy) / R(3);
    } else {
        u.x = (R)philox_uniform(seed, 0, id);
        u.y = (R)philox_uniform(seed, 1, id);
    }
    int h = (int)(philox_uniform(seed, 2, id) * 360.0);
    h = h > 359 ? 359 : h;
    p.uv[i] = u;
    Pos3<R> X = {R(0), R(0), R(0), (R)h};
    p.pos[i] = X;
    p.aux[i] = make_int4(0, -1, (int)id, i);
    Real2<R> z = {R(0), R(0)};
    p.rdot[i] = z;
    p.color[i] = 0;
}

// slot s holds the particle that sits at index aux[s].w of the caller's arrays.  Slab mode (offsets != null): halo
// copies are skipped and owned particles are written in slot order at their compaction offset.
template <typename R>
__global__ void __launch_bounds__(256) k_egest(int resident, int N, const int* __restrict__ offsets, ParticleArrays<R> p,
                                               const Real2<R>* F, const int* new_heading, HostViewOut out)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= resident) return;
    const int4 ax = p.aux[s];
    int o = ax.w;
    if (offsets) {
        if (ax.w < 0) return;
        o = offsets[s];
    }
    if (out.uv) {
        Real2<R> u = p.uv[s];
        out.uv[o] = (double)u.x;
        out.uv[N + o] = (double)u.y;
    }
    if (out.vid) out.vid[o] = ax.x;
    if (out.face) out.face[o] = ax.y;
    if (out.ids) out.ids[o] = (uint32_t)ax.z;
    if (out.r3d || out.heading) {
        Pos3<R> X = p.pos[s];
        if (out.heading) out.heading[o] = (int)X.w;
        if (out.r3d) {
            out.r3d[o] = (double)X.x;
            out.r3d[N + o] = (double)X.y;
            out.r3d[2 * N + o] = (double)X.z;
        }
    }
    if (out.rdot) {
        Real2<R> r = p.rdot[s];
        out.rdot[o] = (double)r.x;
        out.rdot[N + o] = (double)r.y;
    }
    if (out.color) out.color[o] = p.color[s];
    if (out.F) {
        Real2<R> f = F[s];
        out.F[o] = (double)f.x;
        out.F[N + o] = (double)f.y;
    }
    if (out.new_heading) out.new_heading[o] = new_heading[s];
}

// fp32 staging for the host-buffer path of fp32 contexts (Engine::step_host32): the PCIe transfers carry floats, the widening
// to the caller's doubles happens on host threads while the next chunk is in flight (float -> double is exact, so the caller
// sees the same values as through k_egest)
struct HostViewIn32 {
    const float* uv;       // [2N]
    const int* heading;    // [N]
    const int* vid;        // [N] or null
    const float* r3d;      // [3N] or null
    const int* face_hint;  // [N] or null: the face every particle had when the previous call returned
};
struct HostViewOut32 {
    float* uv;             // [2N]
    int* heading;
    int* vid;
    float* r3d;            // [3N]
    float* rdot;           // [2N]
    int* color;
    int* face_hint;        // [N] device-resident, for the next call's re-projection
};
static __global__ void __launch_bounds__(256) k_ingest32(int N, HostViewIn32 in, ParticleArrays<float> p)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    Real2<float> u = {in.uv[i], in.uv[N + i]};
    p.uv[i] = u;
    Pos3<float> X = {0.0f, 0.0f, 0.0f, (float)in.heading[i]};
    if (in.r3d) {
        X.x = in.r3d[i];
        X.y = in.r3d[N + i];
        X.z = in.r3d[2 * N + i];
    }
    p.pos[i] = X;
    p.aux[i] = make_int4(in.vid ? in.vid[i] : 0, in.face_hint ? in.face_hint[i] : -1, i, i);
    Real2<float> z = {0.0f, 0.0f};
    p.rdot[i] = z;
    p.color[i] = 0;
}
static __global__ void __launch_bounds__(256) k_egest32(int N, ParticleArrays<float> p, HostViewOut32 out)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const int4 ax = p.aux[s];
    const int o = ax.w;
    const Real2<float> u = p.uv[s];
    out.uv[o] = u.x;
    out.uv[N + o] = u.y;
    out.vid[o] = ax.x;
    const Pos3<float> X = p.pos[s];
    out.heading[o] = (int)X.w;
    out.r3d[o] = X.x;
    out.r3d[N + o] = X.y;
    out.r3d[2 * N + o] = X.z;
    const Real2<float> r = p.rdot[s];
    out.rdot[o] = r.x;
    out.rdot[N + o] = r.y;
    out.color[o] = p.color[s];
    if (out.face_hint) out.face_hint[o] = ax.y;
}

// The same export straight from the LEAN state, one thread per CALLER index: the writes into the ten output arrays are
// coalesced and what is gathered are whole 32-byte records (k_egest32 scatters 4-byte pieces by caller index: 245 us at
// 2 M particles against 60 us here).  cur = sorted records + aux, pre = the pre-sort side holding r_dot / colour.
static __global__ void __launch_bounds__(256) k_egest32_lean(int N, const int* __restrict__ inv, const int* __restrict__ src,
                                                             ParticleArrays<float> cur, ParticleArrays<float> pre, HostViewOut32 out)
{
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= N) return;
    const int s = inv[o];
    const float4 r0 = cur.rec[2 * (size_t)s], r1 = cur.rec[2 * (size_t)s + 1];
    const int4 ax = cur.aux[s];
    const int i = src[s];
    const Real2<float> rd = pre.rdot[i];
    out.uv[o] = r1.x;
    out.uv[N + o] = r1.y;
    out.vid[o] = ax.x;
    out.heading[o] = __float_as_int(r1.w);
    out.r3d[o] = r0.x;
    out.r3d[N + o] = r0.y;
    out.r3d[2 * N + o] = r0.z;
    out.rdot[o] = rd.x;
    out.rdot[N + o] = rd.y;
    out.color[o] = pre.color[i];
    if (out.face_hint) out.face_hint[o] = ax.y;
}

template <typename R> __global__ void __launch_bounds__(256) k_owned_flags(int n, ParticleArrays<R> p, int* flags)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) flags[s] = p.aux[s].w >= 0 ? 1 : 0;
}

template <typename R> __global__ void __launch_bounds__(256) k_in2(int N, const double* src, Real2<R>* dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    Real2<R> u = {(R)src[i], (R)src[N + i]};
    dst[i] = u;
}
template <typename R> __global__ void __launch_bounds__(256) k_out2(int N, const Real2<R>* src, double* dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    Real2<R> u = src[i];
    dst[i] = (double)u.x;
    dst[N + i] = (double)u.y;
}
template <typename R> __global__ void __launch_bounds__(256) k_outN(int N, int cols, const R* src, double* dst)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * cols) return;
    dst[i] = (double)src[i];
}

template <typename R> struct IoLaunch {
    static void ingest(int N, const HostViewIn& in, const ParticleArrays<R>& p, cudaStream_t s);
    static void egest(int resident, int N, const int* offsets, const ParticleArrays<R>& p, const Real2<R>* F,
                      const int* new_heading, const HostViewOut& out, cudaStream_t s);
    static void seed(int N, uint64_t seed, int mode, uint32_t first_id, int F, const TriUV<R>* tri, const ParticleArrays<R>& p,
                     cudaStream_t s);
    static void owned_flags(int n, const ParticleArrays<R>& p, int* flags, cudaStream_t s);
    static void ingest32(int N, const HostViewIn32& in, const ParticleArrays<R>& p, cudaStream_t s);
    static void egest32(int N, const ParticleArrays<R>& p, const HostViewOut32& out, cudaStream_t s);
    static void egest32_lean(int N, const int* inv, const int* src, const ParticleArrays<R>& cur, const ParticleArrays<R>& pre,
                             const HostViewOut32& out, cudaStream_t s);
    static void in2(int N, const double* src, Real2<R>* dst, cudaStream_t s);
    static void out2(int N, const Real2<R>* src, double* dst, cudaStream_t s);
    static void outN(int N, int cols, const R* src, double* dst, cudaStream_t s);
};

#ifdef T2D_IO_IMPL
template <typename R>
void IoLaunch<R>::ingest(int N, const HostViewIn& in, const ParticleArrays<R>& p, cudaStream_t s)
{
    if (N > 0) k_ingest<R><<<(N + 255) / 256, 256, 0, s>>>(N, in, p);
}
template <typename R>
void IoLaunch<R>::egest(int resident, int N, const int* offsets, const ParticleArrays<R>& p, const Real2<R>* F,
                        const int* new_heading, const HostViewOut& out, cudaStream_t s)
{
    if (resident > 0) k_egest<R><<<(resident + 255) / 256, 256, 0, s>>>(resident, N, offsets, p, F, new_heading, out);
}
template <typename R>
void IoLaunch<R>::seed(int N, uint64_t seed, int mode, uint32_t first_id, int F, const TriUV<R>* tri, const ParticleArrays<R>& p,
                       cudaStream_t s)
{
    if (N > 0) k_seed<R><<<(N + 255) / 256, 256, 0, s>>>(N, seed, mode, first_id, F, tri, p);
}
template <typename R> void IoLaunch<R>::ingest32(int N, const HostViewIn32& in, const ParticleArrays<R>& p, cudaStream_t s)
{
    if constexpr (sizeof(R) == 4) {
        if (N > 0) k_ingest32<<<(N + 255) / 256, 256, 0, s>>>(N, in, p);
    }
}
template <typename R> void IoLaunch<R>::egest32(int N, const ParticleArrays<R>& p, const HostViewOut32& out, cudaStream_t s)
{
    if constexpr (sizeof(R) == 4) {
        if (N > 0) k_egest32<<<(N + 255) / 256, 256, 0, s>>>(N, p, out);
    }
}
template <typename R>
void IoLaunch<R>::egest32_lean(int N, const int* inv, const int* src, const ParticleArrays<R>& cur, const ParticleArrays<R>& pre,
                               const HostViewOut32& out, cudaStream_t s)
{
    if constexpr (sizeof(R) == 4) {
        if (N > 0) k_egest32_lean<<<(N + 255) / 256, 256, 0, s>>>(N, inv, src, cur, pre, out);
    }
}
template <typename R> void IoLaunch<R>::owned_flags(int n, const ParticleArrays<R>& p, int* flags, cudaStream_t s)
{
    if (n > 0) k_owned_flags<R><<<(n + 255) / 256, 256, 0, s>>>(n, p, flags);
}
template <typename R> void IoLaunch<R>::in2(int N, const double* src, Real2<R>* dst, cudaStream_t s)
{
    if (N > 0) k_in2<R><<<(N + 255) / 256, 256, 0, s>>>(N, src, dst);
}
template <typename R> void IoLaunch<R>::out2(int N, const Real2<R>* src, double* dst, cudaStream_t s)
{
    if (N > 0) k_out2<R><<<(N + 255) / 256, 256, 0, s>>>(N, src, dst);
}
template <typename R> void IoLaunch<R>::outN(int N, int cols, const R* src, double* dst, cudaStream_t s)
{
    if (N > 0) k_outN<R><<<(N * cols + 255) / 256, 256, 0, s>>>(N, cols, src, dst);
}
#endif

}  // namespace t2d
