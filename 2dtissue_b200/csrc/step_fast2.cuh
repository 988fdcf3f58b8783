// step_fast2.cuh — k_step_fast2: stages 2-5 of the step for the Euclidean criterion, fp32 fast path (the benchmarked kernel).
//
// Replaces (reference, /root/reference/src/simulation): Locomotion::get_dist_vect + get_distances_between_particles
// (Locomotion.cpp:94-162), ForceHelper::calculate_forces_between_particles (Locomotion/ForceHelper.cpp:34-104),
// OrientationHelper::calculate_average_n_within_distance (Locomotion/OrientationHelper.cpp:29-116), the Euler step
// (Locomotion.cpp:71-84), EuclideanTiling (Locomotion/EuclideanTiling.cpp:31-208), CellHelper::get_r3d (CellHelper.cpp:71-229),
// Validation (Validation.cpp:40-72) and count_particle_neighbors (2DTissue.cpp:254-268).
//
// Same arithmetic as the round-1 kernel (k_step_euclid_fast, kept behind T2D_STEP=legacy); what changed is where the
// instructions and the waiting went (ncu, profiles/r01_step_fast_lines.md: 17 % of the kernel's instructions found the nine
// candidate ranges, 18 % of its stall samples waited for the cs[j] gather, issue slots 53 % busy):
//   * ONE 32-byte record per candidate {x, y, z, trig slot | u, v, cell, heading} (made by the counting-sort scatter): the distance
//     test reads the first half, an in-range pair the second half of the SAME sector — no second and third gather;
//   * (cos n, sin n) of a neighbour's heading comes from a 361-entry copy of the host-libm table in shared memory, loaded
//     once per CTA by the TMA unit (cp.async.bulk + mbarrier); the 16-byte-per-particle cs array no longer exists;
//   * the nine candidate ranges of a particle are two loads each from the STATIC neighbourhood table of its cell
//     (t2d_internal.h NBR_STRIDE) instead of a word lookup, two popcounts and bounds logic per row; the ranges are ordered
//     longest-first by a sorting network on packed 32-bit keys (2 instructions per compare-exchange);
//   * persistent CTAs, warps independent of each other (no CTA barrier after the table load): every warp pulls one 32-slot
//     row per ticket from a device-side queue, and takes the next ticket only when its candidate loops have ended (rows
//     differ ~100x in cost; a ticket taken a row ahead waited behind that row);
//   * launched with programmatic stream serialisation: the prologue up to pdl_wait() overlaps the tail of the scatter.
#pragma once
#include "kernels.cuh"

namespace t2d {

constexpr int F2_THREADS = 128;
#ifndef T2D_F2_MIN_BLOCKS
#define T2D_F2_MIN_BLOCKS 8
#endif
#ifndef T2D_F2_GRAB
#define T2D_F2_GRAB 1   // consecutive 32-slot rows per queue ticket (measured: 1 -> 0.329 ms, 2 -> 0.362 ms, 4 -> 0.409 ms: the tail wins)
#endif
#ifndef T2D_F2_PACKED
#define T2D_F2_PACKED 1   // 1: packed-fp32 candidate block (sm_100 FADD2 / FMUL2 / FFMA2; measured 0.343 -> 0.330 ms); 0: scalar fp32
#endif
#ifndef T2D_F2_UNROLL
#define T2D_F2_UNROLL 4
#endif
constexpr int F2_TRIG_N = 361;   // headings 0..360: what alignment produces (OrientationHelper.cpp:102-116); seam re-entry makes the rest
constexpr unsigned F2_TIE_ULPS = 8;

#ifdef T2D_F2_TIMELINE
static __device__ unsigned long long g_f2_timeline[2 * 8192];   // dev builds: (start, end) ns of every warp of the last launch
static __device__ unsigned g_f2_rowstat[4 * 131072];            // per row: start ns (relative, low 32 bits), duration ns, max lane trips, sum lane trips
#endif

struct alignas(128) F2Smem {
    double2 trig[F2_TRIG_N + 3];            // entries 361 and 362 are overwritten with (0, 0): see f2_candidate
    int rkey[NRANGE][F2_THREADS];           // per thread (column): ranges longest first, (length << 4) | row
    int rbeg[NRANGE][F2_THREADS];           // per thread: first slot of row m's range (not sorted)
    unsigned long long bar;
};

// ---- mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP) --------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* b, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity)
{
    while (!mbar_try_wait(b, parity)) {
    }
}
// 1-D bulk copy global -> shared; dst, src and bytes are multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}

struct F2Consts {
    float r2s, r2c, g1, g0;
    unsigned tie_s_lo, tie_c_lo;   // bits(r^2) - F2_TIE_ULPS - 1
};
struct F2Acc {
    float fx = 0.0f, fy = 0.0f;
    double mx = 0.0, my = 0.0;
    int color = 0, hits = 0;
    unsigned ties = 0;
};

// one 32-byte record with ONE load instruction (LDG.E.256, new on sm_100): {x, y, z, trig slot | u, v, cell, heading}
struct F2Rec { float x, y, z; unsigned slot; float u, v; int cell, heading; };
__device__ __forceinline__ F2Rec f2_load(const float4* __restrict__ q)
{
    F2Rec r;
    asm("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=r"(r.slot), "=f"(r.u), "=f"(r.v), "=r"(r.cell), "=r"(r.heading)
        : "l"(q));
    return r;
}
__device__ __forceinline__ double2 f2_lds_trig(uint32_t addr)
{
    double2 t;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t.x), "=d"(t.y) : "r"(addr));
    return t;
}

// One candidate, branch-free.  d^2 in fp32; colour count (0 != d <= 2.4 sigma, 2DTissue.cpp:254-268); in-range pair:
// F_ij / d = -k / d + k / 2 sigma (ForceHelper.cpp:84-104) as one FMA on 1/d, d = 0 -> 0.001 (ForceHelper.cpp:59-62); the unit
// vector of the neighbour's heading from the shared-memory copy of the host-libm table, summed in double
// (OrientationHelper.cpp:43-58).  About 56 % of the candidates are in range, so some lane of the warp needs the pair term
// almost every time: it is computed for every candidate and masked — an out-of-range candidate reads the table's zero
// entry (slot 361) and gets g = 0.  A heading outside the table has slot 362 (also zero); `oobm` (the largest slot seen) then
// sends the thread through the rare global-table path after the group.  TIES: log candidates within F2_TIE_ULPS ulps of a squared cutoff — a
// neighbour-set or colour difference against the fp64 oracle must be one of these (tests/test_gpu_fastpath.py).
template <bool TIES>
__device__ __forceinline__ void f2_candidate(const F2Rec& J, const float px, const float py, const float pz, const float2 ui,
                                             const F2Consts& k, const uint32_t s_trig, F2Acc& acc, unsigned& oobm)
{
    // hand-scheduled PTX: 25 instructions per candidate, every value defined on every path (nothing for ptxas to spill or
    // to turn into branches); slot 361 = zero entry for "not in range", 362 = zero entry for "heading not in the table"
    float d2;
#if T2D_F2_PACKED
    // packed fp32 (sm_100 FADD2 / FMUL2 / FFMA2): (dx, dy), (ux, uy) and the force accumulation are one instruction each
    asm("{\n\t"
        ".reg .pred p, nz, c;\n\t"
        ".reg .f32 dx2, dy2, dz, inv, g;\n\t"
        ".reg .u32 idx, ad;\n\t"
        ".reg .f64 tc, ts;\n\t"
        ".reg .b64 pxy, jxy, dxy, sq, uxy, juv, gg, fxy;\n\t"
        "mov.b64 pxy, {%8, %9};\n\t"
        "mov.b64 jxy, {%11, %12};\n\t"
        "sub.ftz.f32x2 dxy, pxy, jxy;\n\t"
        "sub.ftz.f32 dz, %10, %13;\n\t"
        "mul.ftz.f32x2 sq, dxy, dxy;\n\t"
        "mov.b64 {dx2, dy2}, sq;\n\t"
        "fma.rn.ftz.f32 %7, dz, dz, dx2;\n\t"
        "add.ftz.f32 %7, %7, dy2;\n\t"
        "setp.lt.ftz.f32 p, %7, %19;\n\t"
        "setp.neu.ftz.f32 nz, %7, 0f00000000;\n\t"
        "setp.le.and.ftz.f32 c, %7, %20, nz;\n\t"
        "@c add.s32 %4, %4, 1;\n\t"
        "@p add.s32 %5, %5, 1;\n\t"
        "selp.u32 idx, %14, 361, p;\n\t"
        "max.u32 %6, %6, idx;\n\t"
        "shl.b32 ad, idx, 4;\n\t"
        "add.u32 ad, ad, %23;\n\t"
        "ld.shared.v2.f64 {tc, ts}, [ad];\n\t"
        "rsqrt.approx.ftz.f32 inv, %7;\n\t"
        "selp.f32 inv, inv, 0f447A0000, nz;\n\t"
        "fma.rn.ftz.f32 g, inv, %21, %22;\n\t"
        "selp.f32 g, g, 0f00000000, p;\n\t"
        "mov.b64 uxy, {%17, %18};\n\t"
        "mov.b64 juv, {%15, %16};\n\t"
        "sub.ftz.f32x2 uxy, uxy, juv;\n\t"
        "mov.b64 gg, {g, g};\n\t"
        "mov.b64 fxy, {%0, %1};\n\t"
        "fma.rn.ftz.f32x2 fxy, gg, uxy, fxy;\n\t"
        "mov.b64 {%0, %1}, fxy;\n\t"
        "add.f64 %2, %2, tc;\n\t"
        "add.f64 %3, %3, ts;\n\t"
        "}"
        : "+f"(acc.fx), "+f"(acc.fy), "+d"(acc.mx), "+d"(acc.my), "+r"(acc.color), "+r"(acc.hits), "+r"(oobm), "=f"(d2)
        : "f"(px), "f"(py), "f"(pz), "f"(J.x), "f"(J.y), "f"(J.z), "r"(J.slot), "f"(J.u), "f"(J.v), "f"(ui.x), "f"(ui.y),
          "f"(k.r2s), "f"(k.r2c), "f"(k.g1), "f"(k.g0), "r"(s_trig));
#else
    asm("{\n\t"
        ".reg .pred p, nz, c;\n\t"
        ".reg .f32 dx, dy, dz, inv, g, ux, uy;\n\t"
        ".reg .u32 idx, ad;\n\t"
        ".reg .f64 tc, ts;\n\t"
        "sub.ftz.f32 dx, %8, %11;\n\t"
        "sub.ftz.f32 dy, %9, %12;\n\t"
        "sub.ftz.f32 dz, %10, %13;\n\t"
        "mul.ftz.f32 %7, dx, dx;\n\t"
        "fma.rn.ftz.f32 %7, dy, dy, %7;\n\t"
        "fma.rn.ftz.f32 %7, dz, dz, %7;\n\t"
        "setp.lt.ftz.f32 p, %7, %19;\n\t"
        "setp.neu.ftz.f32 nz, %7, 0f00000000;\n\t"
        "setp.le.and.ftz.f32 c, %7, %20, nz;\n\t"
        "@c add.s32 %4, %4, 1;\n\t"
        "@p add.s32 %5, %5, 1;\n\t"
        "selp.u32 idx, %14, 361, p;\n\t"
        "max.u32 %6, %6, idx;\n\t"
        "shl.b32 ad, idx, 4;\n\t"
        "add.u32 ad, ad, %23;\n\t"
        "ld.shared.v2.f64 {tc, ts}, [ad];\n\t"
        "rsqrt.approx.ftz.f32 inv, %7;\n\t"
        "selp.f32 inv, inv, 0f447A0000, nz;\n\t"
        "fma.rn.ftz.f32 g, inv, %21, %22;\n\t"
        "selp.f32 g, g, 0f00000000, p;\n\t"
        "sub.ftz.f32 ux, %17, %15;\n\t"
        "sub.ftz.f32 uy, %18, %16;\n\t"
        "fma.rn.ftz.f32 %0, g, ux, %0;\n\t"
        "fma.rn.ftz.f32 %1, g, uy, %1;\n\t"
        "add.f64 %2, %2, tc;\n\t"
        "add.f64 %3, %3, ts;\n\t"
        "}"
        : "+f"(acc.fx), "+f"(acc.fy), "+d"(acc.mx), "+d"(acc.my), "+r"(acc.color), "+r"(acc.hits), "+r"(oobm), "=f"(d2)
        : "f"(px), "f"(py), "f"(pz), "f"(J.x), "f"(J.y), "f"(J.z), "r"(J.slot), "f"(J.u), "f"(J.v), "f"(ui.x), "f"(ui.y),
          "f"(k.r2s), "f"(k.r2c), "f"(k.g1), "f"(k.g0), "r"(s_trig));
#endif
    if (TIES) {
        const unsigned bm1 = __float_as_uint(d2) - 1u;
        if ((bm1 - k.tie_s_lo) <= 2u * F2_TIE_ULPS || (bm1 - k.tie_c_lo) <= 2u * F2_TIE_ULPS) acc.ties++;
    }
}
// the rare path: an in-range neighbour whose heading is outside 0..360 (it crossed the seam in the last step)
template <int = 0>
__device__ __noinline__ double2 f2_oob_term(const double2* g_trig, float r2s, float px, float py, float pz, float jx, float jy, float jz,
                                            unsigned slot, int heading)
{
    const float dx = px - jx, dy = py - jy, dz = pz - jz;
    const float d2 = dx * dx + dy * dy + dz * dz;
    if (!(d2 < r2s) || slot != (unsigned)(F2_TRIG_N + 1)) return make_double2(0.0, 0.0);
    unsigned long long fb = 0;
    return trig_lookup(g_trig, heading, fb);
}

// One trip of the candidate loop: T2D_F2_UNROLL records loaded together, then their candidate blocks (which ptxas interleaves).
// MASKED: only the first `cnt` records exist; the others are read from the sentinel record (far away, zero trig slot).
template <bool TIES, bool MASKED>
__device__ __forceinline__ void f2_trip(const float4* __restrict__ q, int cnt, const float4* __restrict__ sent,
                                        const double2* __restrict__ g_trig, const float px, const float py, const float pz,
                                        const float2 ui, const F2Consts& k, const uint32_t s_trig, F2Acc& acc)
{
    F2Rec J[T2D_F2_UNROLL];
#pragma unroll
    for (int u = 0; u < T2D_F2_UNROLL; ++u) J[u] = f2_load((!MASKED || u < cnt) ? q + 2 * u : sent);
    unsigned oob = 0;
#pragma unroll
    for (int u = 0; u < T2D_F2_UNROLL; ++u) f2_candidate<TIES>(J[u], px, py, pz, ui, k, s_trig, acc, oob);
    if (__builtin_expect(oob > (unsigned)F2_TRIG_N, 0)) {
#pragma unroll 1
        for (int u = 0; u < (MASKED ? cnt : T2D_F2_UNROLL); ++u) {
            const F2Rec O = f2_load(q + 2 * u);
            const double2 tr = f2_oob_term(g_trig, k.r2s, px, py, pz, O.x, O.y, O.z, O.slot, O.heading);
            acc.mx += tr.x;
            acc.my += tr.y;
        }
    }
}

// a particle outside the static index (overflow bucket; cannot happen for points on the mesh): its ranges from its coordinates,
// straight into the thread's shared-memory columns (arguments by value: a reference to the kernel's parameter block would
// force a copy of it into local memory)
static __device__ __noinline__ void f2_ranges_by_coords(DevVox<float> vx, const int* start, float x, float y, float z, int* rbeg,
                                                        int* rkey)
{
    Pos3<float> P = {x, y, z, 0.0f};
    int c[3];
    cell_coords<float>(vx, P, c);
    for (int m = 0; m < 9; ++m) {
        int lo, hi, b = 0, len = 0;
        row_cells<float>(vx, c[0], c[1] + (m % 3) - 1, c[2] + (m / 3) - 1, lo, hi);
        if (hi > lo) {
            b = start[lo];
            len = start[hi] - b;
        }
        rbeg[m * F2_THREADS] = b;
        rkey[m * F2_THREADS] = (min(len, 0x07ffffff) << 4) | m;
    }
}

template <bool MOVING, bool TIES>
__global__ void __launch_bounds__(F2_THREADS, T2D_F2_MIN_BLOCKS) k_step_fast2(StepArgs<float> a, int* queue, int* queue_next)
{
    typedef float R;
    __shared__ F2Smem sm;
    const int tid = threadIdx.x, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&sm.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {    // cos/sin of 0..360 degrees as the host's libm gives them: one bulk copy per CTA
        constexpr unsigned bytes = (unsigned)(sizeof(double2) * (F2_TRIG_N + 3));
        mbar_arrive_expect_tx(&sm.bar, bytes);
        bulk_g2s(sm.trig, a.trig_d + (0 - TRIG_MIN), bytes, &sm.bar);
    }

    F2Consts k;
    {
        k.r2s = a.two_sigma * a.two_sigma;
        k.r2c = a.color_r * a.color_r;
        k.g1 = -a.k;
        k.g0 = a.k / a.two_sigma;
        k.tie_s_lo = __float_as_uint(k.r2s) - F2_TIE_ULPS - 1u;
        k.tie_c_lo = __float_as_uint(k.r2c) - F2_TIE_ULPS - 1u;
    }
    pdl_wait();   // everything above touches only constants and this CTA's shared memory (programmatic dependent launch)
    if (blockIdx.x == 0 && tid == 0) *queue_next = 0;   // the queue of the NEXT launch (launches alternate between two counters)
    const int nres = resident_count<R>(a);
    const int ngrabs = (nres + 32 * T2D_F2_GRAB - 1) / (32 * T2D_F2_GRAB);
    const int M = a.vox.M;
    const float4* __restrict__ rec = a.cur.rec;
    const float4* __restrict__ sent = a.rec_sentinel;
    const int* __restrict__ start = a.start;
    const int ob = start[M], ol = start[M + 1] - ob;   // overflow bucket: normally empty
    unsigned npairs_w = 0, nties_w = 0, ncut_w = 0, fb_w = 0;

#ifdef T2D_F2_TIMELINE
    unsigned long long t_start;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
#endif
    int next = 0;
    if (lane == 0) next = atomicAdd(queue, 1);
    mbar_wait(&sm.bar, 0);   // the table has landed (every thread observes the barrier itself)
    if (tid < 2) sm.trig[F2_TRIG_N + tid] = make_double2(0.0, 0.0);
    __syncthreads();
    const uint32_t s_trig = smem_u32(sm.trig);
    for (;;) {
        // The ticket was taken when the previous row's candidate loops ended, not a whole row ahead: rows differ a lot in
        // cost (after 200 steps of the headline run a row's candidate trips spread from 3 to 330 around a mean of 50, and
        // its duration follows them), and a ticket taken a row ahead sat behind that row — measured with per-row timers,
        // the last rows of the queue started ~100 us after half the warps had run out of work.  Late binding: -10 %.
        const int grab = __shfl_sync(0xffffffffu, next, 0);
        if (grab >= ngrabs) break;
#if defined(T2D_F2_TIMELINE) && T2D_F2_TIMELINE == 1   // per-row timers (heavier than the per-warp ones)
        int lane_trips = 0;
        unsigned long long t_row;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_row));
#endif
#pragma unroll 1
        for (int sub = 0; sub < T2D_F2_GRAB; ++sub) {
            const int i = (grab * T2D_F2_GRAB + sub) * 32 + lane;
            const bool resident = i < nres;
            const int4 ai = resident ? a.cur.aux[i] : make_int4(0, 0, 0, ORIGIN_DEAD);
            const bool live = resident && ai.w >= 0;   // slab mode: halo copies are read by others, never advanced
            if (MOVING && resident && !live) {
                a.alt.aux[i] = make_int4(0, -1, ai.z, ORIGIN_DEAD);
                a.key[i] = KEY_DROP;
            }
            F2Acc acc;
            float2 ui = make_float2(0.0f, 0.0f);
            int nh = 0, own_cell = -1;
            float ox = 0.0f, oy = 0.0f, oz = 0.0f;
#ifdef T2D_F2_ABLATE
            for (int rep = 0; rep < (a.ablate == 8 ? 2 : 1); ++rep) {   // dev: second pass = the same work with warm caches
            if (rep) acc = F2Acc();
#endif
            if (live) {
                const F2Rec self = f2_load(rec + 2 * (size_t)i);
                const float px = self.x, py = self.y, pz = self.z;
                nh = self.heading;
                ui = make_float2(self.u, self.v);
                const int cell = self.cell;
                own_cell = cell;
                ox = px;
                oy = py;
                oz = pz;
                int key[9];
#ifdef T2D_F2_ABLATE
                if (a.ablate == 5) {
#pragma unroll
                    for (int m = 0; m < 9; ++m) key[m] = m;
                } else
#endif
                if (cell < M) {
                    const int4* nb = reinterpret_cast<const int4*>(a.nbr + (size_t)cell * NBR_STRIDE);
                    int4 q[5];
#pragma unroll
                    for (int h = 0; h < 5; ++h) q[h] = __ldg(nb + h);
#pragma unroll
                    for (int m = 0; m < 9; ++m) {
                        const int lo = (m & 1) ? q[m >> 1].z : q[m >> 1].x, hi = (m & 1) ? q[m >> 1].w : q[m >> 1].y;
                        int b = 0, len = 0;
                        if (hi > lo) {
                            b = start[lo];
                            len = start[hi] - b;
                        }
                        sm.rbeg[m][tid] = b;
                        key[m] = (min(len, 0x07ffffff) << 4) | m;
                    }
                } else {
                    f2_ranges_by_coords(a.vox, start, px, py, pz, &sm.rbeg[0][tid], &sm.rkey[0][tid]);
#pragma unroll
                    for (int m = 0; m < 9; ++m) key[m] = sm.rkey[m][tid];
                }
                // 9-input sorting network (25 compare-exchanges), longest range first: the lanes of a warp then finish their
                // m-th range at about the same time (measured in round 1: 67 instead of 92 warp iterations per particle row)
#define T2D_CSWAP(x, y)                          \
    {                                            \
        const int hi_ = max(key[x], key[y]);     \
        key[y] = min(key[x], key[y]);            \
        key[x] = hi_;                            \
    }
                T2D_CSWAP(0, 1) T2D_CSWAP(3, 4) T2D_CSWAP(6, 7) T2D_CSWAP(1, 2) T2D_CSWAP(4, 5) T2D_CSWAP(7, 8)
                T2D_CSWAP(0, 1) T2D_CSWAP(3, 4) T2D_CSWAP(6, 7) T2D_CSWAP(0, 3) T2D_CSWAP(3, 6) T2D_CSWAP(0, 3)
                T2D_CSWAP(1, 4) T2D_CSWAP(4, 7) T2D_CSWAP(1, 4) T2D_CSWAP(2, 5) T2D_CSWAP(5, 8) T2D_CSWAP(2, 5)
                T2D_CSWAP(1, 3) T2D_CSWAP(5, 7) T2D_CSWAP(2, 6) T2D_CSWAP(4, 6) T2D_CSWAP(2, 4) T2D_CSWAP(2, 3)
                T2D_CSWAP(5, 6)
#undef T2D_CSWAP
                int nr = 0;
#pragma unroll
                for (int m = 0; m < 9; ++m) {
                    sm.rkey[m][tid] = key[m];
                    nr += key[m] >= 16 ? 1 : 0;
#if defined(T2D_F2_TIMELINE) && T2D_F2_TIMELINE == 1   // per-row timers (heavier than the per-warp ones)
                    lane_trips += ((key[m] >> 4) + T2D_F2_UNROLL - 1) / T2D_F2_UNROLL;
#endif
                }
#ifdef T2D_F2_ABLATE
                if (a.ablate == 1 || a.ablate == 4) nr = 0;   // dev: no candidate loop
#endif
                if (ol > 0) {   // the overflow bucket is everybody's candidate
                    sm.rbeg[9][tid] = ob;
                    sm.rkey[nr][tid] = (min(ol, 0x07ffffff) << 4) | 9;
                    nr++;
                }
#pragma unroll 1
                for (int m = 0; m < nr; ++m) {
#ifdef T2D_F2_ABLATE
                    if (a.ablate == 3) {   // dev: loads only
                        const int kk2 = sm.rkey[m][tid];
                        const float4* q2 = rec + 2 * (size_t)sm.rbeg[kk2 & 15][tid];
                        for (int t2 = 0; t2 < (kk2 >> 4); t2 += 4) {
                            const F2Rec a0 = f2_load(q2 + 2 * t2), a1 = f2_load(q2 + 2 * min(t2 + 1, (kk2 >> 4) - 1)),
                                        a2 = f2_load(q2 + 2 * min(t2 + 2, (kk2 >> 4) - 1)), a3 = f2_load(q2 + 2 * min(t2 + 3, (kk2 >> 4) - 1));
                            acc.fx += a0.x + a1.y + a2.z + a3.u;
                        }
                        continue;
                    }
#endif
                    const int kk = sm.rkey[m][tid];
                    const int len = kk >> 4, jb = sm.rbeg[kk & 15][tid];
                    const float4* q = rec + 2 * (size_t)jb;
#ifdef T2D_F2_ABLATE
                    if (a.ablate == 6) q = rec + 2 * (size_t)max(0, min(i - lane, nres - 4 - len));   // dev: every candidate load hits L1
#endif
                    int t = 0;
                    for (; t + T2D_F2_UNROLL <= len; t += T2D_F2_UNROLL, q += 2 * T2D_F2_UNROLL) {
                        f2_trip<TIES, false>(q, T2D_F2_UNROLL, sent, a.trig_d, px, py, pz, ui, k, s_trig, acc);
                    }
                    // the rest of the range as ONE masked trip (the candidates beyond the end read the sentinel record): a
                    // one-candidate-per-trip remainder loop had 38 % of the kernel's load waits for 13 % of its candidates
                    if (t < len) f2_trip<TIES, true>(q, len - t, sent, a.trig_d, px, py, pz, ui, k, s_trig, acc);
                }
            }
#ifdef T2D_F2_ABLATE
            }
#endif
            __syncwarp();   // reconverge: lanes leave the candidate loops at different times, the tail is the same for all
            if (sub == T2D_F2_GRAB - 1 && lane == 0) next = atomicAdd(queue, 1);   // next ticket: in flight during the epilogue
            unsigned npairs = 0, nties = 0;
            if (live) {
                double2 own;
                if ((unsigned)nh < (unsigned)F2_TRIG_N) {
                    own = sm.trig[nh];
                } else {
                    unsigned long long fb = 0;
                    own = trig_lookup(a.trig_d, nh, fb);
                    fb_w += (unsigned)fb;
                }
                PairAcc pa;
                pa.fx = acc.fx;
                pa.fy = acc.fy;
                pa.mx = acc.mx;
                pa.my = acc.my;
                Real2<R> uir = {ui.x, ui.y};
#ifdef T2D_F2_ABLATE
                if (a.ablate == 2 || a.ablate == 4) {   // dev: no epilogue — the state is copied through unchanged (valid input for the sort)
                    const Pos3<R> X = {ox, oy, oz, (R)nh};
                    a.alt.pos[i] = X;
                    a.alt.uv[i] = uir;
                    a.alt.aux[i] = ai;
                    Real2<R> rd = {pa.fx + (float)own.x, pa.fy + (float)(pa.mx + pa.my)};
                    a.alt.rdot[i] = rd;
                    a.alt.color[i] = acc.color + acc.hits;
                    a.key[i] = (uint32_t)own_cell;
                    a.rank[i] = (uint32_t)atomicAdd(&a.count[own_cell], 1);
                    continue;
                }
#endif
                fast_epilogue<MOVING>(a, i, ai, uir, own, pa, acc.color, acc.hits, npairs, nties, own_cell, ox, oy, oz);
            }
            npairs_w += npairs;
            nties_w += nties;
            ncut_w += acc.ties;
        }
#if defined(T2D_F2_TIMELINE) && T2D_F2_TIMELINE == 1   // per-row timers (heavier than the per-warp ones)
        {
            __syncwarp();
            const int mx = __reduce_max_sync(0xffffffffu, lane_trips), sm_ = __reduce_add_sync(0xffffffffu, lane_trips);
            unsigned long long t_row_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_row_end));
            if (lane == 0 && grab < 131072) {
                g_f2_rowstat[4 * grab] = (unsigned)t_row;
                g_f2_rowstat[4 * grab + 1] = (unsigned)(t_row_end - t_row);
                g_f2_rowstat[4 * grab + 2] = (unsigned)mx;
                g_f2_rowstat[4 * grab + 3] = (unsigned)sm_;
            }
        }
#endif
    }
#ifdef T2D_F2_TIMELINE
    if (lane == 0) {   // dev: per-warp timeline (start, end in ns)
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        const int gw = blockIdx.x * (F2_THREADS / 32) + (tid >> 5);
        if (gw < 8192) { g_f2_timeline[2 * gw] = t_start; g_f2_timeline[2 * gw + 1] = t_end; }
    }
#endif
    // diagnostic counters: one atomic per warp and counter
    npairs_w = __reduce_add_sync(0xffffffffu, npairs_w);
    nties_w = __reduce_add_sync(0xffffffffu, nties_w);
    ncut_w = __reduce_add_sync(0xffffffffu, ncut_w);
    fb_w = __reduce_add_sync(0xffffffffu, fb_w);
    if (lane == 0) {
        if (npairs_w) atomicAdd(&a.counters->pairs_in_range, (unsigned long long)npairs_w);
        if (nties_w) atomicAdd(&a.counters->ties_trunc, (unsigned long long)nties_w);
        if (ncut_w) atomicAdd(&a.counters->ties_cutoff, (unsigned long long)ncut_w);
        if (fb_w) atomicAdd(&a.counters->trig_fallbacks, (unsigned long long)fb_w);
    }
}

template <typename R> bool Launch<R>::step_fast2(const StepArgs<R>& a, bool moving, int sm_count, cudaStream_t s)
{
    if constexpr (sizeof(R) == 4) {
        if (!a.cur.rec || !a.nbr || !a.rec_sentinel) return false;
        const int n = a.comm.on ? a.comm.capacity : a.N;
        if (n <= 0) return true;
        const int ngrabs = div_up(n, 32 * T2D_F2_GRAB);
        int grid = sm_count * T2D_F2_MIN_BLOCKS;
        if (grid > div_up(ngrabs, F2_THREADS / 32)) grid = div_up(ngrabs, F2_THREADS / 32);
        int* q0 = a.work_counter + 1 + (a.queue_flip & 1);   // two queue counters, used alternately: every launch zeroes the
        int* q1 = a.work_counter + 1 + ((a.queue_flip + 1) & 1);   // other one (the caller advances queue_flip after each launch)
        if (moving) {
            if (a.count_ties)
                launch_pdl(a.pdl != 0, k_step_fast2<true, true>, grid, F2_THREADS, s, a, q0, q1);
            else
                launch_pdl(a.pdl != 0, k_step_fast2<true, false>, grid, F2_THREADS, s, a, q0, q1);
        } else {
            launch_pdl(a.pdl != 0, k_step_fast2<false, true>, grid, F2_THREADS, s, a, q0, q1);
        }
        return true;
    } else {
        return false;
    }
}

template <typename R> void Launch<R>::build_nbr(const DevVox<R>& vx, int2* nbr, cudaStream_t s)
{
    const size_t nwords = (size_t)vx.ncz * vx.ncy * vx.nwx;
    if (nwords) k_build_nbr<R><<<(unsigned)((nwords + 255) / 256), 256, 0, s>>>(vx, nbr);
}

}  // namespace t2d
