// step_tiled.cuh — k_step_euclid_tiled: stages 2-5 of the step for the Euclidean criterion, fp32 fast path, with the
// candidate data of a whole tile staged in shared memory by the TMA unit.
//
// Replaces (reference, /root/reference/src/simulation): Locomotion::get_dist_vect + get_distances_between_particles
// (Locomotion.cpp:94-162), ForceHelper::calculate_forces_between_particles (Locomotion/ForceHelper.cpp:34-104),
// OrientationHelper::calculate_average_n_within_distance (Locomotion/OrientationHelper.cpp:29-116), the Euler step
// (Locomotion.cpp:71-84), EuclideanTiling (Locomotion/EuclideanTiling.cpp:31-208), CellHelper::get_r3d (CellHelper.cpp:71-229),
// Validation (Validation.cpp:40-72) and count_particle_neighbors (2DTissue.cpp:254-268) — same arithmetic as
// k_step_euclid_fast (kernels.cuh), different data movement:
//
//   * the sorted state is cut into static tiles of consecutive surface cells (DevTiles); persistent CTAs pull tiles from
//     a queue.  For a tile, warp 0 turns the tile's pre-merged cell intervals into slot ranges (two start[] loads and a
//     warp scan) and issues one cp.async.bulk per interval and array (pos: 16 B, uv: 8 B per particle) onto an mbarrier;
//     the copies land while all threads compute their 3 x 3 row ranges.  ncu on k_step_euclid_fast had shown the candidate
//     loop waiting on L1/L2 (long scoreboard 8 warps per issue, issue slots 53 % busy): here every candidate read is an
//     LDS from a tile that is already resident;
//   * (cos n, sin n) of a neighbour's heading comes from a 360-entry copy of the host-libm table in shared memory, so the
//     per-particle `cs` array (16 B written by the sort, 16 B gathered per in-range pair) no longer exists in HBM;
//   * intervals that do not fit the staging buffer (dense clumps) and tiles with more than TILE_IMAX intervals fall back
//     to the same loop over global memory, range by range — results do not depend on what was staged.
#pragma once
#include "kernels.cuh"

#ifndef T2D_TILE_GENERIC
#define T2D_TILE_GENERIC 1
#endif

namespace t2d {

constexpr int TILE_WARPS = 4;              // warps per CTA; every warp is an independent worker with its own tile
constexpr int TILE_THREADS = TILE_WARPS * 32;
#ifndef T2D_TILE_CAP
#define T2D_TILE_CAP 192
#endif
constexpr int TILE_CAP = T2D_TILE_CAP;   // staged particles per tile (pos 16 B + uv 8 B each)
constexpr int TILE_IMAX = 32;            // merged intervals per tile: one lane each
constexpr int TRIG_SMEM_N = 360;         // headings 0..359 (alignment leaves them there; seam re-entry makes the rest)
#ifndef T2D_TILE_MIN_BLOCKS
#define T2D_TILE_MIN_BLOCKS 6
#endif

struct alignas(16) WarpTile {
    float4 pos[TILE_CAP];
    float2 uv[TILE_CAP];
    int rb[NRANGE][32];   // candidate ranges of each lane, longest first: index into pos/uv (>= 0) or ~slot (< 0: global)
    int rl[NRANGE][32];
    int ilo[TILE_IMAX], islot[TILE_IMAX], ioff[TILE_IMAX];   // intervals: first compact cell, first staged slot, offset in pos/uv or -1
    unsigned long long bar;
    unsigned long long pad;
};
struct TiledSmem {
    WarpTile w[TILE_WARPS];
    double2 trig[TRIG_SMEM_N];
    unsigned long long bar_trig;
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

// where does the candidate range that starts at slot b (first compact cell clo) live?  index into the staged arrays, or ~b
__device__ __forceinline__ int tiled_lookup(const WarpTile& sm, int clo, int b)
{
    int q = sm.ilo[16] <= clo ? 16 : 0;   // last interval whose first cell is <= clo (ilo is ascending, padded with INT_MAX)
    q += sm.ilo[q + 8] <= clo ? 8 : 0;
    q += sm.ilo[q + 4] <= clo ? 4 : 0;
    q += sm.ilo[q + 2] <= clo ? 2 : 0;
    q += sm.ilo[q + 1] <= clo ? 1 : 0;
    const int off = sm.ioff[q];
    return off >= 0 ? off + (b - sm.islot[q]) : ~b;
}

// thresholds of the candidate tests, shared by every candidate of a thread
struct PairConsts {
    float r2s, g1, g0;
    unsigned r2c_bits, tie_s_lo, tie_c_lo;   // tie_*_lo = bits(r^2) - TIE_ULPS - 1 (see cand_loop)
};
constexpr unsigned TIE_ULPS = 8;

// One candidate range.  Same arithmetic as k_step_euclid_fast's loop: d^2 in fp32, colour count as one unsigned compare
// on the bit pattern, pair term = pair_term() with (cos, sin) from the shared-memory table.  TIES: count the candidates
// whose d^2 is within TIE_ULPS ulps of (2 sigma)^2 or (color_factor sigma)^2 — the "logged near-cutoff ties" of the
// parity bar: a neighbour-set or colour difference against the fp64 oracle must be one of these.
template <bool TIES>
__device__ __forceinline__ void cand_loop(const float4* __restrict__ cp, const float2* __restrict__ cu, int len, const float4& Pi,
                                          const Real2<float>& ui, const PairConsts& k, const double2* __restrict__ trig_s,
                                          const double2* __restrict__ trig_g, PairAcc& acc, int& color, int& hits, unsigned& ties,
                                          unsigned long long& trig_fb)
{
#pragma unroll FAST_UNROLL
    for (int t = 0; t < len; ++t) {
        const float4 Pj = cp[t];
        const float dx = Pi.x - Pj.x, dy = Pi.y - Pj.y, dz = Pi.z - Pj.z;
        const float d2 = dx * dx + dy * dy + dz * dz;
        const unsigned bm1 = __float_as_uint(d2) - 1u;   // wraps to 0xffffffff for d2 == 0: the particle itself never counts
        // _2DTissue::count_particle_neighbors: 0 != d <= 2.4 sigma
        asm("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(color) : "r"(bm1), "r"(k.r2c_bits));
        if (TIES) {
            const bool near_s = (bm1 - k.tie_s_lo) <= 2u * TIE_ULPS, near_c = (bm1 - k.tie_c_lo) <= 2u * TIE_ULPS;
            if (near_s || near_c) ties++;
        }
        if (d2 < k.r2s) {
            const float2 uj = cu[t];
            const int nj = (int)Pj.w;
            double2 tr;
            if ((unsigned)nj < (unsigned)TRIG_SMEM_N)
                tr = trig_s[nj];
            else
                tr = trig_lookup(trig_g, nj, trig_fb);
            acc.mx += tr.x;
            acc.my += tr.y;
            const float g = fmaf(d2 == 0.0f ? 1000.0f : rsqrtf(d2), k.g1, k.g0);   // pair_term(), kernels.cuh
            acc.fx = fmaf(g, ui.x - uj.x, acc.fx);
            acc.fy = fmaf(g, ui.y - uj.y, acc.fy);
            hits++;
        }
    }
}

template <bool TIES> __global__ void __launch_bounds__(TILE_THREADS, T2D_TILE_MIN_BLOCKS) k_step_euclid_tiled(StepArgs<float> a)
{
    typedef float R;
    __shared__ __align__(128) TiledSmem sm;
    const int tid = threadIdx.x, lane = tid & 31;
    WarpTile& wt = sm.w[tid >> 5];

    if (lane == 0) mbar_init(&wt.bar, 1);
    if (tid == 0) mbar_init(&sm.bar_trig, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();   // the only CTA-wide barrier: from here on the warps never wait for each other
    if (tid == 0) {    // cos/sin of 0..359 degrees as the host's libm gives them: one bulk copy per CTA
        mbar_arrive_expect_tx(&sm.bar_trig, (unsigned)sizeof(sm.trig));
        bulk_g2s(sm.trig, a.trig_d + (0 - TRIG_MIN), (unsigned)sizeof(sm.trig), &sm.bar_trig);
    }
    bool trig_ready = false;
    unsigned parity = 0;

    PairConsts k;
    {
        const float r2c = a.color_r * a.color_r;
        k.r2s = a.two_sigma * a.two_sigma;
        k.g1 = -a.k;
        k.g0 = a.k / a.two_sigma;
        k.r2c_bits = __float_as_uint(r2c);
        k.tie_s_lo = __float_as_uint(k.r2s) - TIE_ULPS - 1u;
        k.tie_c_lo = __float_as_uint(r2c) - TIE_ULPS - 1u;
    }
    const int M = a.vox.M;
    unsigned npairs_w = 0, nties_w = 0, ncut_w = 0;
    unsigned long long trig_fb = 0;

    for (;;) {
        int tile = 0;
        if (lane == 0) tile = atomicAdd(a.tiles.queue, 1);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile > a.tiles.ntiles) break;
        const bool last = tile == a.tiles.ntiles;   // the overflow bucket: particles outside the static index, never staged
        const int c0 = last ? M : tile * a.tiles.tile_cells;
        const int c1 = last ? M + 1 : min(M, c0 + a.tiles.tile_cells);
        const int T0 = a.start[c0], T1 = a.start[c1];
        if (T1 <= T0) continue;   // no particles in this tile

        // ---- staging: intervals -> slot ranges -> offsets in the staging buffer (warp scan) -> bulk copies ----
        bool armed;
        {
            int nq = 0, q0 = 0;
            if (!last) {
                q0 = __ldg(&a.tiles.istart[tile]);
                nq = __ldg(&a.tiles.istart[tile + 1]) - q0;
            }
            int lo = 0x7fffffff, slot0 = 0, cnt = 0;
            if (lane < nq) {
                const int2 iv = __ldg(&a.tiles.ints[q0 + lane]);
                const int s0 = a.start[iv.x], s1 = a.start[iv.y];
                lo = iv.x;
                slot0 = s0 & ~1;                                   // uv records are 8 B: keep both arrays 16-byte aligned
                cnt = s1 > s0 ? ((s1 + 1) & ~1) - slot0 : 0;
            }
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const bool staged = cnt > 0 && incl <= TILE_CAP;
            const int off = incl - cnt;
            __syncwarp();   // every lane is done with the previous tile's staged data and interval table
            wt.ilo[lane] = lo;
            wt.islot[lane] = slot0;
            wt.ioff[lane] = staged ? off : -1;
            const unsigned total = __reduce_add_sync(0xffffffffu, staged ? (unsigned)cnt * 24u : 0u);
            armed = total > 0;
            if (lane == 0 && armed) mbar_arrive_expect_tx(&wt.bar, total);
            __syncwarp();
            if (staged) {
                bulk_g2s(wt.pos + off, a.cur.pos + slot0, (unsigned)cnt * 16u, &wt.bar);
                bulk_g2s(wt.uv + off, a.cur.uv + slot0, (unsigned)cnt * 8u, &wt.bar);
            }
        }
        bool waited = false;

        for (int base = T0; base < T1; base += 32) {
            const int i = base + lane;
            const bool resident = i < T1;
            const int4 ai = resident ? a.cur.aux[i] : make_int4(0, 0, 0, ORIGIN_DEAD);
            const bool live = resident && ai.w >= 0;   // slab mode: halo copies are read by others, never advanced
            if (resident && !live) {
                a.alt.aux[i] = make_int4(0, -1, ai.z, ORIGIN_DEAD);
                a.key[i] = KEY_DROP;
            }
            float4 Pi = make_float4(0.f, 0.f, 0.f, 0.f);
            Real2<R> ui = {0.0f, 0.0f};
            int nr = 0;
            if (live) {
                const Pos3<R> P = a.cur.pos[i];
                Pi = make_float4(P.x, P.y, P.z, P.w);
                ui = a.cur.uv[i];
                int rb[9], rl[9];
                int c[3];
                cell_coords<R>(a.vox, P, c);
#pragma unroll
                for (int m = 0; m < 9; ++m) {
                    int b, e, clo;
                    row_range_lo<R>(a, c[0], c[1] + (m % 3) - 1, c[2] + (m / 3) - 1, b, e, clo);
                    rl[m] = e - b;
                    rb[m] = tiled_lookup(wt, clo, b);
                }
#define T2D_CSWAP(x, y)                                \
    if (rl[x] < rl[y]) {                               \
        int t_ = rl[x]; rl[x] = rl[y]; rl[y] = t_;     \
        t_ = rb[x]; rb[x] = rb[y]; rb[y] = t_;         \
    }
                // 9-input sorting network (25 compare-exchanges): longest range first, so the lanes of a warp finish
                // their m-th range at about the same time
                T2D_CSWAP(0, 1) T2D_CSWAP(3, 4) T2D_CSWAP(6, 7) T2D_CSWAP(1, 2) T2D_CSWAP(4, 5) T2D_CSWAP(7, 8)
                T2D_CSWAP(0, 1) T2D_CSWAP(3, 4) T2D_CSWAP(6, 7) T2D_CSWAP(0, 3) T2D_CSWAP(3, 6) T2D_CSWAP(0, 3)
                T2D_CSWAP(1, 4) T2D_CSWAP(4, 7) T2D_CSWAP(1, 4) T2D_CSWAP(2, 5) T2D_CSWAP(5, 8) T2D_CSWAP(2, 5)
                T2D_CSWAP(1, 3) T2D_CSWAP(5, 7) T2D_CSWAP(2, 6) T2D_CSWAP(4, 6) T2D_CSWAP(2, 4) T2D_CSWAP(2, 3)
                T2D_CSWAP(5, 6)
#undef T2D_CSWAP
#pragma unroll
                for (int m = 0; m < 9; ++m) {
                    wt.rb[m][lane] = rb[m];
                    wt.rl[m][lane] = rl[m];
                    nr += rl[m] > 0 ? 1 : 0;
                }
                const int ob = a.start[M], ol = a.start[M + 1] - ob;   // overflow bucket: normally empty, always global
                if (ol > 0) {
                    wt.rb[nr][lane] = ~ob;
                    wt.rl[nr][lane] = ol;
                    nr++;
                }
            }
            __syncwarp();
            if (!waited) {   // the tile's candidates have landed (once per tile)
                if (armed) mbar_wait(&wt.bar, parity);
                if (!trig_ready) {
                    mbar_wait(&sm.bar_trig, 0);
                    trig_ready = true;
                }
                waited = true;
            }
            PairAcc acc;
            int color = 0, hits = 0;
            unsigned ncut = 0;
            if (live) {
#pragma unroll 1
                for (int m = 0; m < nr; ++m) {
                    const int jb = wt.rb[m][lane], len = wt.rl[m][lane];
#if T2D_TILE_GENERIC
                    // one copy of the loop, generic loads: the pointer is in shared memory when the range was staged
                    const float4* cp = jb >= 0 ? wt.pos + jb : reinterpret_cast<const float4*>(a.cur.pos) + ~jb;
                    const float2* cu = jb >= 0 ? wt.uv + jb : reinterpret_cast<const float2*>(a.cur.uv) + ~jb;
                    cand_loop<TIES>(cp, cu, len, Pi, ui, k, sm.trig, a.trig_d, acc, color, hits, ncut, trig_fb);
#else
                    if (jb >= 0)
                        cand_loop<TIES>(wt.pos + jb, wt.uv + jb, len, Pi, ui, k, sm.trig, a.trig_d, acc, color, hits, ncut, trig_fb);
                    else
                        cand_loop<TIES>(reinterpret_cast<const float4*>(a.cur.pos) + ~jb, reinterpret_cast<const float2*>(a.cur.uv) + ~jb,
                                        len, Pi, ui, k, sm.trig, a.trig_d, acc, color, hits, ncut, trig_fb);
#endif
                }
            }
            __syncwarp();   // reconverge: lanes leave the candidate loops at different times, the tail is the same for all
            unsigned npairs = 0, nties = 0;
            if (live) {
                const int nh = (int)Pi.w;
                double2 own;
                if ((unsigned)nh < (unsigned)TRIG_SMEM_N)
                    own = sm.trig[nh];
                else
                    own = trig_lookup(a.trig_d, nh, trig_fb);
                fast_epilogue<true>(a, i, ai, ui, own, acc, color, hits, npairs, nties);
            }
            npairs_w += npairs;
            nties_w += nties;
            ncut_w += ncut;
        }
        if (armed) parity ^= 1u;
    }
    // diagnostic counters: one atomic per warp and counter
    npairs_w = __reduce_add_sync(0xffffffffu, npairs_w);
    nties_w = __reduce_add_sync(0xffffffffu, nties_w);
    ncut_w = __reduce_add_sync(0xffffffffu, ncut_w);
    const unsigned fb_w = __reduce_add_sync(0xffffffffu, (unsigned)trig_fb);
    if (lane == 0) {
        if (npairs_w) atomicAdd(&a.counters->pairs_in_range, (unsigned long long)npairs_w);
        if (nties_w) atomicAdd(&a.counters->ties_trunc, (unsigned long long)nties_w);
        if (ncut_w) atomicAdd(&a.counters->ties_cutoff, (unsigned long long)ncut_w);
        if (fb_w) atomicAdd(&a.counters->trig_fallbacks, (unsigned long long)fb_w);
    }
}

template <typename R> bool Launch<R>::step_euclid_tiled(const StepArgs<R>& a, int sm_count, cudaStream_t s)
{
    if constexpr (sizeof(R) == 4) {
        if (!a.tiles.istart || a.tiles.ntiles <= 0) return false;
        int grid = sm_count * T2D_TILE_MIN_BLOCKS;
        if (grid > (a.tiles.ntiles + TILE_WARPS) / TILE_WARPS) grid = (a.tiles.ntiles + TILE_WARPS) / TILE_WARPS;
        cudaMemsetAsync(a.tiles.queue, 0, sizeof(int), s);
        if (a.count_ties)
            k_step_euclid_tiled<true><<<grid, TILE_THREADS, 0, s>>>(a);
        else
            k_step_euclid_tiled<false><<<grid, TILE_THREADS, 0, s>>>(a);
        return true;
    } else {
        return false;
    }
}

}  // namespace t2d
