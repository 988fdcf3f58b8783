// common.cu — precision-independent kernels: exclusive scan of the bucket histogram, observables.
#include "t2d_internal.h"

namespace t2d {

// ---------------------------------------------------------------------------------------------------
// exclusive scan of count[0..M) into start[0..M], zeroing count for the next step.
// Three launches: per-tile sums -> scan of the tile sums (one block) -> per-tile scan + offset.
// A tile is 256 threads x 16 ints (four 16-byte loads per thread, all issued before the first use: the kernels
// are pure streaming, so memory-level parallelism per thread is what reaches HBM bandwidth).
// The allocator pads count/start to a multiple of 4 ints.
// ---------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_VEC = 4;                           // int4 loads per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_VEC * 4;

int scan_blocks(int M) { return (M + SCAN_TILE - 1) / SCAN_TILE; }

__device__ __forceinline__ int warp_incl_scan(int v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// inclusive scan across the block; returns this thread's inclusive value and the block total
__device__ __forceinline__ int block_incl_scan(int v, int* total)
{
    __shared__ int s_w[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = warp_incl_scan(v);
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
        int t = (lane < (int)(blockDim.x >> 5)) ? s_w[lane] : 0;
        t = warp_incl_scan(t);
        s_w[lane] = t;
    }
    __syncthreads();
    int base = (w > 0) ? s_w[w - 1] : 0;
    *total = s_w[(blockDim.x >> 5) - 1];
    __syncthreads();
    return inc + base;
}

// this thread's SCAN_VEC int4 of the tile (vector v covers ints [base + (v * SCAN_THREADS + tid) * 4, +4)): coalesced
__device__ __forceinline__ void tile_load(const int* __restrict__ count, int M, int tile, int4 q[SCAN_VEC])
{
#pragma unroll
    for (int v = 0; v < SCAN_VEC; ++v) {
        const int base = tile * SCAN_TILE + (v * SCAN_THREADS + threadIdx.x) * 4;
        if (base + 3 < M) {
            q[v] = *reinterpret_cast<const int4*>(count + base);
        } else {
            int t[4] = {0, 0, 0, 0};
            for (int k = 0; k < 4; ++k)
                if (base + k < M) t[k] = count[base + k];
            q[v] = make_int4(t[0], t[1], t[2], t[3]);
        }
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const int* __restrict__ count, int* __restrict__ sums, int M)
{
    int4 q[SCAN_VEC];
    tile_load(count, M, blockIdx.x, q);
    int v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_VEC; ++k) v += q[k].x + q[k].y + q[k].z + q[k].w;
    int total;
    block_incl_scan(v, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_sums(int* sums, int nb)
{
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        int i = b0 + threadIdx.x;
        int v = (i < nb) ? sums[i] : 0;
        int total;
        int inc = block_incl_scan(v, &total);
        int c = carry;
        if (i < nb) sums[i] = c + inc - v;   // exclusive
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[nb] = carry;   // grand total
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(int* __restrict__ count, int* __restrict__ start,
                                                             const int* __restrict__ sums, int M, int nb)
{
    int4 q[SCAN_VEC];
    tile_load(count, M, blockIdx.x, q);
    // vector v of the tile is scanned as a row of SCAN_THREADS * 4 ints; rows are chained through `carry`
    int carry = sums[blockIdx.x];
#pragma unroll
    for (int v = 0; v < SCAN_VEC; ++v) {
        const int base = blockIdx.x * SCAN_TILE + (v * SCAN_THREADS + threadIdx.x) * 4;
        const int s = q[v].x + q[v].y + q[v].z + q[v].w;
        int total;
        const int inc = block_incl_scan(s, &total);
        int4 o;
        o.x = carry + inc - s;
        o.y = o.x + q[v].x;
        o.z = o.y + q[v].y;
        o.w = o.z + q[v].z;
        if (base + 3 < M) {
            *reinterpret_cast<int4*>(start + base) = o;
            *reinterpret_cast<int4*>(count + base) = make_int4(0, 0, 0, 0);
        } else {
            const int t[4] = {o.x, o.y, o.z, o.w};
            for (int k = 0; k < 4; ++k)
                if (base + k < M) {
                    start[base + k] = t[k];
                    count[base + k] = 0;
                }
        }
        carry += total;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) start[M] = sums[nb];
}

void launch_scan(int* count, int* start, int* blocksums, int M, cudaStream_t s)
{
    const int nb = scan_blocks(M);
    k_scan_tile_sums<<<nb, SCAN_THREADS, 0, s>>>(count, blocksums, M);
    k_scan_sums<<<1, 1024, 0, s>>>(blocksums, nb);
    k_scan_apply<<<nb, SCAN_THREADS, 0, s>>>(count, start, blocksums, M, nb);
}

// ---------------------------------------------------------------------------------------------------
// The same scan in ONE launch (decoupled look-back): a tile publishes {launch number, state, value} as one 64-bit word —
// state 1 = the tile's own sum, 2 = its inclusive prefix — and looks back over its predecessors' words until it meets an
// inclusive one.  Tiles are numbered by a ticket counter in the order they start running, so a tile only ever waits for
// tiles that are already resident; the launch number in the word makes resetting the status array unnecessary, and
// `ticket_base` (the counter's value before this launch, tracked by the host) does the same for the counter.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_onepass(int* __restrict__ count, int* __restrict__ start,
                                                               unsigned long long* status, int* ticket, int ticket_base, unsigned seq,
                                                               int M, int nb)
{
    __shared__ int s_tile, s_prefix;
    pdl_wait();
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1) - ticket_base;
    __syncthreads();
    const int tile = s_tile;
    // thread t owns the 16 consecutive buckets [base, base + 16) of the tile: a serial scan in registers, ONE block scan of the
    // per-thread totals (the three-launch version chains four block scans per tile)
    const int base = tile * SCAN_TILE + threadIdx.x * (SCAN_VEC * 4);
    int4 q[SCAN_VEC];
#pragma unroll
    for (int k = 0; k < SCAN_VEC; ++k) {
        const int b0 = base + 4 * k;
        if (b0 + 3 < M) {
            q[k] = *reinterpret_cast<const int4*>(count + b0);
        } else {
            int t[4] = {0, 0, 0, 0};
            for (int j = 0; j < 4; ++j)
                if (b0 + j < M) t[j] = count[b0 + j];
            q[k] = make_int4(t[0], t[1], t[2], t[3]);
        }
    }
    int v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_VEC; ++k) v += q[k].x + q[k].y + q[k].z + q[k].w;
    int total;
    const int incl = block_incl_scan(v, &total);
    volatile unsigned long long* st = status;
    const unsigned long long tag = (unsigned long long)(seq & 0x3fffffffu) << 34;
    if (threadIdx.x == 0) st[tile] = tag | ((unsigned long long)(tile == 0 ? 2u : 1u) << 32) | (unsigned)total;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int excl = 0;
        for (int lb = tile - 1; lb >= 0; lb -= 32) {
            const int idx = lb - lane;
            unsigned long long w = tag | (2ull << 32);   // before the first tile: inclusive, value 0
            if (idx >= 0) {
                do {
                    w = st[idx];
                } while ((w >> 34) != (tag >> 34) || ((w >> 32) & 3u) == 0u);
            }
            const unsigned inclm = __ballot_sync(0xffffffffu, ((w >> 32) & 3u) == 2u);
            const int last = inclm ? __ffs(inclm) - 1 : 31;   // nearest predecessor that already knows its inclusive prefix
            int val = lane <= last ? (int)(unsigned)w : 0;
            for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(0xffffffffu, val, o);
            excl += __shfl_sync(0xffffffffu, val, 0);
            if (inclm) break;
        }
        if (lane == 0) {
            s_prefix = excl;
            if (tile > 0) st[tile] = tag | (2ull << 32) | (unsigned)(excl + total);
        }
    }
    __syncthreads();
    int run = s_prefix + incl - v;   // exclusive prefix of this thread's first bucket
#pragma unroll
    for (int k = 0; k < SCAN_VEC; ++k) {
        const int b0 = base + 4 * k;
        int4 o;
        o.x = run;
        o.y = o.x + q[k].x;
        o.z = o.y + q[k].y;
        o.w = o.z + q[k].z;
        run = o.w + q[k].w;
        if (b0 + 3 < M) {
            *reinterpret_cast<int4*>(start + b0) = o;
            *reinterpret_cast<int4*>(count + b0) = make_int4(0, 0, 0, 0);
        } else {
            const int t[4] = {o.x, o.y, o.z, o.w};
            for (int j = 0; j < 4; ++j)
                if (b0 + j < M) {
                    start[b0 + j] = t[j];
                    count[b0 + j] = 0;
                }
        }
    }
    if (tile == nb - 1 && threadIdx.x == 0) start[M] = s_prefix + total;
}

void launch_scan_onepass(int* count, int* start, unsigned long long* status, int* ticket, int ticket_base, unsigned seq, int M,
                         cudaStream_t s, bool pdl)
{
    const int nb = scan_blocks(M);
    if (nb > 0) launch_pdl(pdl, k_scan_onepass, (unsigned)nb, (unsigned)SCAN_THREADS, s, count, start, status, ticket, ticket_base, seq, M, nb);
}

// ---------------------------------------------------------------------------------------------------
// observables: sum cos n, sum sin n, sum |rdot| over the resident particles
// ---------------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(256) k_observables(const Pos3<R>* __restrict__ pos, const Real2<R>* __restrict__ rdot,
                                                     const int4* __restrict__ aux, int N, const int* __restrict__ dN,
                                                     const double2* __restrict__ trig, double* out)
{
    double sc = 0, ss = 0, sp = 0, cnt = 0;
    if (dN) N = *dN;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        if (aux && aux[i].w < 0) continue;   // halo copy
        cnt += 1.0;
        int n = (int)pos[i].w;
        double c, s;
        if (n >= TRIG_MIN && n <= TRIG_MAX) {
            double2 t = trig[n - TRIG_MIN];
            c = t.x;
            s = t.y;
        } else {
            double r = (double)n * DEG_TO_RAD_D;
            c = cos(r);
            s = sin(r);
        }
        sc += c;
        ss += s;
        double rx = (double)rdot[i].x, ry = (double)rdot[i].y;
        sp += sqrt(rx * rx + ry * ry);
    }
    __shared__ double sh[4][8];
    for (int o = 16; o > 0; o >>= 1) {
        sc += __shfl_down_sync(0xffffffffu, sc, o);
        ss += __shfl_down_sync(0xffffffffu, ss, o);
        sp += __shfl_down_sync(0xffffffffu, sp, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
        sh[0][w] = sc;
        sh[1][w] = ss;
        sh[2][w] = sp;
        sh[3][w] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0, d = 0;
        for (int k = 0; k < 8; ++k) {
            a += sh[0][k];
            b += sh[1][k];
            c += sh[2][k];
            d += sh[3][k];
        }
        atomicAdd(&out[T2D_OBS_SUM_COS], a);
        atomicAdd(&out[T2D_OBS_SUM_SIN], b);
        atomicAdd(&out[T2D_OBS_SUM_SPEED], c);
        atomicAdd(&out[T2D_OBS_COUNT], d);
    }
}

void launch_observables(const void* pos, const void* rdot, const int4* aux, int is_f32, int N, const int* dN,
                        const double2* trig, double* out8, cudaStream_t s)
{
    cudaMemsetAsync(out8, 0, sizeof(double) * T2D_OBS_LEN, s);
    if (N <= 0) return;
    int grid = (N + 255) / 256;
    if (grid > 1184) grid = 1184;   // 148 SMs x 8 resident blocks
    if (is_f32)
        k_observables<float><<<grid, 256, 0, s>>>((const Pos3<float>*)pos, (const Real2<float>*)rdot, aux, N, dN, trig, out8);
    else
        k_observables<double><<<grid, 256, 0, s>>>((const Pos3<double>*)pos, (const Real2<double>*)rdot, aux, N, dN, trig, out8);
}

}  // namespace t2d
