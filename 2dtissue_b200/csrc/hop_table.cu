// hop_table.cu — the vertex-distance table of stage 2 built on the GPU.
//
// Replaces DijkstraDistanceHelper::get_mesh_distance_matrix / calculate_edge_count_distance
// (/root/reference/MeshCartographyLib/src/GeodesicDistance/DijkstraDistanceHelper.cpp:27-111): Dijkstra with unit
// edge weights from every vertex of <stem>_open.off == BFS hop counts over the mesh's edge graph.
// One CTA per source vertex, the distance row lives in shared memory (uint8, 255 = unreached / > 254 hops),
// level-synchronous sweeps; the row is written once, coalesced.
#include "t2d_internal.h"

namespace t2d {

__global__ void __launch_bounds__(256) k_hop_bfs(int V, const int* __restrict__ adj_start, const int* __restrict__ adj,
                                                 uint8_t* __restrict__ out, int src0, int nsrc)
{
    extern __shared__ uint8_t dist[];
    __shared__ int changed;
    for (int sidx = blockIdx.x; sidx < nsrc; sidx += gridDim.x) {
        const int src = src0 + sidx;
        for (int v = threadIdx.x; v < V; v += blockDim.x) dist[v] = 255;
        __syncthreads();
        if (threadIdx.x == 0) dist[src] = 0;
        __syncthreads();
        for (int level = 0; level < 254; ++level) {
            if (threadIdx.x == 0) changed = 0;
            __syncthreads();
            int mine = 0;
            for (int v = threadIdx.x; v < V; v += blockDim.x) {
                if (dist[v] == level) {
                    for (int q = adj_start[v]; q < adj_start[v + 1]; ++q) {
                        int w = adj[q];
                        if (dist[w] == 255) {   // benign race: every writer stores level + 1
                            dist[w] = (uint8_t)(level + 1);
                            mine = 1;
                        }
                    }
                }
            }
            if (mine) changed = 1;
            __syncthreads();
            int c = changed;
            __syncthreads();
            if (!c) break;
        }
        uint8_t* row = out + (size_t)src * V;
        for (int v = threadIdx.x; v < V; v += blockDim.x) row[v] = dist[v];
        __syncthreads();
    }
}

// The same BFS cut off after `hops` levels, emitted as CSR rows (ascending vertex id, diagonal included) instead of a dense
// row: what stage 2 reads of the table is only d < 2 sigma and d <= color_factor sigma, a handful of vertices per row, while
// the dense table of a refined chart (V = 75 k) would be 5.6 GB as uint8 and 45 GB in the reference's double format.
// Pass 1 (start == nullptr) writes the row lengths, pass 2 the columns and hop counts at start[src].
__global__ void __launch_bounds__(256) k_hop_csr(int V, const int* __restrict__ adj_start, const int* __restrict__ adj, int hops,
                                                 int* __restrict__ row_len, const int* __restrict__ start, int* __restrict__ col,
                                                 uint8_t* __restrict__ val)
{
    extern __shared__ uint8_t dist[];
    __shared__ int changed, s_base, s_warp[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int src = blockIdx.x; src < V; src += gridDim.x) {
        for (int v = threadIdx.x; v < V; v += blockDim.x) dist[v] = 255;
        __syncthreads();
        if (threadIdx.x == 0) dist[src] = 0;
        __syncthreads();
        for (int level = 0; level < hops; ++level) {
            if (threadIdx.x == 0) changed = 0;
            __syncthreads();
            int mine = 0;
            for (int v = threadIdx.x; v < V; v += blockDim.x) {
                if (dist[v] == level) {
                    for (int q = adj_start[v]; q < adj_start[v + 1]; ++q) {
                        const int u = adj[q];
                        if (dist[u] == 255) {   // benign race: every writer stores level + 1
                            dist[u] = (uint8_t)(level + 1);
                            mine = 1;
                        }
                    }
                }
            }
            if (mine) changed = 1;
            __syncthreads();
            const int c = changed;
            __syncthreads();
            if (!c) break;
        }
        // ordered emission: chunks of 256 vertices, ascending; a running base keeps the row sorted
        if (threadIdx.x == 0) s_base = 0;
        __syncthreads();
        for (int v0 = 0; v0 < V; v0 += blockDim.x) {
            const int v = v0 + threadIdx.x;
            const bool hit = v < V && dist[v] != 255;
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_warp[w] = __popc(m);
            __syncthreads();
            int off = s_base;
            for (int k = 0; k < w; ++k) off += s_warp[k];
            if (hit && start) {
                const int at = start[src] + off + __popc(m & ((1u << lane) - 1u));
                col[at] = v;
                val[at] = dist[v];
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int t = 0;
                for (int k = 0; k < 8; ++k) t += s_warp[k];
                s_base += t;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0 && row_len) row_len[src] = s_base;
        __syncthreads();
    }
}

int launch_hop_csr(int V, const int* d_adj_start, const int* d_adj, int hops, int* row_len, const int* start, int* col, uint8_t* val,
                   int sm_count, cudaStream_t s)
{
    size_t smem = (size_t)V;
    if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k_hop_csr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = (int)((220 * 1024) / (smem + 2048));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int grid = sm_count * per_sm;
    if (grid > V) grid = V;
    k_hop_csr<<<grid, 256, smem, s>>>(V, d_adj_start, d_adj, hops, row_len, start, col, val);
    return (int)cudaGetLastError();
}

// out_dev: uint8[V][V] on the device.  Returns cudaError as int.
int launch_hop_table(int V, const int* d_adj_start, const int* d_adj, uint8_t* out_dev, int sm_count, cudaStream_t s)
{
    size_t smem = (size_t)V;
    if (smem > 200 * 1024) return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k_hop_bfs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int grid = sm_count * per_sm;
    if (grid > V) grid = V;
    k_hop_bfs<<<grid, 256, smem, s>>>(V, d_adj_start, d_adj, out_dev, 0, V);
    return (int)cudaGetLastError();
}

}  // namespace t2d
