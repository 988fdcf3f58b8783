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
