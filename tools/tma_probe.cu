// tma_probe.cu — micro-benchmark behind DESIGN.md §4: how fast does one SM turn around SMALL cp.async.bulk
// (global -> shared, 1-D, mbarrier completion) copies?  The tiled step kernel stages 10-30 candidate intervals of
// 0.1-3 KB per tile, so the per-request cost matters more than bandwidth.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu
// Prints: bytes per copy, copies per round, CTAs/SM, ns per round, copies/us/SM, GB/s.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned n)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned phase)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT;\n\tDONE:\n\t}" ::"r"(
            smem_u32(b)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* b)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}

__global__ void __launch_bounds__(128) k_probe(const unsigned char* src, size_t src_bytes, int copies, int bytes, int rounds,
                                               unsigned* sink)
{
    extern __shared__ __align__(128) unsigned char buf[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned acc = 0;
    unsigned h = blockIdx.x * 2654435761u + 12345u;
    for (int r = 0; r < rounds; ++r) {
        if (threadIdx.x == 0) mbar_expect_tx(&bar, (unsigned)(copies * bytes));
        if (threadIdx.x < copies) {
            unsigned hh = (h + threadIdx.x * 40503u + r * 2246822519u) * 2654435761u;
            size_t off = ((size_t)hh * 16) % (src_bytes - (size_t)bytes - 16);
            off &= ~(size_t)15;
            bulk_g2s(buf + (size_t)threadIdx.x * bytes, src + off, (unsigned)bytes, &bar);
        }
        mbar_wait(&bar, r & 1);
        acc += buf[(threadIdx.x * 16) % (copies * bytes)];
        __syncthreads();
    }
    if (acc == 0xdeadbeef) *sink = acc;
}

int main()
{
    const size_t SRC = 256u << 20;
    unsigned char* d;
    unsigned* sink;
    cudaMalloc(&d, SRC);
    cudaMemset(d, 1, SRC);
    cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int rounds = 400;
    printf("bytes copies cta_per_sm  ns_per_round  copies_per_us_per_sm  GBps_total\n");
    for (int bytes : {96, 384, 1536, 6144})
        for (int copies : {4, 16, 32})
            for (int cps : {1, 4, 8}) {
                size_t smem = (size_t)copies * bytes;
                if (smem * cps > 200 * 1024 || smem > 100 * 1024) continue;
                int grid = 148 * cps;
                k_probe<<<grid, 128, smem>>>(d, SRC, copies, bytes, 20, sink);
                cudaEventRecord(e0);
                k_probe<<<grid, 128, smem>>>(d, SRC, copies, bytes, rounds, sink);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                cudaError_t err = cudaGetLastError();
                if (err != cudaSuccess) {
                    printf("error %s\n", cudaGetErrorString(err));
                    return 1;
                }
                double ns_round = ms * 1e6 / rounds;
                double cpus = (double)copies * cps / (ns_round * 1e-3);
                double gbs = (double)grid * copies * bytes * rounds / (ms * 1e-3) / 1e9;
                printf("%5d %5d %5d %12.1f %12.2f %12.1f\n", bytes, copies, cps, ns_round, cpus, gbs);
            }
    return 0;
}
