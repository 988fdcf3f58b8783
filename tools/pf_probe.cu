// dev probe: does prefetch.global.L1 / .L2 shorten the latency of a later load on sm_100a?  (nvcc -arch=sm_100a pf_probe.cu)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(const float4* data, size_t n, int mode, long long* out, float* sink)
{
    // one warp; every lane reads a different 4 KB-strided line that nobody touched before
    const int lane = threadIdx.x;
    float acc = 0;
    long long tot = 0;
    for (int it = 0; it < 64; ++it) {
        const float4* p = data + ((size_t)(it * 32 + lane) * 4096 + (size_t)blockIdx.x * 64 * 32 * 4096) % n;
        if (mode == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
        if (mode == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        if (mode == 3) asm volatile("cp.async.bulk.prefetch.L2.global [%0], 32;" ::"l"(p));
        // ~3000 cycles of unrelated work
        float x = (float)lane;
        for (int k = 0; k < 600; ++k) x = x * 1.0001f + 0.5f;
        acc += x;
        const long long t0 = clock64();
        float vx, vy, vz, vw;
        asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(vx), "=f"(vy), "=f"(vz), "=f"(vw) : "l"(p) : "memory");
        acc += vx + vw;
        const long long t1 = clock64();
        tot += t1 - t0;
    }
    sink[blockIdx.x * 32 + lane] = acc;
    if (lane == 0) out[blockIdx.x] = tot / 64;
}
int main()
{
    const size_t n = (size_t)1 << 28;   // 4 GB of float4
    float4* d;
    cudaMalloc(&d, n * sizeof(float4));
    cudaMemset(d, 0, n * sizeof(float4));
    long long* out;
    float* sink;
    cudaMalloc(&out, 64 * sizeof(long long));
    cudaMalloc(&sink, 64 * 32 * sizeof(float));
    for (int mode = 0; mode < 4; ++mode) {
        probe<<<4, 32>>>(d + (size_t)mode * (n / 4), n / 4, mode, out, sink);
        long long h[4];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("mode %d (0 none, 1 prefetch.L1, 2 prefetch.L2, 3 bulk prefetch L2): load latency %lld %lld %lld %lld cycles  (%s)\n", mode, h[0],
               h[1], h[2], h[3], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
