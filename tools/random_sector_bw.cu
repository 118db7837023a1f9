// random_sector_bw.cu -- what HBM3e delivers for the search kernel's access pattern:
// independent random reads of one aligned 32-byte bucket (one sector, one LDG.E.256), no
// dependency between reads, as many in flight as the SMs can hold.  Prints million reads/s and
// the DRAM-side GB/s at 32 B (bytes used) and 64 B (bytes the memory system fetches per miss)
// for footprints well above L2.  Diagnostic only (profiles/gpu_session.sh runs it); the roofline
// denominator stays the streaming number in MEASURED_PEAKS.json, this figure says how much of it
// a random-sector workload can reach at all.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o random_sector_bw random_sector_bw.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__device__ __forceinline__ uint32_t ld256(const uint4 *p)
{
    uint32_t a, b, c, d, e, f, g, h;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

// each thread reads `per` random buckets, UNROLL of them in flight at a time
template <int UNROLL>
__global__ void gather(const uint4 *__restrict__ buckets, uint32_t mask, uint32_t per, uint32_t seed, uint32_t *sink)
{
    uint32_t x = mix(seed + blockIdx.x * blockDim.x + threadIdx.x);
    uint32_t acc = 0;
    for (uint32_t k = 0; k < per; k += UNROLL) {
        uint32_t v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) { x = mix(x + k + u); v[u] = ld256(buckets + (size_t)(x & mask) * 2); }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u];
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int UNROLL>
static double run(const uint4 *d, uint32_t mask, uint32_t *sink, int blocks, int threads, uint32_t per)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<UNROLL><<<blocks, threads>>>(d, mask, per, 1, sink);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0);
        gather<UNROLL><<<blocks, threads>>>(d, mask, per, 77 + it, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return (double)blocks * threads * per / (best * 1e-3);      // reads per second
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *sink; cudaMalloc(&sink, 4);
    for (int logb : {25, 26, 27}) {                              // 1, 2, 4 GiB of 32-byte buckets
        const size_t bytes = (size_t)32 << logb;
        uint4 *d; if (cudaMalloc(&d, bytes) != cudaSuccess) { printf("alloc failed\n"); continue; }
        cudaMemset(d, 1, bytes);
        const uint32_t mask = (1u << logb) - 1;
        for (int warpsPerSm : {24, 64}) {
            const int threads = 128, blocks = sms * warpsPerSm / 4;
            const double r1 = run<1>(d, mask, sink, blocks, threads, 512), r2 = run<2>(d, mask, sink, blocks, threads, 512),
                         r8 = run<8>(d, mask, sink, blocks, threads, 512);
            printf("{\"footprint_gib\": %.0f, \"warps_per_sm\": %d, \"mreads_per_s\": {\"1_in_flight\": %.0f, \"2_in_flight\": %.0f, \"8_in_flight\": %.0f}, "
                   "\"gbs_at_32B\": %.0f, \"gbs_at_64B\": %.0f}\n",
                   bytes / 1073741824.0, warpsPerSm, r1 / 1e6, r2 / 1e6, r8 / 1e6, r8 * 32 / 1e9, r8 * 64 / 1e9);
        }
        cudaFree(d);
    }
    return 0;
}
