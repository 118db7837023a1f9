// random_sector_bw.cu -- what HBM3e delivers for the search kernel's access pattern:
// independent random reads of one aligned 64-byte bucket (one 32-byte-sector pair), no
// dependency between reads, as many in flight as the SMs can hold.  Prints GB/s for
// footprints well above L2.  Diagnostic only (profiles/gpu_session.sh runs it); the
// roofline denominator stays the streaming number in MEASURED_PEAKS.json, this figure says
// how much of it a random-sector workload can reach at all.
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

// each thread reads `per` random buckets; `bytes` = 64 (cnt + both halves), 40 (what s3_rank4 touches) or 32
template <int BYTES>
__global__ void gather(const uint4 *__restrict__ buckets, uint32_t numBuckets, uint32_t per, uint32_t seed, uint32_t *sink)
{
    uint32_t x = mix(seed + blockIdx.x * blockDim.x + threadIdx.x);
    uint32_t acc = 0;
#pragma unroll 4
    for (uint32_t k = 0; k < per; ++k) {
        x = mix(x + k);
        const uint4 *p = buckets + (size_t)(x % numBuckets) * 4;
        uint4 a = __ldg(p);
        acc += a.x ^ a.w;
        if (BYTES >= 40) { uint4 b = __ldg(p + 1); uint2 c = __ldg(reinterpret_cast<const uint2 *>(p) + 4); acc += b.y ^ c.x; }
        if (BYTES >= 64) { uint4 d = __ldg(p + 3); acc += d.z; }
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int BYTES>
static double run(const uint4 *d, uint32_t numBuckets, uint32_t *sink, int blocks, int threads, uint32_t per)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<BYTES><<<blocks, threads>>>(d, numBuckets, per, 1, sink);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0);
        gather<BYTES><<<blocks, threads>>>(d, numBuckets, per, 77 + it, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double reads = (double)blocks * threads * per;
    return reads * 64.0 / (best * 1e-3) / 1e9;      // GB/s counted at 64 B per bucket read, like the roofline
}

int main(int argc, char **argv)
{
    const double gbs[] = {1.0, 2.0, 8.0};
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *sink; cudaMalloc(&sink, 4);
    for (double gb : gbs) {
        const size_t bytes = (size_t)(gb * (1u << 30));
        uint4 *d; if (cudaMalloc(&d, bytes) != cudaSuccess) { printf("alloc %.0f GB failed\n", gb); continue; }
        cudaMemset(d, 1, bytes);
        const uint32_t nb = (uint32_t)(bytes / 64);
        for (int threads : {256, 1024}) {
            const int blocks = sms * (2048 / threads) * 4;
            printf("{\"footprint_gb\": %.0f, \"threads_per_block\": %d, \"gbs_at_64B_per_read\": {\"touch16B\": %.0f, \"touch40B\": %.0f, \"touch64B\": %.0f}}\n",
                   gb, threads, run<16>(d, nb, sink, blocks, threads, 256), run<40>(d, nb, sink, blocks, threads, 256),
                   run<64>(d, nb, sink, blocks, threads, 256));
        }
        cudaFree(d);
    }
    return 0;
}
