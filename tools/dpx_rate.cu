// dpx_rate.cu -- issue rate of the 16x2 DPX instructions the DP kernel is built from, measured:
// independent chains of VIADDMNMX.S16x2 (max(a+b,c)), VIMNMX3.S16x2 (max3), VIADD.16x2 and PRMT,
// every SM full of warps.  Prints warp-instructions per clock per SM and the resulting
// ceiling in cell updates per second (GCUPS) for an N-instruction cell pair:
//     GCUPS_peak(N) = rate[warp-inst/clk/SM] * SMs * clock * 32 lanes * 2 cells / N
// The DP roofline of bench.py uses N = 5 (E, F, H(add), H(max3), best) as the
// denominator and reports the kernel's own ALU-pipe instruction count per cell pair next to it.
// Diagnostic only.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dpx_rate dpx_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CHAINS 8
#define ITERS 4096

template <int OP>
__global__ void __launch_bounds__(256) spin(uint32_t seed, uint32_t *sink)
{
    uint32_t v[CHAINS];
    const uint32_t b = seed * 0x00010001u + threadIdx.x, c = 0x83008300u;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) v[k] = seed + k * 0x00030005u + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) {
            if (OP == 0) v[k] = __viaddmax_s16x2(v[k], b, c);
            else if (OP == 1) v[k] = __vimax3_s16x2(v[k], b, c);
            else if (OP == 2) v[k] = __vadd2(v[k], b);
            else if (OP == 3) v[k] = __vmaxs2(v[k], b);
            else if (OP == 4) v[k] = __byte_perm(v[k], b, v[k] & 0x7777u);
            else v[k] = v[k] * b + c;                            // IMAD (fma pipe), for comparison
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) acc ^= v[k];
    if (acc == 0x12345678u) *sink = acc;
}

template <int OP>
static double rate(int sms, double clockHz, uint32_t *sink)
{
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    spin<OP><<<blocks, threads>>>(1, sink);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0);
        spin<OP><<<blocks, threads>>>(2 + it, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double warpInst = (double)blocks * (threads / 32) * ITERS * CHAINS;
    return warpInst / (best * 1e-3) / clockHz / sms;             // warp-instructions per clock per SM
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clkKHz = 0;
    cudaDeviceGetAttribute(&clkKHz, cudaDevAttrClockRate, 0);
    const double clk = clkKHz * 1e3;
    uint32_t *sink;
    cudaMalloc(&sink, 4);
    const char *names[6] = {"VIADDMNMX.S16x2 (__viaddmax_s16x2)", "VIMNMX3.S16x2 (__vimax3_s16x2)", "VIADD.16x2 (__vadd2)",
                            "VIMNMX.S16x2 (__vmaxs2)", "PRMT (__byte_perm)", "IMAD (fma pipe)"};
    double r[6];
    r[0] = rate<0>(p.multiProcessorCount, clk, sink);
    r[1] = rate<1>(p.multiProcessorCount, clk, sink);
    r[2] = rate<2>(p.multiProcessorCount, clk, sink);
    r[3] = rate<3>(p.multiProcessorCount, clk, sink);
    r[4] = rate<4>(p.multiProcessorCount, clk, sink);
    r[5] = rate<5>(p.multiProcessorCount, clk, sink);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f", p.name, p.multiProcessorCount, clk / 1e6);
    for (int k = 0; k < 6; ++k) printf(", \"%s\": %.3f", names[k], r[k]);
    const double dpx = r[0] < r[1] ? r[0] : r[1];
    printf(", \"warp_inst_per_clk_per_sm_dpx\": %.3f, \"gcups_peak_5op\": %.1f, \"gcups_peak_9op\": %.1f}\n", dpx,
           dpx * p.multiProcessorCount * clk * 64 / 5 / 1e9, dpx * p.multiProcessorCount * clk * 64 / 9 / 1e9);
    return 0;
}
