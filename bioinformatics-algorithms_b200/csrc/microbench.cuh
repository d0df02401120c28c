// microbench.cuh -- sustained issue rate of the packed int16x2 DPX instructions on this GPU.
// the integer-ALU roofline is not in MEASURED_PEAKS.json and must be measured.
// Independent chains (8-way ILP per thread), every SM full, CUDA-event timed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2a {

// KIND 0: VIADDMNMX.S16x2 only.  KIND 1: the SW fill's arithmetic (PRMT, VIADD.16x2, 2x VIADDMNMX).
// KIND 2: KIND 1 plus one IMAD per cell pair (the FMA-pipe delta-word update).
// KIND 3-5: VIMNMX.S16x2, VIMNMX3.S16x2, VIADD.16x2.  KIND 6-9: PRMT, LOP3 (xor), IADD3, IMAD alone.
// KIND 10: the NW cell pair as the kernel issues it (PRMT, VIADDMNMX.S16x2, VIMNMX.S16x2, IMAD).
// KIND 11: the SW cell pair as the kernel issues it (PRMT, VIADD.16x2, VIADDMNMX, VIADDMNMX.RELU, VIMNMX.S16x2, IMAD).
// The mixes are the INSTRUCTION-MIX CEILINGS of the fill kernels: no loads, shuffles, stores or loop overhead.
template <int KIND>
__global__ void __launch_bounds__(256) microbench_kernel(uint32_t* out, int iters, uint32_t g, uint32_t sel, uint32_t radix, uint32_t x)
{
    uint32_t h[8], acc[8], best[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { h[c] = threadIdx.x * 65537u + c; acc[c] = c; best[c] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (KIND == 0) {
                h[c] = __viaddmax_s16x2(h[c], g, x);
            } else if (KIND == 3) {
                h[c] = __vmaxs2(h[c], x + it);                  // VIMNMX.S16x2 (plain packed max)
            } else if (KIND == 4) {
                h[c] = __vimax3_s16x2(h[c], x, g + it);         // VIMNMX3.S16x2
            } else if (KIND == 5) {
                h[c] = __vadd2(h[c], x);                        // VIADD.16x2
            } else if (KIND == 6) {
                asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(h[c]) : "r"(h[c]), "r"(x), "r"(sel));
            } else if (KIND == 7) {
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(h[c]) : "r"(h[c]), "r"(x), "r"(g));
            } else if (KIND == 8) {
                asm volatile("add.u32 %0, %1, %2;" : "=r"(h[c]) : "r"(h[c]), "r"(x));
            } else if (KIND == 9) {
                h[c] = h[c] * radix + x;                        // IMAD
            } else if (KIND == 10) {
                uint32_t s;
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(s) : "r"(h[c]), "r"(x), "r"(sel));
                const uint32_t a = __viaddmax_s16x2(acc[c], s, h[c]);
                h[c] = __vmaxs2(a, x);
                acc[c] = acc[c] * radix + h[c];
            } else if (KIND == 11) {
                uint32_t s;
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(s) : "r"(h[c]), "r"(x), "r"(sel));
                const uint32_t ds = __vadd2(h[c], s);
                const uint32_t a = __viaddmax_s16x2(h[c], g, ds);
                h[c] = __viaddmax_s16x2_relu(x, g, a);
                best[c] = __vmaxs2(best[c], h[c]);
                acc[c] = acc[c] * radix + h[c];
            } else {
                uint32_t s;
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(s) : "r"(h[c]), "r"(x), "r"(sel));
                const uint32_t ds = __vadd2(h[c], s);
                const uint32_t a = __viaddmax_s16x2(h[c], g, ds);
                h[c] = __viaddmax_s16x2(x, g, a);
                if (KIND == 2) acc[c] = acc[c] * radix + h[c];
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) r ^= h[c] ^ acc[c] ^ best[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// gops = 1e9 lane-instructions per second (ALU-pipe instructions only, so KIND 2 shows whether the IMADs are free)
inline cudaError_t run_microbench(int kind, int sm_count, cudaStream_t st, double* gops, float* sm_mhz)
{
    const int threads = 256, blocks = sm_count * 8, iters = 4096;
    uint32_t* d = nullptr;
    cudaError_t e = cudaMalloc((void**)&d, (size_t)threads * blocks * 4);
    if (e != cudaSuccess) return e;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a, st);
        if (kind == 0) microbench_kernel<0><<<blocks, threads, 0, st>>>(d, iters, 0xFFFFFFFFu, 0x5140u, 4u, 0x00100010u);
        else if (kind == 1) microbench_kernel<1><<<blocks, threads, 0, st>>>(d, iters, 0xFFFFFFFFu, 0x5140u, 4u, 0x00100010u);
        else if (kind == 2) microbench_kernel<2><<<blocks, threads, 0, st>>>(d, iters, 0xFFFFFFFFu, 0x5140u, 4u, 0x00100010u);
        else if (kind == 3) microbench_kernel<3><<<blocks, threads, 0, st>>>(d, iters, 0xFFFFFFFFu, 0x5140u, 4u, 0x00100010u);
        else if (kind == 4) microbench_kernel<4><<<blocks, threads, 0, st>>>(d, iters, 0xFFFFFFFFu, 0x5140u, 4u, 0x00100010u);
        else if (kind == 5) microbench_kernel<5><<<blocks, threads, 0, st>>>(d, iters, 0xFFFFFFFFu, 0x5140u, 4u, 0x00100010u);
        else if (kind == 6) microbench_kernel<6><<<blocks, threads, 0, st>>>(d, iters, 0xFFFFFFFFu, 0x5140u, 4u, 0x00100010u);
        else if (kind == 7) microbench_kernel<7><<<blocks, threads, 0, st>>>(d, iters, 0xFFFFFFFFu, 0x5140u, 4u, 0x00100010u);
        else if (kind == 8) microbench_kernel<8><<<blocks, threads, 0, st>>>(d, iters, 0xFFFFFFFFu, 0x5140u, 4u, 0x00100010u);
        else if (kind == 9) microbench_kernel<9><<<blocks, threads, 0, st>>>(d, iters, 0xFFFFFFFFu, 0x5140u, 4u, 0x00100010u);
        else if (kind == 10) microbench_kernel<10><<<blocks, threads, 0, st>>>(d, iters, 0xFFFFFFFFu, 0x5140u, 4u, 0x00100010u);
        else if (kind == 11) microbench_kernel<11><<<blocks, threads, 0, st>>>(d, iters, 0xFFFFFFFFu, 0x5140u, 4u, 0x00100010u);
        else { cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d); return cudaErrorInvalidValue; }
        cudaEventRecord(b, st);
        e = cudaEventSynchronize(b);
        if (e != cudaSuccess) break;
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    // kinds 1, 2: ALU-pipe instructions (4 per cell pair); kinds 10, 11: CELL PAIRS (one per chain and iteration); else instructions
    const double per_iter = (kind == 1 || kind == 2) ? 32.0 : 8.0;
    *gops = (double)threads * blocks * iters * per_iter / (best * 1e-3) / 1e9;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    *sm_mhz = khz / 1000.0f;
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    return e;
}

} // namespace b2a
