// Issue-rate microbenchmarks for the pipes the field arithmetic of this engine lives on:
//   imad_wide : 32 x 32 + 64 -> 64 multiply-accumulate (mad.wide.u32 -> IMAD.WIDE, the unit DESIGN.md counts in)
//   imad_lo   : 32 x 32 + 32 -> 32 (IMAD)
//   dfma      : FP64 fused multiply-add (DFMA), the pipe the integer kernels leave idle
//   iadd3     : 32-bit integer add with carry (IADD3 / IADD3.X), the ALU pipe
//   mixes     : the same instruction streams interleaved in one warp - do the pipes overlap or share issue slots?
// Every thread runs ILP independent dependency chains (each instruction's multiplicand is its own accumulator, so
// nothing is loop-invariant), so a single warp can keep a pipe busy; the grid fills every SM
// with 8 warps per scheduler.  Output: one JSON object with giga warp-instructions per second per kind and the
// cycles per warp instruction per scheduler that implies at the measured SM clock.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/build/pipe_peaks scripts/pipe_peaks.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

static const int ILP = 8;
static const int ITERS = 65536;   // ~10-20 ms per kernel: long enough for boost clocks to settle

__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

__global__ void spin(long long cycles, long long* sink) {      // warm-up: brings the SM clock to its boost state
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = clock64() - t0;
}

template <int MODE>
__global__ void __launch_bounds__(256) bench(uint64_t* out, uint32_t seed, long long* clocks) {
    uint64_t acc[ILP];
    double facc[ILP];
    uint32_t iacc[ILP];
    const uint32_t a = seed * (threadIdx.x + 1) | 1u, b = (seed ^ 0x9e3779b9u) + blockIdx.x;
    const double fa = 1.0 + 1e-9 * threadIdx.x, fb = 1e-12 * (blockIdx.x + 1);
#pragma unroll
    for (int k = 0; k < ILP; ++k) { acc[k] = k + threadIdx.x; facc[k] = k * 0.5; iacc[k] = k ^ seed; }
    const unsigned long long g0 = globaltimer_ns();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (MODE == 0 || MODE == 4 || MODE == 5 || MODE == 6) asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(acc[k]) : "r"(b));
            if (MODE == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(iacc[k]) : "r"(a), "r"(b));
            if (MODE == 2 || MODE == 4 || MODE == 6) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(facc[k]) : "d"(fa), "d"(fb));
            if (MODE == 3 || MODE == 5 || MODE == 6) asm volatile("add.u32 %0, %0, %1;" : "+r"(iacc[k]) : "r"(a));
            if (MODE == 7) {   // two IADD3 per IMAD.WIDE: the ratio a DFMA-assisted product would need
                asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(acc[k]) : "r"(b));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(iacc[k]) : "r"(a));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(iacc[k]) : "r"(b));
            }
        }
    }
    const long long t1 = clock64();
    const unsigned long long g1 = globaltimer_ns();
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += acc[k] + (uint64_t)__double_as_longlong(facc[k]) + iacc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) { clocks[0] = t1 - t0; clocks[1] = (long long)(g1 - g0); }
}

template <int MODE>
static double run(int blocks, uint64_t* out, long long* clk_dev, double* sm_cycles, double* block_ns) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE><<<blocks, 256>>>(out, 12345u, clk_dev);          // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    long long clk[2] = {0, 0};
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        bench<MODE><<<blocks, 256>>>(out, 12345u + rep, clk_dev);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) { best = ms; cudaMemcpy(clk, clk_dev, sizeof clk, cudaMemcpyDeviceToHost); }
    }
    *sm_cycles = (double)clk[0];
    *block_ns = (double)clk[1];
    return best;
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    const int sms = p.multiProcessorCount;
    const int blocks = sms * 4;                 // 4 x 256 threads = 32 warps per SM = 8 per scheduler
    uint64_t* out;
    long long* clk;
    cudaMalloc(&out, sizeof(uint64_t) * blocks * 256);
    cudaMalloc(&clk, 2 * sizeof(long long));
    spin<<<sms, 128>>>(600000000LL, clk);        // ~0.3 s of load before anything is measured
    cudaDeviceSynchronize();
    const char* names[8] = {"imad_wide", "imad_lo", "dfma", "iadd3", "imad_wide+dfma", "imad_wide+iadd3", "imad_wide+dfma+iadd3", "imad_wide+2xiadd3"};
    const int per_iter[8] = {1, 1, 1, 1, 2, 2, 3, 3};
    double ms[8], cyc[8], ns[8];
    ms[0] = run<0>(blocks, out, clk, &cyc[0], &ns[0]); ms[1] = run<1>(blocks, out, clk, &cyc[1], &ns[1]);
    ms[2] = run<2>(blocks, out, clk, &cyc[2], &ns[2]); ms[3] = run<3>(blocks, out, clk, &cyc[3], &ns[3]);
    ms[4] = run<4>(blocks, out, clk, &cyc[4], &ns[4]); ms[5] = run<5>(blocks, out, clk, &cyc[5], &ns[5]);
    ms[6] = run<6>(blocks, out, clk, &cyc[6], &ns[6]); ms[7] = run<7>(blocks, out, clk, &cyc[7], &ns[7]);
    printf("{\"device\": \"%s\", \"sms\": %d, \"warps_per_scheduler\": 8, \"ilp\": %d, \"results\": {", p.name, sms, ILP);
    for (int m = 0; m < 8; ++m) {
        const double warp_instr = (double)blocks * 8 /*warps*/ * ITERS * ILP * per_iter[m];
        const double gwips = warp_instr / (ms[m] * 1e-3) / 1e9;
        // cycles per warp instruction per scheduler from block 0's clock64() span (8 warps share a scheduler); the
        // clock64 rate itself is calibrated against %globaltimer over the same span
        const double per_sched = cyc[m] / ((double)8 * ITERS * ILP * per_iter[m]);
        printf("%s\"%s\": {\"ms\": %.3f, \"giga_warp_instr_per_s\": %.1f, \"clock64_per_warp_instr_per_scheduler\": %.3f, \"clock64_mhz\": %.0f}",
               m ? ", " : "", names[m], ms[m], gwips, per_sched, cyc[m] / ns[m] * 1e3);
    }
    printf("}}\n");
    return 0;
}
