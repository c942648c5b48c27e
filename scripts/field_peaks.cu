// Throughput of the field arithmetic this engine is made of, as the kernels use it ("measure, don't
// reason about, the second pipe"):
//   mul_cios        Fq Montgomery product (136 IMAD.WIDE)              - the integer-multiply peak of the engine
//   sqr_cios        Fq Montgomery square  (108 IMAD.WIDE)
//   mul_shoup       fixed-operand product (99 IMAD.WIDE + 16 IMAD)
//   mul_cios+dfmaN  the same product stream with N independent DFMA per product in the same thread: if the FP64 pipe
//                   ran beside the integer-multiply pipe, products/s would not move until the issue slots run out
//   dfma68_only     the FP64 stream of the last row alone (68 DFMA per slot): the DFMA peak = slots/s x 68
// Every thread keeps ILP independent chains x_k <- x_k * y; grids are one wave of `w` warps per scheduler.
// Output: giga products per second (whole GPU) and the IMAD.WIDE rate that implies.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I zk-email-verify_b200/csrc -o scripts/build/field_peaks scripts/field_peaks.cu
#include <cstdio>
#include <cstdint>
#include "ff.cuh"
using namespace zke::dev;
namespace zke { namespace dev { ZKE_DEFINE_CONSTANT_UPLOAD(upload_consts) } }

static const int ILP = 2;
static const int ITERS = 2048;

template <int MODE, int NDFMA>
__global__ void __launch_bounds__(128) bench(uint32_t* out, uint32_t seed) {
    Fq x[ILP], y, yq;
    double f[8];
    for (int i = 0; i < 8; ++i) {
        y.v[i] = seed * (i + 3) + threadIdx.x;
        yq.v[i] = seed * (i + 11) ^ blockIdx.x;
        for (int k = 0; k < ILP; ++k) x[k].v[i] = (seed ^ (k + 1)) * (i + 7) + threadIdx.x * 977;
        f[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    }
    y.v[7] &= 0x0fffffffu; yq.v[7] &= 0x0fffffffu;
    for (int k = 0; k < ILP; ++k) x[k].v[7] &= 0x0fffffffu;
    const double fa = 1.0000001, fb = 1e-12 * (blockIdx.x + 1);
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (MODE == 0) x[k] = Fq::mul_cios(x[k], y);
            if (MODE == 1) x[k] = Fq::sqr_cios(x[k]);
            if (MODE == 2) x[k] = Fq::mul_shoup(x[k], y.v, yq.v);
#pragma unroll
            for (int d = 0; d < NDFMA; ++d) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[(k * NDFMA + d) & 7]) : "d"(fa), "d"(fb));
        }
    }
    uint32_t s = 0;
    for (int k = 0; k < ILP; ++k) for (int i = 0; i < 8; ++i) s += x[k].v[i];
    double fs = 0;
    for (int i = 0; i < 8; ++i) fs += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (uint32_t)__double_as_longlong(fs);
}

template <int MODE, int NDFMA>
static void sweep(const char* name, int imad_wide, int sms, uint32_t* out, bool last) {
    printf("\"%s\": {", name);
    double best = 0;
    const int bps_list[5] = {1, 2, 3, 4, 5};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int k = 0; k < 5; ++k) {
        const int blocks = sms * bps_list[k];
        bench<MODE, NDFMA><<<blocks, 128>>>(out, 7u);
        cudaDeviceSynchronize();
        float ms_best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            bench<MODE, NDFMA><<<blocks, 128>>>(out, 7u + rep);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < ms_best) ms_best = ms;
        }
        const double gprod = (double)blocks * 128 * ITERS * ILP / (ms_best * 1e-3) / 1e9;
        if (gprod > best) best = gprod;
        printf("\"warps_per_scheduler_%d\": %.2f, ", bps_list[k], gprod);
    }
    printf("\"peak_giga_products_per_s\": %.2f, \"giga_warp_imad_wide_per_s\": %.1f}%s", best, best * imad_wide / 32.0, last ? "" : ", ");
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    FieldConsts fq = {{0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u},
                      {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u},
                      {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u},
                      0xe4866389u,
                      {0x278302b9u, 0xc3df73e9u, 0x978e3572u, 0x687e956eu, 0x7e7ea7a2u, 0x47afba49u, 0x1ece5fd6u, 0xcf9bb18du}};
    upload_consts(&fq, &fq);
    const int sms = p.multiProcessorCount;
    uint32_t* out;
    cudaMalloc(&out, 4 * (size_t)sms * 8 * 128);
    bench<0, 0><<<sms * 4, 128>>>(out, 1u);     // warm the clocks
    cudaDeviceSynchronize();
    printf("{\"device\": \"%s\", \"sms\": %d, \"ilp\": %d, \"unit\": \"giga Fq products per second, whole GPU (lane products)\", \"results\": {", p.name, sms, ILP);
    sweep<0, 0>("mul_cios", 136, sms, out, false);
    sweep<1, 0>("sqr_cios", 108, sms, out, false);
    sweep<2, 0>("mul_shoup", 99, sms, out, false);
    sweep<0, 8>("mul_cios+dfma8", 136, sms, out, false);
    sweep<0, 32>("mul_cios+dfma32", 136, sms, out, false);
    sweep<0, 68>("mul_cios+dfma68", 136, sms, out, false);
    sweep<3, 68>("dfma68_only", 0, sms, out, true);     // the FP64 stream alone: 68 DFMA per "product" slot
    printf("}}\n");
    return 0;
}
