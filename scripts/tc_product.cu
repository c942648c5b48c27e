// Montgomery products whose REDUCTION runs on the tensor cores (zk-email-verify_b200/csrc/ff_tc.cuh): device-side check
// against the integer product (mul_cios) and throughput of both, in the harness of scripts/field_peaks.cu.
//   mul_cios      136 IMAD.WIDE                               (the integer-multiply peak of round 2)
//   mul_tc        64 IMAD.WIDE + 8 IMMA.16832 per warp + 9    (schoolbook 512-bit product, tensor-core reduction)
//   sqr_tc        36 + 9
//   mul_sos       64 + 72, separated (what the wide product costs without the tensor-core reduction)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I zk-email-verify_b200/csrc -o scripts/build/tc_product scripts/tc_product.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include "ff_tc.cuh"
using namespace zke::dev;
namespace zke { namespace dev { ZKE_DEFINE_CONSTANT_UPLOAD(upload_consts) } }

static const int ILP = 2;
static const int ITERS = 2048;

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__global__ void __launch_bounds__(128) check(const TcTable* tab, uint32_t seed, uint32_t* mismatches, uint32_t* first_bad) {
    __shared__ __align__(16) uint32_t scratch[4][2 * TC_SCRATCH_WORDS];
    TcLane L; L.init(tab, scratch[threadIdx.x >> 5]);
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    Fq a, b;
    for (int i = 0; i < 8; ++i) { a.v[i] = mix(seed + id * 16 + i); b.v[i] = mix(seed * 3 + id * 16 + 8 + i); }
    a.v[7] &= 0x0fffffffu; b.v[7] &= 0x0fffffffu;
    const FieldConsts& C = FQ_C;
    if (id == 0) { for (int i = 0; i < 8; ++i) { a.v[i] = C.mod[i]; b.v[i] = C.mod[i]; } a.v[0] -= 1; b.v[0] -= 1; }   // (p-1)^2
    if (id == 1) a = Fq::zero();
    if (id == 2) { a = Fq::zero(); a.v[0] = 1; b = a; }
    if (id == 3) { for (int i = 0; i < 8; ++i) a.v[i] = 0xffffffffu; b = a; }     // non-canonical operands: same residue, may differ by p
    uint32_t bad = 0;
    for (int round = 0; round < 8; ++round) {
        const Fq want = Fq::mul_cios(a, b), got = FpTc<FqTag>::mul(a, b, L);
        const Fq want2 = Fq::sqr_cios(a), got2 = FpTc<FqTag>::sqr(a, L);
        Fq g3, g4, g5, g6;
        FpTc<FqTag>::mul2(g3, g4, a, b, b, want2, L);
        FpTc<FqTag>::sqr_mul(g5, g6, b, a, want2, L);
        uint32_t Ta[16], Tb[16];
        Fq::mul_wide<8>(Ta, a.v, b.v);
        const Fq g7 = FpTc<FqTag>::redc_mul(Ta, Tb, b, want2, L);        // a*b reduced while b*want2 is formed
        const Fq g8 = FpTc<FqTag>::redc(Tb, L);
        if (id != 3 && (g7 != want || g8 != Fq::mul_cios(b, want2))) { ++bad; if (atomicAdd(first_bad, 1u) == 0) first_bad[1] = 0x80000000u | (id * 8 + round); }
        const bool pair_ok = g3 == want && g4 == Fq::mul_cios(b, want2) && g5 == Fq::sqr_cios(b) && g6 == Fq::mul_cios(a, want2);
        if (id != 3 && (want != got || want2 != got2 || !pair_ok)) { ++bad; if (atomicAdd(first_bad, 1u) == 0) first_bad[1] = id * 8 + round; }
        a = got; b = want2;
    }
    if (bad) atomicAdd(mismatches, bad);
}

__device__ __forceinline__ int r_dummy(int k) { return k & 7; }
template <int MODE>
__global__ void __launch_bounds__(128) bench(const TcTable* tab, uint32_t* out, uint32_t seed) {
    __shared__ __align__(16) uint32_t scratch[4][2 * TC_SCRATCH_WORDS];
    TcLane L; L.init(tab, scratch[threadIdx.x >> 5]);
    Fq x[ILP], y;
    for (int i = 0; i < 8; ++i) {
        y.v[i] = seed * (i + 3) + threadIdx.x;
        for (int k = 0; k < ILP; ++k) x[k].v[i] = (seed ^ (k + 1)) * (i + 7) + threadIdx.x * 977;
    }
    y.v[7] &= 0x0fffffffu;
    for (int k = 0; k < ILP; ++k) x[k].v[7] &= 0x0fffffffu;
    uint32_t Tw[2][16];
    if (MODE == 10) Fq::mul_wide<8>(Tw[0], x[0].v, y.v);
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 10) {       // software-pipelined: the reduction of one chain's product under the rows of the other chain's
            x[0] = FpTc<FqTag>::redc_mul(Tw[0], Tw[1], x[1], y, L);
            x[1] = FpTc<FqTag>::redc_mul(Tw[1], Tw[0], x[0], y, L);
            continue;
        }
        if (MODE == 8) { FpTc<FqTag>::mul2(x[0], x[1], x[0], y, x[1], y, L); continue; }      // two products per call
        if (MODE == 9) { FpTc<FqTag>::sqr_mul(x[0], x[1], x[0], x[1], y, L); continue; }
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (MODE == 0) x[k] = Fq::mul_cios(x[k], y);
            if (MODE == 1) x[k] = FpTc<FqTag>::mul(x[k], y, L);
            if (MODE == 2) x[k] = FpTc<FqTag>::sqr(x[k], L);
            if (MODE == 3) x[k] = Fq::mul_sos_plain(x[k], y);
            if (MODE == 4) x[k] = Fq::sqr_cios(x[k]);
            if (MODE == 5) {        // the tensor-core reduction alone (T_hi = y, T_lo = x)
                uint32_t T[16];
                for (int i = 0; i < 8; ++i) { T[i] = x[k].v[i]; T[8 + i] = y.v[i]; }
                x[k] = FpTc<FqTag>::redc(T, L);
            }
            if (MODE == 6) {        // eight IMMA per "product" and nothing else
                int32_t d[4] = {0, 0, 0, 0};
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                                 : "r"(x[k].v[0]), "r"(x[k].v[1]), "r"(x[k].v[2]), "r"(x[k].v[3]), "r"(L.b[r & 7]), "r"(L.b[(r + 1) & 7]));
                x[k].v[r_dummy(k)] ^= (uint32_t)(d[0] + d[1] + d[2] + d[3]);
            }
            if (MODE == 7) {        // the 512-bit schoolbook product alone (64 IMAD.WIDE + merge), folded back by XOR
                uint32_t T[16];
                Fq::mul_wide<8>(T, x[k].v, y.v);
                for (int i = 0; i < 8; ++i) x[k].v[i] = T[i] ^ T[8 + i];
                x[k].v[7] &= 0x0fffffffu;
            }
        }
    }
    uint32_t s = 0;
    for (int k = 0; k < ILP; ++k) for (int i = 0; i < 8; ++i) s += x[k].v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void sweep(const char* name, const TcTable* tab, int sms, uint32_t* out, bool last) {
    printf("\"%s\": {", name);
    double best = 0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int bps = 1; bps <= 6; ++bps) {
        const int blocks = sms * bps;
        bench<MODE><<<blocks, 128>>>(tab, out, 7u);
        cudaDeviceSynchronize();
        float ms_best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            bench<MODE><<<blocks, 128>>>(tab, out, 7u + rep);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < ms_best) ms_best = ms;
        }
        const double gprod = (double)blocks * 128 * ITERS * ILP / (ms_best * 1e-3) / 1e9;
        if (gprod > best) best = gprod;
        printf("\"warps_per_scheduler_%d\": %.2f, ", bps, gprod);
    }
    printf("\"peak_giga_products_per_s\": %.2f}%s", best, last ? "" : ", ");
}

template <int MODE>
static void single(const TcTable* tab, int sms, int bps, uint32_t* out) {
    for (int rep = 0; rep < 3; ++rep) bench<MODE><<<sms * bps, 128>>>(tab, out, 7u + rep);
    cudaDeviceSynchronize();
}

int main(int argc, char** argv) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    FieldConsts fq = {{0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u},
                      {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u},
                      {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u},
                      0xe4866389u,
                      {0x278302b9u, 0xc3df73e9u, 0x978e3572u, 0x687e956eu, 0x7e7ea7a2u, 0x47afba49u, 0x1ece5fd6u, 0xcf9bb18du}};
    upload_consts(&fq, &fq);
    TcTable host_tab;
    tc_build_table(fq.mod, &host_tab);
    TcTable* tab;
    cudaMalloc(&tab, sizeof(TcTable));
    cudaMemcpy(tab, &host_tab, sizeof(TcTable), cudaMemcpyHostToDevice);
    const int sms = p.multiProcessorCount;
    uint32_t *out, *mism;
    cudaMalloc(&out, 4 * (size_t)sms * 8 * 128);
    cudaMalloc(&mism, 16);
    cudaMemset(mism, 0, 16);
    if (argc == 3) {      // one configuration only (for ncu): tc_product <mode> <blocks per SM>
        const int mode = atoi(argv[1]), bps = atoi(argv[2]);
        if (mode == 0) single<0>(tab, sms, bps, out);
        if (mode == 1) single<1>(tab, sms, bps, out);
        if (mode == 2) single<2>(tab, sms, bps, out);
        if (mode == 5) single<5>(tab, sms, bps, out);
        if (mode == 8) single<8>(tab, sms, bps, out);
        if (mode == 10) single<10>(tab, sms, bps, out);
        return 0;
    }
    check<<<1024, 128>>>(tab, 12345u, mism, mism + 1);
    uint32_t h[4] = {0, 0, 0, 0};
    cudaMemcpy(h, mism, 16, cudaMemcpyDeviceToHost);
    const cudaError_t err = cudaDeviceSynchronize();
    printf("{\"device\": \"%s\", \"sms\": %d, \"ilp\": %d, \"mu\": \"0x%08x\", \"check\": {\"products\": %d, \"mismatches\": %u, \"first_bad\": %u, \"cuda\": \"%s\"}, "
           "\"unit\": \"giga Fq products per second, whole GPU (lane products)\", \"results\": {",
           p.name, sms, ILP, host_tab.mu, 1024 * 128 * 16, h[0], h[2], cudaGetErrorString(err));
    bench<0><<<sms * 4, 128>>>(tab, out, 1u);     // warm the clocks
    cudaDeviceSynchronize();
    sweep<0>("mul_cios", tab, sms, out, false);
    sweep<1>("mul_tc", tab, sms, out, false);
    sweep<4>("sqr_cios", tab, sms, out, false);
    sweep<2>("sqr_tc", tab, sms, out, false);
    sweep<3>("mul_sos_plain", tab, sms, out, false);
    sweep<5>("redc_tc_only", tab, sms, out, false);
    sweep<6>("imma8_only", tab, sms, out, false);
    sweep<7>("mul_wide_only", tab, sms, out, false);
    sweep<8>("mul_tc_pair", tab, sms, out, false);
    sweep<9>("sqr_mul_tc_pair", tab, sms, out, false);
    sweep<10>("mul_tc_pipelined", tab, sms, out, true);
    printf("}}\n");
    return h[0] ? 2 : 0;
}
