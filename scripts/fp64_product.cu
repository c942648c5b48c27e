// Can the FP64 pipe carry whole field products beside the integer pipe?  (DESIGN.md section 5 / "What comes next" item 0)
//   An Fq Montgomery product in a redundant FP64 representation: 12 limbs of 22 bits held in doubles (R = 2^264).  Limb
//   products are < 2^44 and a column of twelve of them < 2^48, so every DFMA accumulation is exact; reduction limb by limb
//   (m = (c_i * (-p^-1 mod 2^22)) mod 2^22, c += m * p, carry c_i / 2^22) in the same representation; floors by a
//   round-down DFMA against 2^52.  ~470 FP64 instructions per product against 136 IMAD.WIDE + ~100 ALU for the integer one.
// The program (1) checks the FP64 product against the integer product on 2^20 random operand pairs on the device,
// (2) measures products/s for: the integer product alone, the FP64 product alone, and kernels in which every 2nd / 3rd / 4th
// warp runs the FP64 product while the others run the integer one (both kinds do whole, independent product chains).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I zk-email-verify_b200/csrc -o scripts/build/fp64_product scripts/fp64_product.cu
#include <cstdio>
#include <cstdint>
#include "ff.cuh"
using namespace zke::dev;
namespace zke { namespace dev { ZKE_DEFINE_CONSTANT_UPLOAD(upload_consts) } }

struct F64 { double l[12]; };
__constant__ double P64[12] = {3996999.0, 3169121.0, 1294856.0, 1864355.0, 2789736.0, 3563013.0, 1578373.0, 1142176.0, 2734160.0, 312960.0, 321326.0, 3097.0};
static __device__ __forceinline__ double floor22(double v) {      // floor(v / 2^22) for |v| < 2^51
    const double MAGIC = 6755399441055744.0;                       // 1.5 * 2^52: keeps the sum in one binade for negative v too
    return __fma_rd(v, 1.0 / 4194304.0, MAGIC) - MAGIC;
}

__device__ F64 to_f64(const Fq& x) {
    F64 r;
    for (int i = 0; i < 12; ++i) {
        const int bit = 22 * i, w = bit >> 5, s = bit & 31;
        unsigned long long two = w < 8 ? x.v[w] : 0;
        if (w + 1 < 8) two |= (unsigned long long)x.v[w + 1] << 32;
        r.l[i] = (double)(uint32_t)((two >> s) & 0x3fffffu);
    }
    return r;
}
__device__ Fq to_int(const F64& a) {    // limbs normalised to [0, 2^22)
    Fq r = Fq::zero();
    for (int i = 0; i < 12; ++i) {
        const unsigned long long v = (unsigned long long)a.l[i];
        const int bit = 22 * i, w = bit >> 5, s = bit & 31;
        if (w < 8) r.v[w] |= (uint32_t)(v << s);
        if (w + 1 < 8 && s > 10) r.v[w + 1] |= (uint32_t)(v >> (32 - s));
    }
    return r;
}

__device__ __forceinline__ F64 mont_mul64(const F64& a, const F64& b) {
    const double T22 = 4194304.0, PINV0 = 418697.0;
    double c[24];
#pragma unroll
    for (int k = 0; k < 24; ++k) c[k] = 0.0;
#pragma unroll
    for (int i = 0; i < 12; ++i)
#pragma unroll
        for (int j = 0; j < 12; ++j) c[i + j] = fma(a.l[i], b.l[j], c[i + j]);
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const double lo = fma(-T22, floor22(c[i]), c[i]);          // c_i mod 2^22
        const double t = lo * PINV0;                                // < 2^44, exact
        const double m = fma(-T22, floor22(t), t);                  // m_i
#pragma unroll
        for (int j = 0; j < 12; ++j) c[i + j] = fma(m, P64[j], c[i + j]);
        c[i + 1] = fma(c[i], 1.0 / 4194304.0, c[i + 1]);            // c_i is now a multiple of 2^22
    }
    F64 r, d;
    double carry = 0.0, borrow = 0.0;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const double v = c[12 + k] + carry;
        carry = floor22(v);
        r.l[k] = fma(-T22, carry, v);
        const double u = r.l[k] - P64[k] + borrow;                  // r - p, limb by limb
        borrow = floor22(u);                                        // -1 or 0
        d.l[k] = fma(-T22, borrow, u);
    }
    if (borrow == 0.0) r = d;                                       // r >= p
    return r;
}

__global__ void check_kernel(uint32_t n, uint32_t seed, unsigned long long* bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fq a, b;
    uint32_t s = seed ^ (i * 2654435761u);
    for (int k = 0; k < 8; ++k) { s = s * 1664525u + 1013904223u; a.v[k] = s; s = s * 1664525u + 1013904223u; b.v[k] = s; }
    a.v[7] &= 0x0fffffffu; b.v[7] &= 0x0fffffffu;                   // < 2^252 < p
    if (i == 0) { for (int k = 0; k < 8; ++k) { a.v[k] = FQ_C.mod[k]; b.v[k] = FQ_C.mod[k]; } a.v[0] -= 1; b.v[0] -= 1; }   // (p-1)^2
    Fq want = Fq::mul_cios(a, b);                                    // a b 2^-256
    Fq got = to_int(mont_mul64(to_f64(a), to_f64(b)));               // a b 2^-264
    for (int k = 0; k < 8; ++k) got = got.dbl();                     // * 2^8
    if (got != want) atomicAdd(bad, 1ull);
}

static const int ITERS = 2048;
// WARP_MOD = 0: every warp the integer product; 1: every warp the FP64 product; m >= 2: warps with (warp % m) == 0 run the
// FP64 product (FP64_ITERS iterations), the others the integer one (ITERS iterations)
template <int WARP_MOD, int FP64_ITERS>
__global__ void __launch_bounds__(128) bench(uint32_t* out, uint32_t seed) {
    const int warp = threadIdx.x >> 5;
    const bool fp = WARP_MOD == 1 || (WARP_MOD >= 2 && (warp % WARP_MOD) == 0);
    Fq x[2], y;
    for (int i = 0; i < 8; ++i) { y.v[i] = seed * (i + 3) + threadIdx.x; x[0].v[i] = (seed ^ 1) * (i + 7) + threadIdx.x * 977; x[1].v[i] = (seed ^ 2) * (i + 5) + blockIdx.x; }
    y.v[7] &= 0x0fffffffu; x[0].v[7] &= 0x0fffffffu; x[1].v[7] &= 0x0fffffffu;
    uint32_t s = 0;
    if (fp) {
        F64 fx0 = to_f64(x[0]), fx1 = to_f64(x[1]), fy = to_f64(y);
#pragma unroll 1
        for (int it = 0; it < FP64_ITERS; ++it) { fx0 = mont_mul64(fx0, fy); fx1 = mont_mul64(fx1, fy); }
        for (int i = 0; i < 12; ++i) s += (uint32_t)fx0.l[i] + (uint32_t)fx1.l[i];
    } else {
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) { x[0] = Fq::mul_cios(x[0], y); x[1] = Fq::mul_cios(x[1], y); }
        for (int i = 0; i < 8; ++i) s += x[0].v[i] + x[1].v[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int WARP_MOD, int FP64_ITERS>
static void run(const char* name, int sms, uint32_t* out, bool last) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0, best_int = 0, best_fp = 0;
    int best_bps = 0;
    for (int bps = 1; bps <= 6; ++bps) {                 // blocks of 4 warps per SM = warps per scheduler
        const int blocks = sms * bps;
        bench<WARP_MOD, FP64_ITERS><<<blocks, 128>>>(out, 7u);
        cudaDeviceSynchronize();
        float ms_best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            bench<WARP_MOD, FP64_ITERS><<<blocks, 128>>>(out, 9u + rep);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < ms_best) ms_best = ms;
        }
        const double warps = (double)blocks * 4;
        const double fp_warps = WARP_MOD == 0 ? 0 : (WARP_MOD == 1 ? warps : (double)blocks * ((4 + WARP_MOD - 1) / WARP_MOD));
        const double int_warps = warps - fp_warps;
        const double gi = int_warps * 32 * ITERS * 2 / (ms_best * 1e-3) / 1e9, gf = fp_warps * 32 * (double)FP64_ITERS * 2 / (ms_best * 1e-3) / 1e9;
        if (gi + gf > best) { best = gi + gf; best_int = gi; best_fp = gf; best_bps = bps; }
    }
    printf("\"%s\": {\"giga_products_per_s\": %.2f, \"integer_part\": %.2f, \"fp64_part\": %.2f, \"warps_per_scheduler\": %d}%s", name, best, best_int, best_fp, best_bps, last ? "" : ", ");
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
    unsigned long long* bad;
    cudaMalloc(&out, 4 * (size_t)sms * 8 * 128);
    cudaMalloc(&bad, 8);
    cudaMemset(bad, 0, 8);
    const uint32_t n_check = 1u << 20;
    check_kernel<<<n_check / 128, 128>>>(n_check, 12345u, bad);
    unsigned long long nbad = 0;
    cudaMemcpy(&nbad, bad, 8, cudaMemcpyDeviceToHost);
    printf("{\"device\": \"%s\", \"fp64_product_checked\": %u, \"mismatches_vs_integer_product\": %llu, \"unit\": \"giga Fq products per second, whole GPU\", \"results\": {", p.name, n_check, nbad);
    run<0, 1>("integer_only", sms, out, false);
    run<1, ITERS>("fp64_only", sms, out, false);
    run<2, ITERS * 5 / 8>("every_2nd_warp_fp64", sms, out, false);
    run<4, ITERS * 5 / 8>("every_4th_warp_fp64", sms, out, false);
    run<4, ITERS>("every_4th_warp_fp64_equal_iters", sms, out, true);
    printf("}}\n");
    return 0;
}
