// Radix-2 NTT over BN254 Fr for the Groth16 quotient ("6 NTTs of size N" - snarkjs groth16_prove steps 3-4,
// SURVEY 3.2 / A.7).  Each global pass fuses up to 8 butterfly stages in shared memory (tile of 1024 elements,
// limb-planar so a warp's accesses are conflict-free), so a 2^22 transform is 3 read+write sweeps over HBM
// (algorithmic bytes per transform: 2 * 32 * N).  The inverse transform is decimation-in-frequency
// (natural in, bit-reversed out) and the forward transform decimation-in-time (bit-reversed in, natural out), so
// no bit-reversal permutation is ever materialised; the coset shift g^j / N is fused into the store of the last
// inverse pass.  Twiddles are 32-byte vector loads from a table of N/2 powers that stays L2-resident (64 MB at 2^22).
#include "device_engine.cuh"
#include "ntt.cuh"

namespace zke {
namespace dev {

static const int MAX_TILE_LOG = 10;   // 1024 elements = 32 KB of shared memory per CTA, 512 threads

struct Planar {
    uint32_t* sm;
    int tile;
    __device__ __forceinline__ Fr get(int e) const { Fr r; for (int i = 0; i < 8; ++i) r.v[i] = sm[i * tile + e]; return r; }
    __device__ __forceinline__ void put(int e, const Fr& x) const { for (int i = 0; i < 8; ++i) sm[i * tile + e] = x.v[i]; }
};

// One fused pass over stage bits [s_lo, s_hi].  DIF: stages descend and multiply after the subtraction;
// DIT: stages ascend and multiply before the butterfly.  `scale`, if non-null, multiplies element idx on store.
template <bool DIF>
__global__ void __launch_bounds__(1 << (MAX_TILE_LOG - 1))
ntt_pass_kernel(uint8_t* __restrict__ data, const uint8_t* __restrict__ tw, const uint8_t* __restrict__ scale,
                int log_n, int s_lo, int s_hi, int tile_log) {
    extern __shared__ uint32_t smem[];
    const int TILE = 1 << tile_log;
    const int NTT_THREADS = TILE / 2;
    Planar S{smem, TILE};
    const int k = s_hi - s_lo + 1;
    const int J = 1 << k;
    const int L = TILE / J;
    const uint32_t bid = blockIdx.x;
    const bool strided = s_lo > 0;
    uint32_t hi = 0, lo_base = 0;
    if (strided) {
        const uint32_t lo_blocks = (1u << s_lo) / L;
        lo_base = (bid % lo_blocks) * L;
        hi = bid / lo_blocks;
    }
    auto gidx = [&](int j, int l) -> uint32_t {
        return strided ? ((hi << (s_hi + 1)) | ((uint32_t)j << s_lo) | (lo_base + l))
                       : (((bid * L + l) << k) | (uint32_t)j);
    };
    auto sidx = [&](int j, int l) -> int { return strided ? (j * L + l) : (l * J + j); };

    for (int e = threadIdx.x; e < TILE; e += NTT_THREADS) {
        int j, l;
        if (strided) { l = e % L; j = e / L; } else { j = e % J; l = e / J; }
        S.put(e, Fr::load(data + 32ull * gidx(j, l)));
    }
    __syncthreads();

    const int t = threadIdx.x;
    int l, jj;
    if (strided) { l = t % L; jj = t / L; } else { jj = t % (J / 2); l = t / (J / 2); }
    for (int step = 0; step < k; ++step) {
        const int st = DIF ? (s_hi - step) : (s_lo + step);   // global stage: butterfly distance 2^st
        const int h = 1 << (st - s_lo);                        // distance in j units
        const int j0 = ((jj / h) * 2 * h) + (jj % h);
        const int e0 = sidx(j0, l), e1 = sidx(j0 + h, l);
        const uint32_t idx0 = gidx(j0, l);
        const uint32_t tw_i = (idx0 & ((1u << st) - 1)) << (log_n - 1 - st);
        const Fr w = Fr::load(tw + 32ull * tw_i);
        Fr u = S.get(e0), v = S.get(e1);
        if (DIF) {
            S.put(e0, u + v);
            S.put(e1, (u - v) * w);
        } else {
            Fr x = v * w;
            S.put(e0, u + x);
            S.put(e1, u - x);
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < TILE; e += NTT_THREADS) {
        int j, l2;
        if (strided) { l2 = e % L; j = e / L; } else { j = e % J; l2 = e / J; }
        const uint32_t g = gidx(j, l2);
        Fr x = S.get(e);
        if (scale) x = x * Fr::load(scale + 32ull * g);
        x.store(data + 32ull * g);
    }
}

static void plan(int log_n, int* lo, int* hi, int* n_pass) {
    // split log_n stage bits into passes of at most 8 stages, as evenly as possible, top bits first
    int passes = (log_n + 7) / 8;
    int base = log_n / passes, extra = log_n % passes;
    int top = log_n - 1;
    for (int p = 0; p < passes; ++p) {
        int k = base + (p < extra ? 1 : 0);
        hi[p] = top;
        lo[p] = top - k + 1;
        top -= k;
    }
    *n_pass = passes;
}

void launch_intt_dif(uint8_t* data, const NttTables& T, const uint8_t* scale_bitrev, cudaStream_t st) {
    int lo[4], hi[4], np;
    plan(T.log_n, lo, hi, &np);
    ZKE_COUNT_LAUNCH(np);
    const int tile_log = T.log_n < MAX_TILE_LOG ? T.log_n : MAX_TILE_LOG;
    const uint32_t blocks = 1u << (T.log_n - tile_log);
    for (int p = 0; p < np; ++p)
        ntt_pass_kernel<true><<<blocks, 1 << (tile_log - 1), 32u << tile_log, st>>>(
            data, T.tw_inv, p == np - 1 ? scale_bitrev : nullptr, T.log_n, lo[p], hi[p], tile_log);
}

void launch_ntt_dit(uint8_t* data, const NttTables& T, cudaStream_t st) {
    int lo[4], hi[4], np;
    plan(T.log_n, lo, hi, &np);
    ZKE_COUNT_LAUNCH(np);
    const int tile_log = T.log_n < MAX_TILE_LOG ? T.log_n : MAX_TILE_LOG;
    const uint32_t blocks = 1u << (T.log_n - tile_log);
    for (int p = np - 1; p >= 0; --p)
        ntt_pass_kernel<false><<<blocks, 1 << (tile_log - 1), 32u << tile_log, st>>>(
            data, T.tw_fwd, nullptr, T.log_n, lo[p], hi[p], tile_log);
}

// c = a o b   (Montgomery in/out)
__global__ void hadamard_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ c, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    (Fr::load(a + 32ull * i) * Fr::load(b + 32ull * i)).store(c + 32ull * i);
}
void launch_hadamard(const uint8_t* a, const uint8_t* b, uint8_t* c, uint32_t n, cudaStream_t st) {
    hadamard_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, b, c, n);
    ZKE_COUNT_LAUNCH(1);
}

// d = a*b - c on the coset, converted to standard form (the scalars of the H multi-exponentiation)
__global__ void quotient_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, const uint8_t* __restrict__ c,
                                uint8_t* __restrict__ d, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr x = Fr::load(a + 32ull * i) * Fr::load(b + 32ull * i) - Fr::load(c + 32ull * i);
    x.from_mont().store(d + 32ull * i);
}
void launch_quotient(const uint8_t* a, const uint8_t* b, const uint8_t* c, uint8_t* d, uint32_t n, cudaStream_t st) {
    quotient_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, b, c, d, n);
    ZKE_COUNT_LAUNCH(1);
}

}  // namespace dev
}  // namespace zke
