// Radix-2 NTT over BN254 Fr for the Groth16 quotient ("6 NTTs of size N" - snarkjs groth16_prove steps 3-4,
// SURVEY 3.2 / A.7).  Each global pass fuses up to 8 butterfly stages in shared memory (tile of 1024 elements,
// limb-planar so a warp's accesses are conflict-free), so a 2^22 transform is 3 read+write sweeps over HBM
// (algorithmic bytes per transform: 2 * 32 * N).  The inverse transform is decimation-in-frequency
// (natural in, bit-reversed out) and the forward transform decimation-in-time (bit-reversed in, natural out), so
// no bit-reversal permutation is ever materialised; the coset shift g^j / N is fused into the store of the last
// inverse pass.  Twiddles (and the coset factors) are constants; the tables hold them either in Montgomery form
// (32 bytes, Montgomery product) or as fixed-operand pairs {w in standard form, floor(w 2^256 / r)} (64 bytes,
// Fp::mul_shoup of ff.cuh: 99 IMAD.WIDE + 16 IMAD instead of 136 IMAD.WIDE) - NttTables::shoup, chosen when the
// context is opened (ZKE_NTT_SHOUP); the data stays in Montgomery form either way.
#include "device_engine.cuh"
#include "ntt.cuh"
#include <cstdlib>
#include <type_traits>

namespace zke {
namespace dev {

// A constant factor (twiddle, coset scale) as the tables hold it: SHOUP: the fixed-operand pair {w standard form,
// floor(w 2^256 / r)} (64 bytes, Fp::mul_shoup); otherwise w in Montgomery form (32 bytes, Montgomery product).
template <bool SHOUP>
struct Twid {
    Fr w, wq;
    static constexpr uint64_t STRIDE = SHOUP ? 64 : 32;
    __device__ __forceinline__ void load(const uint8_t* base, uint64_t idx) {
        w = Fr::load(base + STRIDE * idx);
        if (SHOUP) wq = Fr::load(base + STRIDE * idx + 32);
    }
    __device__ __forceinline__ Fr mul(const Fr& x) const { return SHOUP ? Fr::mul_shoup(x, w.v, wq.v) : x * w; }
};

static const int MAX_TILE_LOG = 10;   // 1024 elements = 32 KB of shared memory per CTA, 512 threads

struct Planar {
    uint32_t* sm;
    int tile;
    __device__ __forceinline__ Fr get(int e) const { Fr r; for (int i = 0; i < 8; ++i) r.v[i] = sm[i * tile + e]; return r; }
    __device__ __forceinline__ void put(int e, const Fr& x) const { for (int i = 0; i < 8; ++i) sm[i * tile + e] = x.v[i]; }
};

// One fused pass over stage bits [s_lo, s_hi].  DIF: stages descend and multiply after the subtraction;
// DIT: stages ascend and multiply before the butterfly.  `scale`, if non-null, multiplies element idx on store.
template <bool DIF, bool SHOUP>
__global__ void __launch_bounds__(1 << (MAX_TILE_LOG - 1))
ntt_pass_kernel(uint8_t* __restrict__ data, const uint8_t* __restrict__ tw, const uint8_t* __restrict__ scale,
                int log_n, int s_lo, int s_hi, int tile_log) {   // log_n: size of the transform the twiddle table belongs to
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
        Twid<SHOUP> w;
        w.load(tw, tw_i);
        Fr u = S.get(e0), v = S.get(e1);
        if (DIF) {
            S.put(e0, u + v);
            S.put(e1, w.mul(u - v));
        } else {
            Fr x = w.mul(v);
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
        if (scale) { Twid<SHOUP> sc; sc.load(scale, g); x = sc.mul(x); }
        x.store(data + 32ull * g);
    }
}

// ---- fast path: tile of 1024 elements, stage count K and addressing mode known at compile time ------------------
// Shared memory holds the tile as two planes of uint4 (limbs 0-3 / limbs 4-7): a quarter-warp reading consecutive
// elements touches 128 contiguous bytes per plane, so the 128-bit shared loads are conflict-free.  All index
// arithmetic is shifts and masks; the twiddle of the next stage is fetched before the current stage's product so
// the L2 latency of the table lookup overlaps the multiplication.
template <bool DIF, int K, bool STRIDED, bool SHOUP>
__global__ void __launch_bounds__(512, 2)
ntt_pass_fast_kernel(uint8_t* __restrict__ data, const uint8_t* __restrict__ tw, const uint8_t* __restrict__ scale,
                     int log_n, int s_lo_arg) {
    const int s_lo = STRIDED ? s_lo_arg : 0;   // the contiguous pass always starts at stage 0 (compile-time stages)
    constexpr int TILE_LOG = 10, TILE = 1 << TILE_LOG, J = 1 << K, L_LOG = TILE_LOG - K, L = 1 << L_LOG;
    __shared__ uint4 plane0[TILE], plane1[TILE];
    const int s_hi = s_lo + K - 1;
    const uint32_t bid = blockIdx.x;
    uint32_t hi = 0, lo_base = 0;
    if constexpr (STRIDED) {
        const uint32_t lo_blocks_log = s_lo - L_LOG;
        lo_base = (bid & ((1u << lo_blocks_log) - 1)) << L_LOG;
        hi = bid >> lo_blocks_log;
    }
    auto gidx = [&](uint32_t j, uint32_t l) -> uint32_t {
        return STRIDED ? ((hi << (s_hi + 1)) | (j << s_lo) | (lo_base + l)) : ((((bid << L_LOG) + l) << K) | j);
    };
    auto sidx = [&](uint32_t j, uint32_t l) -> uint32_t { return STRIDED ? ((j << L_LOG) + l) : ((l << K) + j); };
    auto sget = [&](uint32_t e) -> Fr {
        const uint4 a = plane0[e], b = plane1[e];
        Fr r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
        return r;
    };
    auto sput = [&](uint32_t e, const Fr& x) {
        plane0[e] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
        plane1[e] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
    };
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const uint32_t e = threadIdx.x + r * 512;
        const uint32_t j = STRIDED ? (e >> L_LOG) : (e & (J - 1)), l = STRIDED ? (e & (L - 1)) : (e >> K);
        const uint4* src = reinterpret_cast<const uint4*>(data + 32ull * gidx(j, l));
        plane0[e] = src[0];
        plane1[e] = src[1];
    }
    const uint32_t t = threadIdx.x;
    const uint32_t l = STRIDED ? (t & (L - 1)) : (t >> (K - 1));
    const uint32_t jj = STRIDED ? (t >> L_LOG) : (t & (J / 2 - 1));
    auto stage_of = [&](int step) { return DIF ? (s_hi - step) : (s_lo + step); };
    auto j0_of = [&](int st) -> uint32_t {
        const uint32_t hb = st - s_lo;                     // log2 of the butterfly distance in j units
        return ((jj >> hb) << (hb + 1)) | (jj & ((1u << hb) - 1));
    };
    auto tw_idx = [&](int st, uint32_t j0) -> uint64_t {
        const uint32_t idx0 = gidx(j0, l);
        return (uint64_t)(idx0 & ((1u << st) - 1)) << (log_n - 1 - st);
    };
    Twid<SHOUP> w_next;
    w_next.load(tw, tw_idx(stage_of(0), j0_of(stage_of(0))));
    __syncthreads();
#pragma unroll
    for (int step = 0; step < K; ++step) {
        const int st = stage_of(step);
        const uint32_t j0 = j0_of(st);
        const uint32_t e0 = sidx(j0, l), e1 = sidx(j0 + (1u << (st - s_lo)), l);
        const Twid<SHOUP> w = w_next;
        if (step + 1 < K) w_next.load(tw, tw_idx(stage_of(step + 1), j0_of(stage_of(step + 1))));
        const Fr u = sget(e0), v = sget(e1);
        const bool trivial = !STRIDED && st == 0;   // stage 0: every twiddle is omega^0 = 1, no product
        if (DIF) {
            sput(e0, u + v);
            sput(e1, trivial ? (u - v) : w.mul(u - v));
        } else {
            const Fr x = trivial ? v : w.mul(v);
            sput(e0, u + x);
            sput(e1, u - x);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const uint32_t e = threadIdx.x + r * 512;
        const uint32_t j = STRIDED ? (e >> L_LOG) : (e & (J - 1)), l2 = STRIDED ? (e & (L - 1)) : (e >> K);
        const uint32_t g = gidx(j, l2);
        if (scale) {
            Twid<SHOUP> sc;
            sc.load(scale, g);
            sc.mul(sget(e)).store(data + 32ull * g);
        } else {
            uint4* dst = reinterpret_cast<uint4*>(data + 32ull * g);
            dst[0] = plane0[e];
            dst[1] = plane1[e];
        }
    }
}

// ---- register-blocked variant: 8 elements per thread, up to three stages between shared-memory exchanges -----------
// The kernel above synchronises the CTA after every stage and gives a thread ONE product per stage: it runs at ~65 % of
// the multiplier pipe because each stage is a chain load -> product -> store -> barrier.  Here a thread holds 2^R
// elements (R <= 3) in registers and does R stages on them - 4 independent butterflies per stage - before the tile is
// exchanged through shared memory again: 3 barriers instead of 8 per pass, a third of the shared-memory traffic, and
// enough independent products per thread for the fixed-operand product (whose two phases are dependent) to pay off.
// Shared memory is padded by one element per eight (index e -> e + (e >> 3)) so that the strides 2, 4, 8 of the
// register blocks stay conflict-free for 128-bit accesses.
template <bool DIF, int K, bool STRIDED, bool SHOUP>
__global__ void __launch_bounds__(128, 3)
ntt_pass_r8_kernel(uint8_t* __restrict__ data, const uint8_t* __restrict__ tw, const uint8_t* __restrict__ scale,
                   int log_n, int s_lo_arg) {
    const int s_lo = STRIDED ? s_lo_arg : 0;
    constexpr int TILE_LOG = 10, TILE = 1 << TILE_LOG, J = 1 << K, L_LOG = TILE_LOG - K, L = 1 << L_LOG;
    constexpr int SM = TILE + TILE / 8;
    __shared__ uint4 plane0[SM], plane1[SM];
    const int s_hi = s_lo + K - 1;
    const uint32_t bid = blockIdx.x;
    uint32_t hi = 0, lo_base = 0;
    if constexpr (STRIDED) {
        const uint32_t lo_blocks_log = s_lo - L_LOG;
        lo_base = (bid & ((1u << lo_blocks_log) - 1)) << L_LOG;
        hi = bid >> lo_blocks_log;
    }
    auto gidx = [&](uint32_t j, uint32_t l) -> uint32_t {
        return STRIDED ? ((hi << (s_hi + 1)) | (j << s_lo) | (lo_base + l)) : ((((bid << L_LOG) + l) << K) | j);
    };
    auto pad = [](uint32_t e) -> uint32_t { return e + (e >> 3); };
    auto sget = [&](uint32_t e) -> Fr {
        const uint4 a = plane0[pad(e)], b = plane1[pad(e)];
        Fr r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
        return r;
    };
    auto sput = [&](uint32_t e, const Fr& x) {
        plane0[pad(e)] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
        plane1[pad(e)] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
    };
    // tile element e <-> (j, l): the same mapping as the global load of ntt_pass_fast_kernel (coalesced 32-byte rows)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint32_t e = threadIdx.x + r * 128;
        const uint32_t j = STRIDED ? (e >> L_LOG) : (e & (J - 1)), l = STRIDED ? (e & (L - 1)) : (e >> K);
        const uint4* src = reinterpret_cast<const uint4*>(data + 32ull * gidx(j, l));
        plane0[pad(e)] = src[0];
        plane1[pad(e)] = src[1];
    }
    __syncthreads();
    // one round: the R stages of local stage bits [B, B + R)
    auto round = [&](auto b_tag, auto r_tag) {
        constexpr int B = decltype(b_tag)::value, R = decltype(r_tag)::value, NR = 1 << R;
        constexpr int BP = STRIDED ? B + L_LOG : B;        // position of local stage bit B in the tile index
        constexpr int ITEMS = TILE >> R;
#pragma unroll 1
        for (uint32_t u = threadIdx.x; u < ITEMS; u += 128) {
            const uint32_t base = (u & ((1u << BP) - 1)) | ((u >> BP) << (BP + R));
            const uint32_t j = STRIDED ? (base >> L_LOG) : (base & (J - 1)), l = STRIDED ? (base & (L - 1)) : (base >> K);
            Fr x[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) x[r] = sget(base + ((uint32_t)r << BP));
#pragma unroll
            for (int qq = 0; qq < R; ++qq) {
                const int q = DIF ? (R - 1 - qq) : qq;
                const int st = s_lo + B + q;
                const bool trivial = !STRIDED && B + q == 0;     // stage 0: every twiddle is 1
                Twid<SHOUP> w[NR / 2];
                if (!trivial) {
#pragma unroll
                    for (int p = 0; p < NR / 2; ++p) {
                        const int r0 = ((p >> q) << (q + 1)) | (p & ((1 << q) - 1));
                        const uint32_t idx0 = gidx(j | ((uint32_t)r0 << B), l);
                        w[p].load(tw, (uint64_t)(idx0 & ((1u << st) - 1)) << (log_n - 1 - st));
                    }
                }
#pragma unroll
                for (int p = 0; p < NR / 2; ++p) {
                    const int r0 = ((p >> q) << (q + 1)) | (p & ((1 << q) - 1)), r1 = r0 | (1 << q);
                    const Fr a = x[r0], b = x[r1];
                    if (DIF) {
                        x[r0] = a + b;
                        x[r1] = trivial ? (a - b) : w[p].mul(a - b);
                    } else {
                        const Fr y = trivial ? b : w[p].mul(b);
                        x[r0] = a + y;
                        x[r1] = a - y;
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < NR; ++r) sput(base + ((uint32_t)r << BP), x[r]);
        }
        __syncthreads();
    };
    using std::integral_constant;
    constexpr int R0 = K >= 3 ? 3 : K, K1 = K - R0, R1 = K1 >= 3 ? 3 : K1, R2 = K1 - R1;      // K = 8: 3 + 3 + 2, 7: 3 + 3 + 1, ...
    if (DIF) {      // stages descend: the top bits first
        round(integral_constant<int, K - R0>{}, integral_constant<int, R0>{});
        if constexpr (R1 > 0) round(integral_constant<int, K - R0 - R1>{}, integral_constant<int, R1>{});
        if constexpr (R2 > 0) round(integral_constant<int, 0>{}, integral_constant<int, R2>{});
    } else {        // stages ascend
        round(integral_constant<int, 0>{}, integral_constant<int, R0>{});
        if constexpr (R1 > 0) round(integral_constant<int, R0>{}, integral_constant<int, R1>{});
        if constexpr (R2 > 0) round(integral_constant<int, R0 + R1>{}, integral_constant<int, R2>{});
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint32_t e = threadIdx.x + r * 128;
        const uint32_t j = STRIDED ? (e >> L_LOG) : (e & (J - 1)), l2 = STRIDED ? (e & (L - 1)) : (e >> K);
        const uint32_t g = gidx(j, l2);
        if (scale) {
            Twid<SHOUP> sc;
            sc.load(scale, g);
            sc.mul(sget(e)).store(data + 32ull * g);
        } else {
            uint4* dst = reinterpret_cast<uint4*>(data + 32ull * g);
            dst[0] = plane0[pad(e)];
            dst[1] = plane1[pad(e)];
        }
    }
}

template <bool DIF, bool SHOUP>
static bool launch_fast(uint8_t* data, const uint8_t* tw, const uint8_t* scale, int log_n, int s_lo, int s_hi, cudaStream_t st, int log_tw) {
    if (log_n < 10) return false;
    const int k = s_hi - s_lo + 1;
    const uint32_t blocks = 1u << (log_n - 10);
    const bool strided = s_lo > 0;
    if (strided && s_lo < 10 - k) return false;
    // ZKE_NTT_R8=1 selects the register-blocked kernel.  Measured on B200 (profiles/ntt_ab_r02.txt): it is faster alone
    // (5.37 vs 5.59 ms for the six 2^22 transforms, Montgomery twiddles) but its 160 registers per thread leave less room
    // for the blocks of other proofs' kernels, and the overlapped steady state is 1.8 % slower (53.1 vs 54.1 proofs/s).
    static const bool r8 = getenv("ZKE_NTT_R8") && atoi(getenv("ZKE_NTT_R8")) != 0;
#define ZKE_NTT_CASE(KK)                                                                                            \
    case KK:                                                                                                       \
        if (r8) {                                                                                                  \
            if (strided) ntt_pass_r8_kernel<DIF, KK, true, SHOUP><<<blocks, 128, 0, st>>>(data, tw, scale, log_tw, s_lo);  \
            else ntt_pass_r8_kernel<DIF, KK, false, SHOUP><<<blocks, 128, 0, st>>>(data, tw, scale, log_tw, s_lo);         \
        } else if (strided) ntt_pass_fast_kernel<DIF, KK, true, SHOUP><<<blocks, 512, 0, st>>>(data, tw, scale, log_tw, s_lo); \
        else ntt_pass_fast_kernel<DIF, KK, false, SHOUP><<<blocks, 512, 0, st>>>(data, tw, scale, log_tw, s_lo);            \
        return true;
    switch (k) {
        ZKE_NTT_CASE(4) ZKE_NTT_CASE(5) ZKE_NTT_CASE(6) ZKE_NTT_CASE(7) ZKE_NTT_CASE(8)
        default: return false;
    }
#undef ZKE_NTT_CASE
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

// `log_local` < T.log_n: the transform is the tail of a larger one - the 2^log_local elements at `data` are one of the
// independent blocks left after the top T.log_n - log_local stages were done elsewhere (msm sharding across GPUs);
// the twiddle tables are those of the full transform.
static void intt_dif_impl(uint8_t* data, const NttTables& T, int log_local, const uint8_t* scale_bitrev, cudaStream_t st) {
    int lo[4], hi[4], np;
    plan(log_local, lo, hi, &np);
    ZKE_COUNT_LAUNCH(np);
    const int tile_log = log_local < MAX_TILE_LOG ? log_local : MAX_TILE_LOG;
    const uint32_t blocks = 1u << (log_local - tile_log);
    for (int p = 0; p < np; ++p) {
        const uint8_t* sc = p == np - 1 ? scale_bitrev : nullptr;
        if (T.shoup ? launch_fast<true, true>(data, T.tw_inv, sc, log_local, lo[p], hi[p], st, T.log_n)
                    : launch_fast<true, false>(data, T.tw_inv, sc, log_local, lo[p], hi[p], st, T.log_n)) continue;
        if (T.shoup) ntt_pass_kernel<true, true><<<blocks, 1 << (tile_log - 1), 32u << tile_log, st>>>(data, T.tw_inv, sc, T.log_n, lo[p], hi[p], tile_log);
        else ntt_pass_kernel<true, false><<<blocks, 1 << (tile_log - 1), 32u << tile_log, st>>>(data, T.tw_inv, sc, T.log_n, lo[p], hi[p], tile_log);
    }
}
static void ntt_dit_impl(uint8_t* data, const NttTables& T, int log_local, cudaStream_t st) {
    int lo[4], hi[4], np;
    plan(log_local, lo, hi, &np);
    ZKE_COUNT_LAUNCH(np);
    const int tile_log = log_local < MAX_TILE_LOG ? log_local : MAX_TILE_LOG;
    const uint32_t blocks = 1u << (log_local - tile_log);
    for (int p = np - 1; p >= 0; --p) {
        if (T.shoup ? launch_fast<false, true>(data, T.tw_fwd, nullptr, log_local, lo[p], hi[p], st, T.log_n)
                    : launch_fast<false, false>(data, T.tw_fwd, nullptr, log_local, lo[p], hi[p], st, T.log_n)) continue;
        if (T.shoup) ntt_pass_kernel<false, true><<<blocks, 1 << (tile_log - 1), 32u << tile_log, st>>>(data, T.tw_fwd, nullptr, T.log_n, lo[p], hi[p], tile_log);
        else ntt_pass_kernel<false, false><<<blocks, 1 << (tile_log - 1), 32u << tile_log, st>>>(data, T.tw_fwd, nullptr, T.log_n, lo[p], hi[p], tile_log);
    }
}
void launch_intt_dif(uint8_t* data, const NttTables& T, const uint8_t* scale_bitrev, cudaStream_t st) { intt_dif_impl(data, T, T.log_n, scale_bitrev, st); }
void launch_ntt_dit(uint8_t* data, const NttTables& T, cudaStream_t st) { ntt_dit_impl(data, T, T.log_n, st); }
void launch_intt_dif_block(uint8_t* block, const NttTables& T, int log_local, const uint8_t* scale_block, cudaStream_t st) { intt_dif_impl(block, T, log_local, scale_block, st); }
void launch_ntt_dit_block(uint8_t* block, const NttTables& T, int log_local, cudaStream_t st) { ntt_dit_impl(block, T, log_local, st); }

// ---- the top LOGG stages of a transform whose blocks live on different GPUs -------------------------------------------
// Column layout: this GPU holds, for every one of the G = 2^LOGG blocks g, the columns [col0, col0 + n_cols) of the
// block (element (g, j) at its global position g * M + j, M = N / G).  One thread per column does the LOGG butterfly
// stages of distance M, 2M, ... in registers.  DIF (inverse transform): these are the FIRST stages, top down;
// DIT (forward): the LAST stages, bottom up.
template <bool DIF, int LOGG, bool SHOUP>
__global__ void __launch_bounds__(128)
ntt_cross_kernel(uint8_t* __restrict__ data, const uint8_t* __restrict__ tw, int log_n, uint32_t col0, uint32_t n_cols) {
    constexpr int G = 1 << LOGG;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cols) return;
    const int log_m = log_n - LOGG;
    const uint32_t j = col0 + t;
    Fr x[G];
#pragma unroll
    for (int g = 0; g < G; ++g) x[g] = Fr::load(data + 32ull * (((uint32_t)g << log_m) + j));
#pragma unroll
    for (int step = 0; step < LOGG; ++step) {
        const int sb = DIF ? (LOGG - 1 - step) : step;        // stage bit within the block index
        const int st = log_m + sb;
#pragma unroll
        for (int g0 = 0; g0 < G; ++g0) {
            if (g0 & (1 << sb)) continue;
            const int g1 = g0 | (1 << sb);
            const uint32_t idx0 = ((uint32_t)g0 << log_m) + j;
            const uint64_t tw_i = (uint64_t)(idx0 & ((1u << st) - 1)) << (log_n - 1 - st);
            Twid<SHOUP> w;
            w.load(tw, tw_i);
            const Fr u = x[g0], v = x[g1];
            if (DIF) {
                x[g0] = u + v;
                x[g1] = w.mul(u - v);
            } else {
                const Fr y = w.mul(v);
                x[g0] = u + y;
                x[g1] = u - y;
            }
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) x[g].store(data + 32ull * (((uint32_t)g << log_m) + j));
}
template <bool DIF, bool SHOUP>
static void launch_cross(uint8_t* data, const uint8_t* tw, int log_n, int log_g, uint32_t col0, uint32_t n_cols, cudaStream_t st) {
    const uint32_t blocks = (n_cols + 127) / 128;
    switch (log_g) {
        case 1: ntt_cross_kernel<DIF, 1, SHOUP><<<blocks, 128, 0, st>>>(data, tw, log_n, col0, n_cols); break;
        case 2: ntt_cross_kernel<DIF, 2, SHOUP><<<blocks, 128, 0, st>>>(data, tw, log_n, col0, n_cols); break;
        case 3: ntt_cross_kernel<DIF, 3, SHOUP><<<blocks, 128, 0, st>>>(data, tw, log_n, col0, n_cols); break;
        default: break;
    }
    ZKE_COUNT_LAUNCH(1);
}
void launch_intt_cross(uint8_t* data, const NttTables& T, int log_g, uint32_t col0, uint32_t n_cols, cudaStream_t st) {
    if (T.shoup) launch_cross<true, true>(data, T.tw_inv, T.log_n, log_g, col0, n_cols, st);
    else launch_cross<true, false>(data, T.tw_inv, T.log_n, log_g, col0, n_cols, st);
}
void launch_ntt_cross(uint8_t* data, const NttTables& T, int log_g, uint32_t col0, uint32_t n_cols, cudaStream_t st) {
    if (T.shoup) launch_cross<false, true>(data, T.tw_fwd, T.log_n, log_g, col0, n_cols, st);
    else launch_cross<false, false>(data, T.tw_fwd, T.log_n, log_g, col0, n_cols, st);
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
// the same on the columns [col0, col0 + 2^log_cols) of every block of 2^log_m elements (global positions)
__global__ void quotient_cols_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, const uint8_t* __restrict__ c,
                                     uint8_t* __restrict__ d, uint32_t n_threads, int log_m, int log_cols, uint32_t col0) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_threads) return;
    const uint32_t i = ((t >> log_cols) << log_m) + col0 + (t & ((1u << log_cols) - 1));
    Fr x = Fr::load(a + 32ull * i) * Fr::load(b + 32ull * i) - Fr::load(c + 32ull * i);
    x.from_mont().store(d + 32ull * i);
}
void launch_quotient_cols(const uint8_t* a, const uint8_t* b, const uint8_t* c, uint8_t* d, int log_n, int log_g, uint32_t col0, uint32_t n_cols_log, cudaStream_t st) {
    const uint32_t n_threads = (1u << log_g) << n_cols_log;
    quotient_cols_kernel<<<(n_threads + 255) / 256, 256, 0, st>>>(a, b, c, d, n_threads, log_n - log_g, (int)n_cols_log, col0);
    ZKE_COUNT_LAUNCH(1);
}

}  // namespace dev
}  // namespace zke
