// H bucket accumulation with the Montgomery REDUCTIONS on the tensor cores (ff_tc.cuh): the same per-chunk XYZZ sums as
// chunk_sum_kernel<Fq> (msm.cu), computed by warp-collective mixed additions - the ten products of an addition are
// five pairs of 512-bit schoolbook / half products (64 / 36 IMAD.WIDE) whose reductions run as int8 contractions on
// the IMMA pipe: 669 multiplier instructions per addition instead of 1,232.  Every lane of a warp walks its own chunk;
// lanes whose chunk is exhausted keep taking part in the collectives with the point at infinity (chunks are handed
// out in size order, so the lanes of a warp have near-equal lengths).  Opt-in: ZKE_H_TC=1 (see DESIGN.md section 5
// for the measurements).  G1 only.
#pragma once
#include "ec.cuh"
#include "ff_tc.cuh"
#include <mutex>

namespace zke {
namespace dev {

// acc += (+/-) p for the lanes with `live` set; all 32 lanes must call it
__device__ __forceinline__ void madd_tc(XYZZ<Fq>& acc, const Affine<Fq>& p, bool negate, bool live, const TcLane& L) {
    typedef FpTc<FqTag> T;
    const bool use = live && !p.is_inf();
    const bool acc_inf = acc.is_inf();
    const Fq py = negate ? p.y.neg() : p.y;
    Fq U2, S2;
    T::mul2(U2, S2, p.x, acc.zz, py, acc.zzz, L);
    const Fq P = U2 - acc.x, R = S2 - acc.y;
    Fq PP, RR;
    T::sqr2(PP, RR, P, R, L);
    Fq PPP, Q;
    T::mul2(PPP, Q, P, PP, acc.x, PP, L);
    const Fq X3 = RR - PPP - Q.dbl();
    Fq Ya, Yb;
    T::mul2(Ya, Yb, R, Q - X3, acc.y, PPP, L);
    Fq ZZ3, ZZZ3;
    T::mul2(ZZ3, ZZZ3, acc.zz, PP, acc.zzz, PPP, L);
    if (use) {
        if (acc_inf) {
            acc.x = p.x; acc.y = py; acc.zz = Fq::one(); acc.zzz = Fq::one();
        } else if (P.is_zero()) {      // same x: doubling or cancellation (never on random points; not collective)
            if (R.is_zero()) acc.dbl(); else acc = XYZZ<Fq>::inf();
        } else {
            acc.x = X3; acc.y = Ya - Yb; acc.zz = ZZ3; acc.zzz = ZZZ3;
        }
    }
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB)
chunk_sum_tc_kernel(const uint8_t* __restrict__ points, const uint32_t* __restrict__ entries, const uint32_t* __restrict__ offsets,
                    const uint32_t* __restrict__ hist, const uint32_t* __restrict__ chunk_off, const uint32_t* __restrict__ work_bucket,
                    const uint32_t* __restrict__ order, uint32_t n_buckets, uint32_t CHUNK, uint8_t* partial, const TcTable* __restrict__ tab) {
    __shared__ __align__(16) uint32_t scratch[4][2 * TC_SCRATCH_WORDS];
    TcLane L;
    L.init(tab, scratch[threadIdx.x >> 5]);
    const uint32_t total = chunk_off[n_buckets];
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < total; base += gridDim.x * blockDim.x) {   // warp-uniform trip count
        const uint32_t t = base + lane;
        const bool valid = t < total;
        uint32_t w = 0, beg = 0, end = 0;
        if (valid) {
            w = order[t];
            const uint32_t b = work_bucket[w];
            const uint32_t ci = w - chunk_off[b];
            beg = offsets[b] + ci * CHUNK;
            end = min(offsets[b] + hist[b], beg + CHUNK);
        }
        const uint32_t len = end - beg;
        const uint32_t maxlen = __reduce_max_sync(0xffffffffu, len);
        XYZZ<Fq> acc = XYZZ<Fq>::inf();
        for (uint32_t i = 0; i < maxlen; ++i) {
            const bool live = i < len;
            Affine<Fq> p;
            bool neg = false;
            if (live) {
                const uint32_t e = entries[beg + i];
                p = Affine<Fq>::load(points + sizeof(Affine<Fq>) * (size_t)(e & 0x7fffffffu));
                neg = (e >> 31) != 0;
            } else {
                p.x = Fq::zero(); p.y = Fq::zero();
            }
            madd_tc(acc, p, neg, live, L);
        }
        if (valid) acc.store(partial + sizeof(XYZZ<Fq>) * (size_t)w);
    }
}

// the reduction matrix of Fq in fragment order, one copy per device (built from the modulus the engine uploaded)
inline const TcTable* tc_table_fq() {
    static std::mutex mu;
    static TcTable* tabs[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!tabs[dev]) {
        FieldConsts fq;
        if (cudaMemcpyFromSymbol(&fq, FQ_C, sizeof(FieldConsts)) != cudaSuccess) return nullptr;
        TcTable host;
        tc_build_table(fq.mod, &host);
        TcTable* d = nullptr;
        if (cudaMalloc(&d, sizeof(TcTable)) != cudaSuccess) return nullptr;
        if (cudaMemcpy(d, &host, sizeof(TcTable), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
        tabs[dev] = d;
    }
    return tabs[dev];
}

}  // namespace dev
}  // namespace zke
