// Batched-affine point accumulation ("BA"): sums lists of affine points pairwise with ONE shared field inversion per
// tree level instead of inversion-free extended coordinates.
//
// An affine addition needs lambda = (y2 - y1) / (x2 - x1).  Montgomery's trick turns the K divisions of K
// independent additions into 3 (K - 1) products and one inversion: with prefix products p_i = d_0 ... d_i,
//   1 / d_i = (1 / p_i) p_{i-1},   1 / p_{i-1} = (1 / p_i) d_i.
// One addition then costs 6 products (1 prefix, 2 back-substitution, lambda, lambda^2, y3) against 10 for the XYZZ
// mixed addition (8M + 2S) - the bucket accumulation of the H multi-exponentiation is bound by exactly that product
// count.  The single inversion (Fermat, ~380 products) is shared by a whole thread block
// (msm.cu: ba_chunk_sum_kernel), so its cost per addition is below 1 % of a product.
//
// This header holds the per-thread arithmetic, written so that it also compiles under ZKE_FF_EMULATE
// (tests/test_ba_emulation.py checks it against Python integers).
#pragma once
#include "ec.cuh"
#include <cstddef>

namespace zke {
namespace dev {

// non-binding request to bring the 32-byte sectors at p (and p + 32) into the cache hierarchy
__device__ __forceinline__ void ba_prefetch(const void* p, bool two_sectors) {
#ifndef ZKE_FF_EMULATE
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    if (two_sectors) asm volatile("prefetch.global.L1 [%0];" ::"l"((const char*)p + 32));
#else
    (void)p; (void)two_sectors;
#endif
}

enum BaKind { BA_ADD = 0, BA_DBL = 1, BA_KEEP_P = 2, BA_KEEP_Q = 3, BA_INF = 4 };

// classifies P + Q and returns the denominator of its slope (never zero; one when no slope is needed)
template <class F>
__device__ __forceinline__ int ba_den(const Affine<F>& P, const Affine<F>& Q, F& den) {
    const bool pi = P.is_inf(), qi = Q.is_inf();
    if (pi || qi) { den = F::one(); return pi ? (qi ? BA_INF : BA_KEEP_Q) : BA_KEEP_P; }
    if (!(P.x == Q.x)) { den = Q.x - P.x; return BA_ADD; }
    if (P.y == Q.y && !P.y.is_zero()) { den = P.y.dbl(); return BA_DBL; }
    den = F::one();
    return BA_INF;   // P + (-P), or a 2-torsion point (none on BN254)
}

template <class F>
__device__ __forceinline__ Affine<F> ba_apply(int kind, const Affine<F>& P, const Affine<F>& Q, const F& inv_den) {
    Affine<F> r;
    if (kind == BA_KEEP_P) return P;
    if (kind == BA_KEEP_Q) return Q;
    if (kind == BA_INF) { r.x = F::zero(); r.y = F::zero(); return r; }
    F lambda, x3;
    if (kind == BA_ADD) {
        lambda = (Q.y - P.y) * inv_den;
        x3 = lambda.sqr() - P.x - Q.x;
    } else {
        const F xx = P.x.sqr();
        lambda = (xx.dbl() + xx) * inv_den;
        x3 = lambda.sqr() - P.x.dbl();
    }
    r.x = x3;
    r.y = lambda * (P.x - x3) - P.y;
    return r;
}

// Phase A over the pairs (slot 2i, slot 2i + 1), i < np, of a point source: the running product `run` (carried in
// from the thread's previous chunks) is multiplied by each slope denominator and stored: pref[i * stride] = product
// of everything up to and including pair i.  Four pairs per step: the eight x-coordinates (table gathers for level 0)
// are requested before the first product so that their latency overlaps; a pair only fetches the y-coordinates when
// its x-coordinates are equal or zero (point at infinity, doubling, cancellation - rare).
template <class F, class Source>
__device__ __forceinline__ void ba_phase_a(const Source& src, int np, F* pref, size_t stride, F& run) {
    for (int i = 0; i < np; i += 4) {
        F px[4], qx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i + u < np) { px[u] = src.get_x(2 * (i + u)); qx[u] = src.get_x(2 * (i + u) + 1); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i + u < np) {
                F den;
                if (px[u].is_zero() || qx[u].is_zero() || px[u] == qx[u]) ba_den(src.get(2 * (i + u)), src.get(2 * (i + u) + 1), den);
                else den = qx[u] - px[u];
                run = run * den;
                run.store(pref + (size_t)(i + u) * stride);
            }
        }
    }
}

// Phase B: v = 1 / (running product after this chunk's last pair) on entry, 1 / (running product before its first
// pair) on exit; `before` = the running product before the chunk's first pair.  Writes out[i * stride] = slot 2i +
// slot 2i + 1.  Software-pipelined by one pair in registers: the operands of pair i - 1 are loaded before pair i is
// computed (measured faster than prefetch.global.L1 hints, which mostly miss by the time the data is used).
template <class F, class Source>
__device__ __forceinline__ void ba_phase_b(const Source& src, int np, const F* pref, size_t stride, const F& before, F& v,
                                           Affine<F>* out, size_t out_stride) {
    if (np <= 0) return;
    Affine<F> P = src.get(2 * (np - 1)), Q = src.get(2 * (np - 1) + 1);
    F prev = np > 1 ? F::load(pref + (size_t)(np - 2) * stride) : before;
    for (int i = np - 1; i >= 0; --i) {
        Affine<F> Pn = P, Qn = Q;
        F prevn = before;
        if (i > 0) {
            Pn = src.get(2 * (i - 1));
            Qn = src.get(2 * (i - 1) + 1);
            if (i > 1) prevn = F::load(pref + (size_t)(i - 2) * stride);
        }
        F den;
        const int kind = ba_den(P, Q, den);
        const F inv_den = v * prev;
        v = v * den;
        ba_apply(kind, P, Q, inv_den).store(out + (size_t)i * out_stride);
        P = Pn; Q = Qn; prev = prevn;
    }
}

// point sources: the (sign-tagged) entries of a bucket chunk in the point table, or a buffer of intermediate sums
template <class F>
struct BaTableSource {
    const uint8_t* points;
    const uint32_t* entries;
    __device__ __forceinline__ Affine<F> get(int slot) const {
        const uint32_t e = entries[slot];
        Affine<F> p = Affine<F>::load(points + sizeof(Affine<F>) * (size_t)(e & 0x7fffffffu));
        if (e >> 31) p.y = p.y.neg();
        return p;
    }
    __device__ __forceinline__ F get_x(int slot) const { return F::load(points + sizeof(Affine<F>) * (size_t)(entries[slot] & 0x7fffffffu)); }
    __device__ __forceinline__ void prefetch(int slot) const { ba_prefetch(points + sizeof(Affine<F>) * (size_t)(entries[slot] & 0x7fffffffu), sizeof(F) == 32); }
};
template <class F>
struct BaBufferSource {
    const Affine<F>* buf;
    size_t stride;      // in points: element `slot` of this thread's list is buf[slot * stride]
    __device__ __forceinline__ Affine<F> get(int slot) const { return Affine<F>::load(buf + (size_t)slot * stride); }
    __device__ __forceinline__ F get_x(int slot) const { return F::load(buf + (size_t)slot * stride); }
    __device__ __forceinline__ void prefetch(int slot) const { ba_prefetch(buf + (size_t)slot * stride, sizeof(F) == 32); }
};

}  // namespace dev
}  // namespace zke
