// Streaming batched-affine bucket accumulation for the H multi-exponentiation (round 2).
//
// The bucket sums of a Pippenger MSM are sums of ~100 affine table points each.  In inversion-free XYZZ coordinates
// an addition costs 8M + 2S (1,232 IMAD.WIDE); in affine coordinates it costs 5M + 1S (788) plus one inversion, and
// Montgomery's trick shares ONE inversion among all additions that are independent of each other.  All pairwise
// additions of one level of the buckets' addition trees are independent, so the accumulation becomes a few
// level-synchronous sweeps over flat arrays - every sweep a plain grid-wide kernel with coalesced streams:
//
//   level l input : per bucket b, n_l(b) = ceil(size(b) / 2^l) affine points, contiguous at off_l[b]
//                   (level 0: the table points themselves, copied into bucket order by the digit scatter kernel - the
//                   table is read once, coalesced, and the sweeps never gather)
//   forward  (A)  : one thread per K consecutive OUTPUT slots: for slot s = (b, i) with 2i + 1 < n_l(b) form the slope
//                   denominator d = x2 - x1 of points 2i, 2i + 1, multiply it into a running product and store the
//                   prefix; the thread's total goes to `tot`
//   invert        : tot[t] <- 1 / tot[t] for all threads, by a hierarchical grid-wide batch inversion (groups of 32,
//                   32, ... until one thread holds the rest: a single Fermat inversion per level of the tree)
//   backward (B)  : the same thread walks its slots backwards: 1/d = v * prefix_before, v *= d, then the affine
//                   addition (lambda, x3, y3) into the level-(l+1) array; an odd last point of a bucket is copied
//
// After `levels` sweeps each bucket holds ceil(size / 2^levels) points; they are finished by the running-sum kernel
// with mixed additions.  Products per bucket entry: (1 - 2^-levels) * 6 + ~0.2 (inversion tree) instead of 10.
// DRAM traffic is ~26 GB per 2^22-point MSM, all of it streaming except the one table gather.
#pragma once
#include "ba.cuh"

namespace zke {
namespace dev {

static const int BA_K = 8;            // output slots per thread
static const int BA_THREADS = 128;
static const int BINV_G = 32;         // group size of the hierarchical batch inversion

// n_l(b) for a bucket with `size` entries
__device__ __forceinline__ uint32_t ba_level_count(uint32_t size, uint32_t level) { return (size + (1u << level) - 1) >> level; }

struct BaLevel {
    const uint32_t* hist;         // bucket sizes (level 0)
    const uint32_t* off_in;       // exclusive scan of n_l
    const uint32_t* off_out;      // exclusive scan of n_{l+1}; off_out[n_buckets] = number of output slots
    const uint32_t* slot_bucket;  // bucket of every output slot
    uint32_t n_buckets, level;
};

// ---- forward sweep ---------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(BA_THREADS, 4)
ba_forward_kernel(BaLevel L, const Affine<F>* __restrict__ pts_in, F* __restrict__ prefix, F* __restrict__ tot, uint32_t n_threads) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_threads) return;
    const uint32_t total = L.off_out[L.n_buckets];
    const uint32_t s0 = t * BA_K;
    F run = F::one();
#pragma unroll 1
    for (uint32_t j = 0; j < BA_K; j += 4) {
        // four slots per step: their x-coordinate loads are issued together
        F px[4], qx[4];
        bool live[4], pair[4];
        uint32_t in0[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t s = s0 + j + u;
            live[u] = s < total;
            pair[u] = false;
            if (!live[u]) continue;
            const uint32_t b = L.slot_bucket[s];
            const uint32_t i = s - L.off_out[b];
            const uint32_t n_in = ba_level_count(L.hist[b], L.level);
            in0[u] = L.off_in[b] + 2 * i;
            pair[u] = 2 * i + 1 < n_in;
            if (pair[u]) {
                // only the x-coordinates are needed unless the pair is a doubling / cancellation / has an infinity
                px[u] = F::load(&pts_in[in0[u]].x);
                qx[u] = F::load(&pts_in[in0[u] + 1].x);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!live[u]) continue;
            const uint32_t s = s0 + j + u;
            if (pair[u]) {
                F den;
                if (px[u].is_zero() || qx[u].is_zero() || px[u] == qx[u])
                    ba_den(Affine<F>::load(pts_in + in0[u]), Affine<F>::load(pts_in + in0[u] + 1), den);
                else den = qx[u] - px[u];
                run = run * den;
            }
            run.store(prefix + s);
        }
    }
    run.store(tot + t);
}

// ---- backward sweep --------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(BA_THREADS, 4)
ba_backward_kernel(BaLevel L, const Affine<F>* __restrict__ pts_in, const F* __restrict__ prefix,
                   const F* __restrict__ tot_inv, Affine<F>* __restrict__ pts_out, uint32_t n_threads) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_threads) return;
    const uint32_t total = L.off_out[L.n_buckets];
    const uint32_t s0 = t * BA_K;
    if (s0 >= total) return;
    F v = F::load(tot_inv + t);
    const uint32_t last = min(total, s0 + BA_K);
    // operands of the slot to be processed next are loaded one slot ahead
    auto fetch = [&](uint32_t s, Affine<F>& P, Affine<F>& Q, bool& pair, F& prev) {
        const uint32_t b = L.slot_bucket[s];
        const uint32_t i = s - L.off_out[b];
        const uint32_t n_in = ba_level_count(L.hist[b], L.level);
        pair = 2 * i + 1 < n_in;
        const uint32_t in0 = L.off_in[b] + 2 * i;
        P = Affine<F>::load(pts_in + in0);
        if (pair) Q = Affine<F>::load(pts_in + in0 + 1);
        if (pair) prev = s > s0 ? F::load(prefix + s - 1) : F::one();
    };
    Affine<F> P, Q;
    bool pair;
    F prev;
    fetch(last - 1, P, Q, pair, prev);
#pragma unroll 1
    for (uint32_t s = last; s-- > s0;) {
        Affine<F> Pn, Qn;
        bool pair_n = false;
        F prev_n;
        if (s > s0) fetch(s - 1, Pn, Qn, pair_n, prev_n);
        if (pair) {
            F den;
            const int kind = ba_den(P, Q, den);
            const F inv_den = v * prev;
            v = v * den;
            ba_apply(kind, P, Q, inv_den).store(pts_out + s);
        } else {
            P.store(pts_out + s);
        }
        P = Pn; Q = Qn; pair = pair_n; prev = prev_n;
    }
}

// ---- grid-wide batch inversion ---------------------------------------------------------------------------------
// up: group g multiplies its BINV_G values, storing the running prefix (inclusive) and the group total
template <class F>
__global__ void __launch_bounds__(128)
binv_up_kernel(const F* __restrict__ vals, uint32_t n, F* __restrict__ pre, F* __restrict__ group_tot, uint32_t n_groups) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint32_t beg = g * BINV_G, end = min(n, beg + BINV_G);
    F run = F::one();
    for (uint32_t i = beg; i < end; ++i) { run = run * F::load(vals + i); run.store(pre + i); }
    run.store(group_tot + g);
}
// down: group g receives 1 / (its total) and replaces every member by its inverse
template <class F>
__global__ void __launch_bounds__(128)
binv_down_kernel(F* __restrict__ vals, uint32_t n, const F* __restrict__ pre, const F* __restrict__ group_inv, uint32_t n_groups) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint32_t beg = g * BINV_G, end = min(n, beg + BINV_G);
    F v = F::load(group_inv + g);
    for (uint32_t i = end; i-- > beg;) {
        const F x = F::load(vals + i);
        const F inv = i > beg ? v * F::load(pre + i - 1) : v;
        v = v * x;
        inv.store(vals + i);
    }
}
// top: one thread inverts the remaining (<= BINV_TOP) values with a single field inversion
static const uint32_t BINV_TOP = 64;
template <class F>
__global__ void binv_top_kernel(F* vals, uint32_t n) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    F pre[BINV_TOP];
    F run = F::one();
    for (uint32_t i = 0; i < n; ++i) { run = run * F::load(vals + i); pre[i] = run; }
    F v = run.inv();
    for (uint32_t i = n; i-- > 0;) {
        const F x = F::load(vals + i);
        const F inv = i > 0 ? v * pre[i - 1] : v;
        v = v * x;
        inv.store(vals + i);
    }
}

// number of field elements of scratch batch_inverse needs for n values
static inline size_t binv_scratch_elems(size_t n) {
    size_t total = 0;
    while (n > BINV_TOP) { const size_t g = (n + BINV_G - 1) / BINV_G; total += n + g; n = g; }
    return total + 8;
}
// vals[i] <- 1 / vals[i] (all non-zero), i < n
template <class F>
static void batch_inverse(F* vals, uint32_t n, F* scratch, cudaStream_t st) {
    struct Lvl { F* vals; F* pre; uint32_t n; };
    Lvl lv[8];
    int depth = 0;
    F* cur = vals;
    uint32_t cn = n;
    F* sp = scratch;
    while (cn > BINV_TOP) {
        const uint32_t g = (cn + BINV_G - 1) / BINV_G;
        F* pre = sp; sp += cn;
        F* nxt = sp; sp += g;
        binv_up_kernel<F><<<(g + 127) / 128, 128, 0, st>>>(cur, cn, pre, nxt, g);
        lv[depth++] = Lvl{cur, pre, cn};
        cur = nxt; cn = g;
    }
    binv_top_kernel<F><<<1, 32, 0, st>>>(cur, cn);
    ZKE_COUNT_LAUNCH(depth + 1);
    for (int d = depth; d-- > 0;) {
        const uint32_t g = (lv[d].n + BINV_G - 1) / BINV_G;
        binv_down_kernel<F><<<(g + 127) / 128, 128, 0, st>>>(lv[d].vals, lv[d].n, lv[d].pre, cur, g);
        cur = lv[d].vals;
    }
    ZKE_COUNT_LAUNCH(depth);
}

// ---- workspace ---------------------------------------------------------------------------------------------------
// Slot counts are data dependent; every array is sized for the upper bound (max_entries = n * windows entries):
// level-(l+1) slots <= max_entries / 2^(l+1) + n_buckets.
template <class F>
struct BaPlan {
    size_t max_slots1 = 0;     // upper bound of level-1 slots (outputs of the first sweep)
    size_t n_threads1 = 0;
    static size_t slots_bound(size_t max_entries, size_t n_buckets, int level) { return (max_entries >> level) + n_buckets + 1; }
    static size_t bytes(size_t max_entries, size_t n_buckets, int levels) {
        const size_t s1 = slots_bound(max_entries, n_buckets, 1), s2 = slots_bound(max_entries, n_buckets, 2);
        const size_t t1 = (s1 + BA_K - 1) / BA_K;
        size_t b = 0;
        auto al = [&](size_t x) { b += (x + 255) & ~(size_t)255; };
        al(sizeof(Affine<F>) * (max_entries + 1));   // level-0 points in bucket order (written by the digit scatter)
        al(sizeof(F) * s1);                     // prefix
        al(sizeof(Affine<F>) * s1);             // points ping
        al(sizeof(Affine<F>) * s2);             // points pong
        al(4 * s1);                             // slot -> bucket
        al(sizeof(F) * t1);                     // thread totals
        al(sizeof(F) * binv_scratch_elems(t1));
        for (int l = 0; l <= levels; ++l) al(4 * (n_buckets + 1));   // per-level offsets
        return b;
    }
};

}  // namespace dev
}  // namespace zke
