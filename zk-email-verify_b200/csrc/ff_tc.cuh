// Warp-collective Montgomery reduction on the tensor cores (sm_100a, legacy mma.sync IMMA path).
//
// Why: every heavy kernel of this engine is bound by the integer-multiply pipe, and 72 of the 136
// IMAD.WIDE of a Montgomery product are the reduction - multiplications by CONSTANTS.  A reduction of a 512-bit
// T = a * b is a linear map of the low half:
//     T * 2^-256  ==  T_hi + sum_{j<32} t_j * C_j   (mod p),     t_j = byte j of T_lo,  C_j = 2^(8j - 256) mod p,
// i.e. a (32 bytes) x (32 x 32 bytes) matrix product with a constant matrix: exactly an int8 tensor-core contraction.
// One warp reduces its 32 elements (one per lane) with 2 x 4 mma.m16n8k32.u8 instructions; the column sums (< 2^21
// each, byte-granular weights) are folded back into limbs on the ALU pipe, T_hi is added, and one 14-bit Barrett
// step (1 IMAD.HI + 8 IMAD.WIDE) brings the 268-bit value into [0, 2p).  Multiplier instructions per product:
// 64 + 9 instead of 136 (square: 36 + 9 instead of 108); the result is the same canonical a*b/2^256 mod p.
//
// Data movement: the rows of the A operand are the lanes' T_lo (2 x STS.128 per lane, 2 x ldmatrix.x4 per warp); the
// columns of the constant matrix are permuted so that lane (g, t) of a quad ends up with the sums of bytes 8t .. 8t+7
// of its four rows, folds them into three words, and hands them to the owner lanes with three stmatrix.x4 (the
// fragment layout of stmatrix is exactly the transposition needed) - 40 shared-memory wavefronts per warp product.
// All 32 lanes must call redc() convergently (mma.sync / ldmatrix are .aligned).
//
// Role in the reference: part of the Fr/Fq layer of wasmcurves 0.2.0 (un-vendored), see ff.cuh.
#pragma once
#include "ff.cuh"

namespace zke {
namespace dev {

static const int TC_SCRATCH_WORDS = 768;     // per warp and per simultaneous reduction: 32 input rows x 48 bytes, then three hand-over planes of 32 x 16 bytes

// Per-lane constants: the eight B-fragment registers of the reduction matrix (4 column tiles x 2 k-halves) and the
// Barrett reciprocal floor(2^285 / p).  Built on the host by tc_build_table(); one table per field in global memory.
struct TcTable {
    uint32_t bfrag[32][8];
    uint32_t mu;
    uint32_t pad[7];
};

struct TcLane {
    uint32_t b[8];
    uint32_t mu;
    uint32_t* scratch;     // this warp's TC_SCRATCH_WORDS words of shared memory (16-byte aligned)
    __device__ __forceinline__ void init(const TcTable* tab, uint32_t* warp_scratch) {
        const int lane = threadIdx.x & 31;
        const uint4 lo = reinterpret_cast<const uint4*>(tab->bfrag[lane])[0], hi = reinterpret_cast<const uint4*>(tab->bfrag[lane])[1];
        b[0] = lo.x; b[1] = lo.y; b[2] = lo.z; b[3] = lo.w; b[4] = hi.x; b[5] = hi.y; b[6] = hi.z; b[7] = hi.w;
        mu = tab->mu;
        scratch = warp_scratch;
    }
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel)); return r; }

template <class Tag>
struct FpTc {
    typedef Fp<Tag> F;

    // T[i][0..16) -> T[i] / 2^256 mod p, canonical, for NB independent values per lane.  Warp-collective.  NB = 2 gives
    // the scheduler two independent dependency graphs between the two warp synchronisations (the column folding and the
    // carry chains are latency-, not throughput-bound); the warp's scratch holds NB areas of TC_SCRATCH_WORDS words.
    template <int NB>
    static __device__ __forceinline__ void redc_n(F* r, const uint32_t (*T)[16], const TcLane& L) {
        const FieldConsts& C = Tag::C();
        const int lane = threadIdx.x & 31;
        uint32_t* S = L.scratch;
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            uint4* row = reinterpret_cast<uint4*>(S + i * TC_SCRATCH_WORDS + lane * 12);    // 48-byte row stride: conflict-free STS.128 / ldmatrix
            row[0] = make_uint4(T[i][0], T[i][1], T[i][2], T[i][3]);
            row[1] = make_uint4(T[i][4], T[i][5], T[i][6], T[i][7]);
        }
        __syncwarp();
        // ldmatrix lane -> row address: matrix m = lane / 8 (m & 1: rows 8..15, m >> 1: bytes 16..31), row lane % 8
        const uint32_t lm = smem_addr(S) + (uint32_t)((((lane >> 3) & 1) * 8 + (lane & 7)) * 48 + (lane >> 4) * 16);
        // slot k = 2 mt + h <-> row 16 mt + 8 h + g = the element of lane 8 k + g; w[i][p][k]: word p of this lane's 8-byte slice
        uint32_t w[NB][3][4];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                uint32_t a[4];
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(lm + (uint32_t)(i * TC_SCRATCH_WORDS * 4 + mt * 16 * 48)) : "memory");
                int32_t d[4][4];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};"
                                 : "=r"(d[nt][0]), "=r"(d[nt][1]), "=r"(d[nt][2]), "=r"(d[nt][3])
                                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(L.b[2 * nt]), "r"(L.b[2 * nt + 1]), "r"(0));
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    // byte b = 2 nt + c of the slice is d[nt][2 h + c] (< 2^21); y_c = bytes 2c, 2c + 1 (< 2^30), value = sum y_c 2^(16 c).
                    // The 16-bit shifts are byte permutes so that ptxas cannot turn them into IMAD.WIDE by 0x10000 (it did, for
                    // a shift-and-add written in C or as add.cc / addc: 16 extra multiplier instructions per product).
                    const uint32_t y0 = (uint32_t)d[0][2 * h] + ((uint32_t)d[0][2 * h + 1] << 8);
                    const uint32_t y1 = (uint32_t)d[1][2 * h] + ((uint32_t)d[1][2 * h + 1] << 8);
                    const uint32_t y2 = (uint32_t)d[2][2 * h] + ((uint32_t)d[2][2 * h + 1] << 8);
                    const uint32_t y3 = (uint32_t)d[3][2 * h] + ((uint32_t)d[3][2 * h + 1] << 8);
                    const uint32_t lo = prmt(y1, 0u, 0x1044u), mid = prmt(y1, y3, 0x5432u), hi = prmt(y3, 0u, 0x4432u);
                    w[i][0][2 * mt + h] = add_cc(y0, lo);
                    w[i][1][2 * mt + h] = addc_cc(y2, mid);
                    w[i][2][2 * mt + h] = addc(hi, 0);
                }
            }
        }
        // hand-over: plane p, 16-byte row e = (word p of the four slices of element e), written in fragment order
        const uint32_t out = smem_addr(S) + 1536u + (uint32_t)lane * 16u;
#pragma unroll
        for (int i = 0; i < NB; ++i)
#pragma unroll
            for (int p = 0; p < 3; ++p)
                asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};"
                             :: "r"(out + (uint32_t)(i * TC_SCRATCH_WORDS * 4 + p * 512)), "r"(w[i][p][0]), "r"(w[i][p][1]), "r"(w[i][p][2]), "r"(w[i][p][3]) : "memory");
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            uint32_t Lw[9];
            const uint4* mine = reinterpret_cast<const uint4*>(S + i * TC_SCRATCH_WORDS + 384 + lane * 4);
            const uint4 p0 = mine[0], p1 = mine[32], p2 = mine[64];
            Lw[0] = p0.x; Lw[1] = p1.x;
            Lw[2] = add_cc(p2.x, p0.y); Lw[3] = addc_cc(p1.y, 0);
            Lw[4] = addc_cc(p2.y, p0.z); Lw[5] = addc_cc(p1.z, 0);
            Lw[6] = addc_cc(p2.z, p0.w); Lw[7] = addc_cc(p1.w, 0);
            Lw[8] = addc(p2.w, 0);
            Lw[0] = add_cc(Lw[0], T[i][8]);
#pragma unroll
            for (int k = 1; k < 8; ++k) Lw[k] = addc_cc(Lw[k], T[i][8 + k]);
            Lw[8] = addc(Lw[8], 0);
            // Barrett: q = floor(floor(r' / 2^240) * mu / 2^45) is the quotient digit or one below it: r' - q p in [0, 2p)
            const uint32_t x = (Lw[8] << 16) | (Lw[7] >> 16);
            const uint32_t q = __umulhi(x, L.mu) >> 13;
            uint32_t od[8];
            F::mul_n(od, C.nmod + 1, q);                            // limbs 1.. of q * (2^256 - p), odd positions
            F::cmad_n(Lw, C.nmod, q);                               // even positions, on top of r'
            r[i].v[0] = Lw[0];
            r[i].v[1] = add_cc(Lw[1], od[0]);
#pragma unroll
            for (int k = 2; k < 7; ++k) r[i].v[k] = addc_cc(Lw[k], od[k - 1]);
            r[i].v[7] = addc(Lw[7], od[6]);
            r[i].reduce_once();
        }
    }
    // ---- software-pipelined form: r = redc(Tp) while the NEXT 512-bit product Tn = a * b is formed by the same warp.
    // The reduction is a latency chain (store, ldmatrix, IMMA, fold, stmatrix, load, carry chains) on the shared-memory,
    // tensor and ALU pipes; the schoolbook product is 64 back-to-back IMAD.WIDE.  Run one after the other, the warps of a
    // scheduler convoy: all of them queue for the multiplier, then all of them wait out the reduction.  Here the eight
    // rows of the product are issued between the steps of the reduction, so every warp always has both kinds of work in
    // flight.  Tn must not depend on the result r (the caller alternates between independent values).
    struct WideRows {
        uint32_t ev[16], od[16];     // od[k] sits at limb position k + 1
        int ee, eo;
        __device__ __forceinline__ WideRows() : ee(0), eo(0) {}
        template <int I>
        __device__ __forceinline__ void row(const uint32_t* a, uint32_t bi) {
            if ((I & 1) == 0) {
                ee = F::template chain<4>(ev, ee, I, a, 0, 2, bi);
                eo = F::template chain<4>(od, eo, I, a, 1, 2, bi);
            } else {
                eo = F::template chain<4>(od, eo, I - 1, a, 0, 2, bi);
                ee = F::template chain<4>(ev, ee, I + 1, a, 1, 2, bi);
            }
        }
        __device__ __forceinline__ void finish(uint32_t* T) const {
            T[0] = ev[0];
            T[1] = add_cc(ev[1], od[0]);
#pragma unroll
            for (int k = 2; k < 16; ++k) {
                const uint32_t e = k < ee ? ev[k] : 0, o = (k - 1) < eo ? od[k - 1] : 0;
                T[k] = k + 1 < 16 ? addc_cc(e, o) : addc(e, o);
            }
        }
    };
    static __device__ __forceinline__ F redc_mul(const uint32_t* Tp, uint32_t* Tn, const F& a, const F& b, const TcLane& L) {
        const FieldConsts& C = Tag::C();
        const int lane = threadIdx.x & 31;
        uint32_t* S = L.scratch;
        WideRows W;
        {
            uint4* row = reinterpret_cast<uint4*>(S + lane * 12);
            row[0] = make_uint4(Tp[0], Tp[1], Tp[2], Tp[3]);
            row[1] = make_uint4(Tp[4], Tp[5], Tp[6], Tp[7]);
        }
        W.template row<0>(a.v, b.v[0]);
        W.template row<1>(a.v, b.v[1]);
        __syncwarp();
        const uint32_t lm = smem_addr(S) + (uint32_t)((((lane >> 3) & 1) * 8 + (lane & 7)) * 48 + (lane >> 4) * 16);
        uint32_t af[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                         : "=r"(af[mt][0]), "=r"(af[mt][1]), "=r"(af[mt][2]), "=r"(af[mt][3]) : "r"(lm + (uint32_t)(mt * 16 * 48)) : "memory");
        W.template row<2>(a.v, b.v[2]);
        W.template row<3>(a.v, b.v[3]);
        uint32_t w[3][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            int32_t d[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};"
                             : "=r"(d[nt][0]), "=r"(d[nt][1]), "=r"(d[nt][2]), "=r"(d[nt][3])
                             : "r"(af[mt][0]), "r"(af[mt][1]), "r"(af[mt][2]), "r"(af[mt][3]), "r"(L.b[2 * nt]), "r"(L.b[2 * nt + 1]), "r"(0));
            if (mt == 0) W.template row<4>(a.v, b.v[4]); else W.template row<5>(a.v, b.v[5]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t y0 = (uint32_t)d[0][2 * h] + ((uint32_t)d[0][2 * h + 1] << 8);
                const uint32_t y1 = (uint32_t)d[1][2 * h] + ((uint32_t)d[1][2 * h + 1] << 8);
                const uint32_t y2 = (uint32_t)d[2][2 * h] + ((uint32_t)d[2][2 * h + 1] << 8);
                const uint32_t y3 = (uint32_t)d[3][2 * h] + ((uint32_t)d[3][2 * h + 1] << 8);
                const uint32_t lo = prmt(y1, 0u, 0x1044u), mid = prmt(y1, y3, 0x5432u), hi = prmt(y3, 0u, 0x4432u);
                w[0][2 * mt + h] = add_cc(y0, lo);
                w[1][2 * mt + h] = addc_cc(y2, mid);
                w[2][2 * mt + h] = addc(hi, 0);
            }
        }
        const uint32_t out = smem_addr(S) + 1536u + (uint32_t)lane * 16u;
#pragma unroll
        for (int p = 0; p < 3; ++p)
            asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};"
                         :: "r"(out + (uint32_t)(p * 512)), "r"(w[p][0]), "r"(w[p][1]), "r"(w[p][2]), "r"(w[p][3]) : "memory");
        W.template row<6>(a.v, b.v[6]);
        W.template row<7>(a.v, b.v[7]);
        __syncwarp();
        uint32_t Lw[9];
        {
            const uint4* mine = reinterpret_cast<const uint4*>(S + 384 + lane * 4);
            const uint4 p0 = mine[0], p1 = mine[32], p2 = mine[64];
            W.finish(Tn);                                          // the merge of the new product covers the load latency
            Lw[0] = p0.x; Lw[1] = p1.x;
            Lw[2] = add_cc(p2.x, p0.y); Lw[3] = addc_cc(p1.y, 0);
            Lw[4] = addc_cc(p2.y, p0.z); Lw[5] = addc_cc(p1.z, 0);
            Lw[6] = addc_cc(p2.z, p0.w); Lw[7] = addc_cc(p1.w, 0);
            Lw[8] = addc(p2.w, 0);
        }
        Lw[0] = add_cc(Lw[0], Tp[8]);
#pragma unroll
        for (int k = 1; k < 8; ++k) Lw[k] = addc_cc(Lw[k], Tp[8 + k]);
        Lw[8] = addc(Lw[8], 0);
        const uint32_t x = (Lw[8] << 16) | (Lw[7] >> 16);
        const uint32_t q = __umulhi(x, L.mu) >> 13;
        uint32_t od[8];
        F::mul_n(od, C.nmod + 1, q);
        F::cmad_n(Lw, C.nmod, q);
        F r;
        r.v[0] = Lw[0];
        r.v[1] = add_cc(Lw[1], od[0]);
#pragma unroll
        for (int k = 2; k < 7; ++k) r.v[k] = addc_cc(Lw[k], od[k - 1]);
        r.v[7] = addc(Lw[7], od[6]);
        r.reduce_once();
        return r;
    }

    static __device__ __forceinline__ F redc(const uint32_t* T, const TcLane& L) {
        F r;
        redc_n<1>(&r, reinterpret_cast<const uint32_t (*)[16]>(T), L);
        return r;
    }
    // (a*b, c*d): the two reductions share their synchronisation points
    static __device__ __forceinline__ void mul2(F& r0, F& r1, const F& a, const F& b, const F& c, const F& d, const TcLane& L) {
        uint32_t T[2][16];
        F::template mul_wide<8>(T[0], a.v, b.v);
        F::template mul_wide<8>(T[1], c.v, d.v);
        F r[2];
        redc_n<2>(r, T, L);
        r0 = r[0]; r1 = r[1];
    }
    // (a^2, c^2)
    static __device__ __forceinline__ void sqr2(F& r0, F& r1, const F& a, const F& c, const TcLane& L) {
        uint32_t T[2][16];
        F::sqr_wide(T[0], a.v);
        F::sqr_wide(T[1], c.v);
        F r[2];
        redc_n<2>(r, T, L);
        r0 = r[0]; r1 = r[1];
    }
    // (a^2, c*d)
    static __device__ __forceinline__ void sqr_mul(F& r0, F& r1, const F& a, const F& c, const F& d, const TcLane& L) {
        uint32_t T[2][16];
        F::sqr_wide(T[0], a.v);
        F::template mul_wide<8>(T[1], c.v, d.v);
        F r[2];
        redc_n<2>(r, T, L);
        r0 = r[0]; r1 = r[1];
    }

    static __device__ __forceinline__ F mul(const F& a, const F& b, const TcLane& L) {
        uint32_t T[16];
        F::template mul_wide<8>(T, a.v, b.v);
        return redc(T, L);
    }
    static __device__ __forceinline__ F sqr(const F& a, const TcLane& L) {
        uint32_t T[16];
        F::sqr_wide(T, a.v);
        return redc(T, L);
    }
};

#ifndef ZKE_FF_EMULATE
// Host side: the constant matrix in fragment order.  mod = the field modulus, 8 little-endian limbs.
// C_j = 2^(8j - 256) mod p by repeated halving from C_32 = 1; column `n` of tile `nt` holds byte 8 (n / 2) + 2 nt + (n & 1).
inline void tc_build_table(const uint32_t* mod, TcTable* out) {
    auto halve = [&](uint32_t* x) {
        uint64_t carry = 0;
        if (x[0] & 1) { for (int i = 0; i < 8; ++i) { carry += (uint64_t)x[i] + mod[i]; x[i] = (uint32_t)carry; carry >>= 32; } }
        for (int i = 0; i < 8; ++i) x[i] = (x[i] >> 1) | (i < 7 ? x[i + 1] << 31 : (uint32_t)carry << 31);
    };
    uint8_t Cb[32][32];
    uint32_t cur[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 31; j >= 0; --j) {
        for (int h = 0; h < 8; ++h) halve(cur);
        for (int k = 0; k < 32; ++k) Cb[j][k] = (uint8_t)(cur[k >> 2] >> (8 * (k & 3)));
    }
    for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3;
        for (int nt = 0; nt < 4; ++nt) {
            const int byte = 8 * (g >> 1) + 2 * nt + (g & 1);
            for (int half = 0; half < 2; ++half) {
                uint32_t w = 0;
                for (int i = 0; i < 4; ++i) w |= (uint32_t)Cb[16 * half + 4 * t + i][byte] << (8 * i);
                out->bfrag[lane][2 * nt + half] = w;
            }
        }
    }
    // mu = floor(2^285 / p): long division of 2^285 by the top 96 bits is not exact enough; do it bit by bit on 9 limbs
    uint32_t rem[10] = {0}, mu = 0;       // rem < p always
    for (int bit = 285; bit >= 0; --bit) {
        // rem = 2 rem + (bit == 285)
        uint32_t c = bit == 285 ? 1u : 0u;
        for (int i = 0; i < 9; ++i) { const uint32_t n = (rem[i] << 1) | c; c = rem[i] >> 31; rem[i] = n; }
        // if rem >= p: rem -= p, quotient bit 1
        bool ge = rem[8] != 0;
        if (!ge) { ge = true; for (int i = 7; i >= 0; --i) if (rem[i] != mod[i]) { ge = rem[i] > mod[i]; break; } }
        if (ge) {
            uint64_t br = 0;
            for (int i = 0; i < 9; ++i) { const uint64_t dd = (uint64_t)rem[i] - (i < 8 ? mod[i] : 0) - br; rem[i] = (uint32_t)dd; br = (dd >> 32) & 1; }
            if (bit < 32) mu |= 1u << bit;
        }
    }
    out->mu = mu;
    for (int i = 0; i < 7; ++i) out->pad[i] = 0;
}
#endif

}  // namespace dev
}  // namespace zke
