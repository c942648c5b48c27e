// Device-side BN254 prime-field arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form (R = 2^256),
// carry chains written with mad.lo/hi.cc PTX (IMAD pipe; B300_MICROARCH "Pipe rates": IMAD 64/clk/SM).
// Memory image of an element = 32 bytes little-endian, identical to the host's U256.
//
// Role in the reference: the Fr/Fq layer of wasmcurves 0.2.0 (un-vendored; /root/reference/yarn.lock:8521-8525)
// that snarkjs' prover and circom's witness calculator run on.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace zke {
namespace dev {

struct FieldConsts {
    uint32_t mod[8];
    uint32_t r[8];    // 2^256 mod p  (Montgomery one)
    uint32_t r2[8];   // 2^512 mod p
    uint32_t inv;     // -p^-1 mod 2^32
};
// One copy per translation unit (whole-program device compilation, no -rdc): every .cu that uses field arithmetic
// instantiates ZKE_DEFINE_CONSTANT_UPLOAD(name) and the engine calls each TU's upload function once per device.
static __constant__ FieldConsts FR_C;
static __constant__ FieldConsts FQ_C;
#define ZKE_DEFINE_CONSTANT_UPLOAD(fn)                                                          \
    cudaError_t fn(const ::zke::dev::FieldConsts* fr, const ::zke::dev::FieldConsts* fq) {       \
        cudaError_t e = cudaMemcpyToSymbol(::zke::dev::FR_C, fr, sizeof(::zke::dev::FieldConsts)); \
        if (e != cudaSuccess) return e;                                                          \
        return cudaMemcpyToSymbol(::zke::dev::FQ_C, fq, sizeof(::zke::dev::FieldConsts));          \
    }

struct FrTag { static __device__ __forceinline__ const FieldConsts& C() { return FR_C; } };
struct FqTag { static __device__ __forceinline__ const FieldConsts& C() { return FQ_C; } };

// ---- carry-chain primitives --------------------------------------------------------------------
__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

template <class Tag>
struct Fp {
    uint32_t v[8];

    static __device__ __forceinline__ Fp zero() { Fp r; for (int i = 0; i < 8; ++i) r.v[i] = 0; return r; }
    static __device__ __forceinline__ Fp one() { Fp r; for (int i = 0; i < 8; ++i) r.v[i] = Tag::C().r[i]; return r; }
    static __device__ __forceinline__ Fp r2() { Fp r; for (int i = 0; i < 8; ++i) r.v[i] = Tag::C().r2[i]; return r; }
    static __device__ __forceinline__ Fp load(const void* p) {  // 32-byte aligned global / shared address
        Fp r;
        const uint4* q = reinterpret_cast<const uint4*>(p);
        uint4 a = q[0], b = q[1];
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
        return r;
    }
    __device__ __forceinline__ void store(void* p) const {
        uint4* q = reinterpret_cast<uint4*>(p);
        q[0] = make_uint4(v[0], v[1], v[2], v[3]);
        q[1] = make_uint4(v[4], v[5], v[6], v[7]);
    }
    __device__ __forceinline__ bool is_zero() const { return (v[0] | v[1] | v[2] | v[3] | v[4] | v[5] | v[6] | v[7]) == 0; }
    __device__ __forceinline__ bool operator==(const Fp& o) const {
        uint32_t d = 0;
        for (int i = 0; i < 8; ++i) d |= v[i] ^ o.v[i];
        return d == 0;
    }
    __device__ __forceinline__ bool operator!=(const Fp& o) const { return !(*this == o); }

    // r = a - p if a >= p (a < 2p)
    __device__ __forceinline__ void reduce_once() {
        const FieldConsts& C = Tag::C();
        uint32_t t[8];
        t[0] = sub_cc(v[0], C.mod[0]);
#pragma unroll
        for (int i = 1; i < 8; ++i) t[i] = subc_cc(v[i], C.mod[i]);
        uint32_t borrow = subc(0, 0);  // 0xffffffff if a < p
        if (borrow == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = t[i];
        }
    }

    friend __device__ __forceinline__ Fp operator+(const Fp& a, const Fp& b) {
        Fp r;
        r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < 8; ++i) r.v[i] = addc_cc(a.v[i], b.v[i]);
        // p < 2^254: no carry out of 256 bits
        r.reduce_once();
        return r;
    }
    friend __device__ __forceinline__ Fp operator-(const Fp& a, const Fp& b) {
        const FieldConsts& C = Tag::C();
        Fp r;
        r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < 8; ++i) r.v[i] = subc_cc(a.v[i], b.v[i]);
        uint32_t borrow = subc(0, 0);
        if (borrow) {
            r.v[0] = add_cc(r.v[0], C.mod[0]);
#pragma unroll
            for (int i = 1; i < 7; ++i) r.v[i] = addc_cc(r.v[i], C.mod[i]);
            r.v[7] = addc(r.v[7], C.mod[7]);
        }
        return r;
    }
    __device__ __forceinline__ Fp neg() const { return is_zero() ? *this : (zero() - *this); }
    __device__ __forceinline__ Fp dbl() const { return *this + *this; }

    // ---- Montgomery product a*b/2^256 mod p -------------------------------------------------------------------
    // Even/odd accumulator formulation: the running total is kept as two interleaved vectors, `even` (limb k at
    // position k) and `odd` (limb k at position k + 1), so that every 32x32 product (lo, hi) lands on an ALIGNED
    // register pair of one of them.  Each `mad.lo.cc / madc.hi.cc` pair on such a register pair is fused by ptxas
    // into a single 64-bit IMAD.WIDE with carry, instead of the IMAD + IADD3.X pair a limb-serial CIOS chain needs:
    // ~20 issue slots per row instead of ~68.  After each row the low limb of `even` is
    // zero (Montgomery step), the roles of the two vectors swap and the former `even` is realigned by the
    // shift-by-two inside madc_n_rshift.
    static __device__ __forceinline__ void mul_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
        for (int j = 0; j < 8; j += 2)
            asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(acc[j]), "=r"(acc[j + 1]) : "r"(a[j]), "r"(bi));
    }
    // acc[0..8) += a[0,2,4,6] * bi (pairs), carry chained; the final carry is left in CC
    static __device__ __forceinline__ void cmad_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(bi));
#pragma unroll
        for (int j = 2; j < 8; j += 2)
            asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[j]), "+r"(acc[j + 1]) : "r"(a[j]), "r"(bi));
    }
    // odd[j], odd[j+1] = a[j] * bi + odd[j+2], odd[j+3] (+ incoming carry): multiply-accumulate and shift down two limbs
    static __device__ __forceinline__ void madc_n_rshift(uint32_t* odd, const uint32_t* a, uint32_t bi) {
#pragma unroll
        for (int j = 0; j < 6; j += 2)
            asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
                         : "=r"(odd[j]), "=r"(odd[j + 1]) : "r"(a[j]), "r"(bi), "r"(odd[j + 2]), "r"(odd[j + 3]));
        asm volatile("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.u32 %1, %2, %3, 0;" : "=r"(odd[6]), "=r"(odd[7]) : "r"(a[6]), "r"(bi));
    }
    template <bool FIRST>
    static __device__ __forceinline__ void mad_n_redc(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi) {
        const FieldConsts& C = Tag::C();
        if (FIRST) {
            mul_n(odd, a + 1, bi);
            mul_n(even, a, bi);
        } else {
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(even[0]) : "r"(odd[1]));
            madc_n_rshift(odd, a + 1, bi);
            cmad_n(even, a, bi);
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(odd[7]));
        }
        const uint32_t mi = even[0] * C.inv;
        cmad_n(odd, C.mod + 1, mi);
        cmad_n(even, C.mod, mi);
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(odd[7]));
    }
    friend __device__ __forceinline__ Fp operator*(const Fp& a, const Fp& b) {
        uint32_t even[8], odd[8];
        mad_n_redc<true>(even, odd, a.v, b.v[0]);
        mad_n_redc<false>(odd, even, a.v, b.v[1]);
#pragma unroll
        for (int i = 2; i < 8; i += 2) {
            mad_n_redc<false>(even, odd, a.v, b.v[i]);
            mad_n_redc<false>(odd, even, a.v, b.v[i + 1]);
        }
        // merge: result limb k = even[k] + odd[k + 1]
        Fp r;
        r.v[0] = add_cc(even[0], odd[1]);
#pragma unroll
        for (int k = 1; k < 7; ++k) r.v[k] = addc_cc(even[k], odd[k + 1]);
        r.v[7] = addc(even[7], 0);
        r.reduce_once();
        return r;
    }
    __device__ __forceinline__ Fp sqr() const { return *this * *this; }
    __device__ __forceinline__ Fp to_mont() const { return *this * r2(); }
    __device__ __forceinline__ Fp from_mont() const {
        Fp o = zero(); o.v[0] = 1;
        return *this * o;
    }
    // Fermat inverse (Montgomery in/out); inv(0) = 0.  ~380 products - use sparingly / batch.
    __device__ Fp inv() const {
        const FieldConsts& C = Tag::C();
        uint32_t e[8];
        e[0] = sub_cc(C.mod[0], 2);
#pragma unroll
        for (int i = 1; i < 8; ++i) e[i] = subc_cc(C.mod[i], 0);
        Fp res = one(), base = *this;
        for (int w = 7; w >= 0; --w) {
            for (int bit = 31; bit >= 0; --bit) {
                res = res.sqr();
                if ((e[w] >> bit) & 1) res = res * base;
            }
        }
        return res;
    }
};

typedef Fp<FrTag> Fr;
typedef Fp<FqTag> Fq;

// Fq2 = Fq[u]/(u^2+1)
struct Fq2 {
    Fq c0, c1;
    static __device__ __forceinline__ Fq2 zero() { Fq2 r; r.c0 = Fq::zero(); r.c1 = Fq::zero(); return r; }
    static __device__ __forceinline__ Fq2 one() { Fq2 r; r.c0 = Fq::one(); r.c1 = Fq::zero(); return r; }
    static __device__ __forceinline__ Fq2 load(const void* p) { Fq2 r; r.c0 = Fq::load(p); r.c1 = Fq::load((const char*)p + 32); return r; }
    __device__ __forceinline__ void store(void* p) const { c0.store(p); c1.store((char*)p + 32); }
    __device__ __forceinline__ bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    __device__ __forceinline__ bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    friend __device__ __forceinline__ Fq2 operator+(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; return r; }
    friend __device__ __forceinline__ Fq2 operator-(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; return r; }
    // not inlined: a G2 addition holds ten of these; inlining them all makes the G2 kernels compile for minutes
    // and spill (the call costs ~20 instructions against ~600 of work)
    friend __device__ __noinline__ Fq2 operator*(const Fq2& a, const Fq2& b) {
        Fq t0 = a.c0 * b.c0, t1 = a.c1 * b.c1;
        Fq2 r;
        r.c1 = (a.c0 + a.c1) * (b.c0 + b.c1) - t0 - t1;
        r.c0 = t0 - t1;
        return r;
    }
    __device__ __noinline__ Fq2 sqr() const {
        Fq2 r;
        Fq t = c0 * c1;
        r.c0 = (c0 + c1) * (c0 - c1);
        r.c1 = t + t;
        return r;
    }
    __device__ __forceinline__ Fq2 neg() const { Fq2 r; r.c0 = c0.neg(); r.c1 = c1.neg(); return r; }
    __device__ __forceinline__ Fq2 dbl() const { return *this + *this; }
    __device__ Fq2 inv() const {
        Fq d = (c0.sqr() + c1.sqr()).inv();
        Fq2 r; r.c0 = c0 * d; r.c1 = (c1 * d).neg();
        return r;
    }
};

}  // namespace dev
}  // namespace zke
