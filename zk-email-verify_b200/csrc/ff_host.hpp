// Host-side BN254 prime-field arithmetic (4 x 64-bit limbs, Montgomery form) used by the circuit
// front-end, the toy trusted setup and the verifier.  Device code has its own 8 x 32-bit version
// in ff.cuh; both agree on the little-endian 32-byte memory image of a field element.
//
// Replaces (for the EmailVerifier path) the field layer the reference obtains from the un-vendored
// ffjavascript 0.2.56 / wasmcurves 0.2.0 stack (/root/reference/yarn.lock:4646-4652, 8521-8525).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>

namespace zke {

typedef unsigned __int128 u128;

struct U256 {
    uint64_t v[4];
    bool operator==(const U256& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2] && v[3] == o.v[3]; }
    bool operator!=(const U256& o) const { return !(*this == o); }
    bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
};

inline int u256_cmp(const U256& a, const U256& b) {
    for (int i = 3; i >= 0; --i) {
        if (a.v[i] < b.v[i]) return -1;
        if (a.v[i] > b.v[i]) return 1;
    }
    return 0;
}
inline uint64_t u256_add(U256& r, const U256& a, const U256& b) {
    u128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (u128)a.v[i] + b.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
inline uint64_t u256_sub(U256& r, const U256& a, const U256& b) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a.v[i] - b.v[i] - borrow;
        r.v[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}
inline bool u256_bit(const U256& a, unsigned i) { return (a.v[i >> 6] >> (i & 63)) & 1; }
inline U256 u256_shr(const U256& a, unsigned s) {
    U256 r = {{0, 0, 0, 0}};
    if (s >= 256) return r;
    unsigned w = s >> 6, b = s & 63;
    for (unsigned i = 0; i + w < 4; ++i) {
        r.v[i] = a.v[i + w] >> b;
        if (b && i + w + 1 < 4) r.v[i] |= a.v[i + w + 1] << (64 - b);
    }
    return r;
}
U256 u256_from_dec(const std::string& s);   // throws on bad digits / overflow
std::string u256_to_dec(const U256& a);
U256 u256_from_hex(const char* s);

// Parameters of one prime field.
struct FieldParams {
    U256 p;        // modulus
    U256 r;        // 2^256 mod p      (Montgomery one)
    U256 r2;       // 2^512 mod p
    uint64_t inv;  // -p^{-1} mod 2^64
};
const FieldParams& fr_params();
const FieldParams& fq_params();

// Element of a prime field, stored in Montgomery form.
template <const FieldParams& (*PARAMS)()>
struct Fp {
    U256 m;  // Montgomery representation

    static const FieldParams& P() { return PARAMS(); }
    static Fp zero() { Fp r; r.m = U256{{0, 0, 0, 0}}; return r; }
    static Fp one() { Fp r; r.m = P().r; return r; }
    static Fp from_u256(const U256& x) {  // x < p, standard form
        Fp t; t.m = x; Fp r2; r2.m = P().r2; return mont_mul(t, r2);
    }
    static Fp from_u64(uint64_t x) { return from_u256(U256{{x, 0, 0, 0}}); }
    static Fp from_i64(int64_t x) { return x >= 0 ? from_u64((uint64_t)x) : from_u64((uint64_t)(-x)).neg(); }
    static Fp from_dec(const std::string& s) {
        U256 x = u256_from_dec(s);
        if (u256_cmp(x, P().p) >= 0) throw std::runtime_error("field element out of range");
        return from_u256(x);
    }
    U256 to_u256() const {  // standard form
        Fp o; o.m = U256{{1, 0, 0, 0}}; return mont_mul(*this, o).m;
    }
    bool is_zero() const { return m.is_zero(); }
    bool operator==(const Fp& o) const { return m == o.m; }
    bool operator!=(const Fp& o) const { return m != o.m; }

    static Fp mont_mul(const Fp& a, const Fp& b) {
        const U256& p = P().p;
        const uint64_t inv = P().inv;
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            u128 c = 0;
            for (int j = 0; j < 4; ++j) {
                c += (u128)a.m.v[j] * b.m.v[i] + t[j];
                t[j] = (uint64_t)c; c >>= 64;
            }
            c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
            uint64_t mm = t[0] * inv;
            c = (u128)mm * p.v[0] + t[0]; c >>= 64;
            for (int j = 1; j < 4; ++j) {
                c += (u128)mm * p.v[j] + t[j];
                t[j - 1] = (uint64_t)c; c >>= 64;
            }
            c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
        }
        Fp r; r.m = U256{{t[0], t[1], t[2], t[3]}};
        if (t[4] || u256_cmp(r.m, p) >= 0) u256_sub(r.m, r.m, p);
        return r;
    }
    Fp operator*(const Fp& o) const { return mont_mul(*this, o); }
    Fp operator+(const Fp& o) const {
        Fp r; uint64_t c = u256_add(r.m, m, o.m);
        if (c || u256_cmp(r.m, P().p) >= 0) u256_sub(r.m, r.m, P().p);
        return r;
    }
    Fp operator-(const Fp& o) const {
        Fp r; if (u256_sub(r.m, m, o.m)) u256_add(r.m, r.m, P().p);
        return r;
    }
    Fp neg() const { if (is_zero()) return *this; Fp r; u256_sub(r.m, P().p, m); return r; }
    Fp& operator+=(const Fp& o) { *this = *this + o; return *this; }
    Fp& operator-=(const Fp& o) { *this = *this - o; return *this; }
    Fp& operator*=(const Fp& o) { *this = *this * o; return *this; }
    Fp sqr() const { return mont_mul(*this, *this); }
    Fp pow(const U256& e) const {
        Fp res = one();
        for (int i = 255; i >= 0; --i) {
            res = res.sqr();
            if (u256_bit(e, i)) res = res * *this;
        }
        return res;
    }
    Fp pow_u64(uint64_t e) const { return pow(U256{{e, 0, 0, 0}}); }
    Fp inv() const {  // Fermat; inv(0) = 0
        U256 e; U256 two = {{2, 0, 0, 0}}; u256_sub(e, P().p, two);
        return pow(e);
    }
};

typedef Fp<fr_params> Fr;
typedef Fp<fq_params> Fq;

// Montgomery batch inversion (zeros stay zero).
template <class F>
void batch_inverse(F* a, size_t n) {
    std::vector<F> pre(n);
    F acc = F::one();
    for (size_t i = 0; i < n; ++i) { pre[i] = acc; if (!a[i].is_zero()) acc = acc * a[i]; }
    acc = acc.inv();
    for (size_t i = n; i-- > 0;) {
        if (a[i].is_zero()) continue;
        F t = acc * pre[i]; acc = acc * a[i]; a[i] = t;
    }
}

// 2-adic root of unity of Fr: generator 5, 2^28 | r - 1.
Fr fr_root_of_unity(unsigned log_n);   // primitive 2^log_n-th root

}  // namespace zke
