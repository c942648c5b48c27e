#include "ff_host.hpp"
#include <algorithm>

namespace zke {

U256 u256_from_dec(const std::string& s) {
    if (s.empty()) throw std::runtime_error("empty decimal string");
    U256 r = {{0, 0, 0, 0}};
    for (char ch : s) {
        if (ch < '0' || ch > '9') throw std::runtime_error("bad decimal digit in '" + s + "'");
        u128 c = (unsigned)(ch - '0');
        for (int i = 0; i < 4; ++i) { c += (u128)r.v[i] * 10; r.v[i] = (uint64_t)c; c >>= 64; }
        if (c) throw std::runtime_error("decimal value exceeds 256 bits");
    }
    return r;
}

std::string u256_to_dec(const U256& a) {
    U256 t = a;
    std::string out;
    if (t.is_zero()) return "0";
    while (!t.is_zero()) {
        u128 rem = 0;
        for (int i = 3; i >= 0; --i) {
            u128 cur = (rem << 64) | t.v[i];
            t.v[i] = (uint64_t)(cur / 10);
            rem = cur % 10;
        }
        out.push_back((char)('0' + (int)rem));
    }
    std::reverse(out.begin(), out.end());
    return out;
}

U256 u256_from_hex(const char* s) {
    U256 r = {{0, 0, 0, 0}};
    if (s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) s += 2;
    for (; *s; ++s) {
        int d;
        if (*s >= '0' && *s <= '9') d = *s - '0';
        else if (*s >= 'a' && *s <= 'f') d = *s - 'a' + 10;
        else if (*s >= 'A' && *s <= 'F') d = *s - 'A' + 10;
        else throw std::runtime_error("bad hex digit");
        r.v[3] = (r.v[3] << 4) | (r.v[2] >> 60);
        r.v[2] = (r.v[2] << 4) | (r.v[1] >> 60);
        r.v[1] = (r.v[1] << 4) | (r.v[0] >> 60);
        r.v[0] = (r.v[0] << 4) | (uint64_t)d;
    }
    return r;
}

static FieldParams make_params(const char* dec) {
    FieldParams fp;
    fp.p = u256_from_dec(dec);
    // inv = -p^{-1} mod 2^64 (Newton)
    uint64_t x = 1;
    for (int i = 0; i < 6; ++i) x *= 2 - fp.p.v[0] * x;
    fp.inv = (uint64_t)0 - x;
    // r = 2^256 mod p by doubling 1, 256 times; r2 by doubling a further 256 times
    U256 t = {{1, 0, 0, 0}};
    for (int i = 0; i < 512; ++i) {
        uint64_t c = u256_add(t, t, t);
        if (c || u256_cmp(t, fp.p) >= 0) u256_sub(t, t, fp.p);
        if (i == 255) fp.r = t;
    }
    fp.r2 = t;
    return fp;
}

const FieldParams& fr_params() {
    static const FieldParams p = make_params("21888242871839275222246405745257275088548364400416034343698204186575808495617");
    return p;
}
const FieldParams& fq_params() {
    static const FieldParams p = make_params("21888242871839275222246405745257275088696311157297823662689037894645226208583");
    return p;
}

Fr fr_root_of_unity(unsigned log_n) {
    if (log_n > 28) throw std::runtime_error("BN254 Fr has 2-adicity 28");
    // w28 = 5^((r-1)/2^28)
    U256 e; U256 one = {{1, 0, 0, 0}};
    u256_sub(e, fr_params().p, one);
    e = u256_shr(e, 28);
    Fr w = Fr::from_u64(5).pow(e);
    for (unsigned i = log_n; i < 28; ++i) w = w.sqr();
    return w;
}

}  // namespace zke
