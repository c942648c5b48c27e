// Big-integer hint of FpMul, shared by host (unit tests) and device (witness kernel).
//
// circom computes q, r of FpMul with hint functions on n-bit limbs held in field registers:
//     var long_div_out[2][100] = long_div(n, k, k, ab_proper, p);   q[i] <-- ...; r[i] <-- ...;
// (/root/reference/packages/circuits/lib/fp.circom:32-50, lib/bigint-func.circom:32-53, 65-103, 169-264).
// Because q and r are range-checked to n bits, r < p and a*b = q*p + r is enforced exactly, the hint is the
// integer statement (q, r) = divmod(A*B, P) with A = sum a_i 2^(n i) etc..  This file computes
// that with Knuth's algorithm D on 32-bit words.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define ZKE_HD __host__ __device__
#else
#define ZKE_HD
#endif

namespace zke {

static const int BIGDIV_MAXW = 80;  // words per operand (n*k + 256 bits <= 2560)

// acc[0..W) += (val[0..8) << shift)
ZKE_HD inline void bd_add_shifted(uint32_t* acc, int W, const uint32_t* val, uint32_t shift) {
    const uint32_t ws = shift >> 5, bs = shift & 31;
    uint64_t carry = 0;
    uint32_t prev = 0;
    for (int i = 0; i < 9; ++i) {
        uint32_t cur = i < 8 ? val[i] : 0;
        uint32_t piece = bs ? ((cur << bs) | (prev >> (32 - bs))) : cur;
        prev = cur;
        int idx = (int)ws + i;
        if (idx >= W) break;
        uint64_t s = (uint64_t)acc[idx] + piece + carry;
        acc[idx] = (uint32_t)s;
        carry = s >> 32;
    }
    for (int idx = (int)ws + 9; carry && idx < W; ++idx) {
        uint64_t s = (uint64_t)acc[idx] + carry;
        acc[idx] = (uint32_t)s;
        carry = s >> 32;
    }
}

// out (8 words) = bits [bit, bit + nbits) of x[0..W)
ZKE_HD inline void bd_extract(const uint32_t* x, int W, uint32_t bit, uint32_t nbits, uint32_t* out) {
    for (int i = 0; i < 8; ++i) out[i] = 0;
    for (uint32_t done = 0; done < nbits; done += 32) {
        uint32_t pos = bit + done;
        uint32_t ws = pos >> 5, bs = pos & 31;
        uint32_t lo = (int)ws < W ? x[ws] : 0;
        uint32_t hi = (int)(ws + 1) < W ? x[ws + 1] : 0;
        uint32_t word = bs ? ((lo >> bs) | (hi << (32 - bs))) : lo;
        uint32_t rem = nbits - done;
        if (rem < 32) word &= (1u << rem) - 1;
        out[done >> 5] = word;
    }
}

// Knuth's algorithm D: U[0..UL] (UL + 2 words allocated, U[UL + 1] scratch) divided by P[0..t) (top word non-zero, t >= 1;
// P is normalised in place and NOT restored).  On return Q[0..UL] holds the quotient and U[0..t) the remainder.
ZKE_HD inline void bd_knuth_div(uint32_t* U, int UL, uint32_t* P, int t, uint32_t* Q) {
    // normalise so that the top bit of P[t-1] is set
    int s = 0;
    while (((P[t - 1] << s) & 0x80000000u) == 0) ++s;
    if (s) {
        for (int i = t - 1; i > 0; --i) P[i] = (P[i] << s) | (P[i - 1] >> (32 - s));
        P[0] <<= s;
        for (int i = UL; i > 0; --i) U[i] = (U[i] << s) | (U[i - 1] >> (32 - s));
        U[0] <<= s;
    }
    if (t == 1) {
        uint64_t rem = 0;
        for (int j = UL; j >= 0; --j) {
            uint64_t cur = (rem << 32) | U[j];
            Q[j] = (uint32_t)(cur / P[0]);
            rem = cur % P[0];
            U[j] = 0;
        }
        U[0] = (uint32_t)rem;
    } else {
        const uint64_t v1 = P[t - 1], v2 = P[t - 2];
        for (int j = UL - t; j >= 0; --j) {
            uint64_t num = ((uint64_t)U[j + t] << 32) | U[j + t - 1];
            uint64_t qhat = num / v1, rhat = num % v1;
            while (qhat >= 0x100000000ull || qhat * v2 > ((rhat << 32) | U[j + t - 2])) {
                --qhat;
                rhat += v1;
                if (rhat >= 0x100000000ull) break;
            }
            // U[j .. j+t] -= qhat * P
            uint64_t borrow = 0, carry = 0;
            for (int i = 0; i < t; ++i) {
                uint64_t prod = qhat * P[i] + carry;
                carry = prod >> 32;
                uint64_t sub = (uint64_t)U[i + j] - (uint32_t)prod - borrow;
                U[i + j] = (uint32_t)sub;
                borrow = (sub >> 32) & 1;
            }
            uint64_t sub = (uint64_t)U[j + t] - carry - borrow;
            U[j + t] = (uint32_t)sub;
            if ((sub >> 32) & 1) {  // qhat was one too large: add back
                --qhat;
                uint64_t c = 0;
                for (int i = 0; i < t; ++i) {
                    uint64_t sum = (uint64_t)U[i + j] + P[i] + c;
                    U[i + j] = (uint32_t)sum;
                    c = sum >> 32;
                }
                U[j + t] += (uint32_t)c;
            }
            Q[j] = (uint32_t)qhat;
        }
    }
    // de-normalise the remainder (low t words of U)
    if (s) {
        for (int i = 0; i < t; ++i) U[i] = (U[i] >> s) | (i + 1 <= UL ? (U[i + 1] << (32 - s)) : 0);
    }
}

// a_limbs / b_limbs / p_limbs: k values of 8 words each (witness values).  q_out / r_out: k values of 8 words each.
// Returns 0 on success, non-zero if the parameters exceed the supported size.
ZKE_HD inline int fpmul_hint_words(uint32_t n, uint32_t k, const uint32_t* a_limbs, const uint32_t* b_limbs,
                                   const uint32_t* p_limbs, uint32_t* q_out, uint32_t* r_out) {
    const int W = (int)((n * k + 256 + 31) / 32);
    if (W > BIGDIV_MAXW || n > 128 || n == 0) return 1;
    uint32_t A[BIGDIV_MAXW], B[BIGDIV_MAXW], P[BIGDIV_MAXW + 1];
    uint32_t U[2 * BIGDIV_MAXW + 2];   // dividend (normalised in place), ends up holding the remainder
    uint32_t Q[2 * BIGDIV_MAXW + 1];
    for (int i = 0; i < W; ++i) { A[i] = 0; B[i] = 0; P[i] = 0; }
    for (uint32_t i = 0; i < k; ++i) {
        bd_add_shifted(A, W, a_limbs + 8 * i, n * i);
        bd_add_shifted(B, W, b_limbs + 8 * i, n * i);
        bd_add_shifted(P, W, p_limbs + 8 * i, n * i);
    }
    const int UL = 2 * W;
    for (int i = 0; i < UL + 2; ++i) U[i] = 0;
    for (int i = 0; i < UL + 1; ++i) Q[i] = 0;
    for (int i = 0; i < W; ++i) {
        uint64_t carry = 0;
        const uint64_t ai = A[i];
        if (ai == 0) continue;
        for (int j = 0; j < W; ++j) {
            uint64_t s = ai * B[j] + U[i + j] + carry;
            U[i + j] = (uint32_t)s;
            carry = s >> 32;
        }
        U[i + W] = (uint32_t)carry;
    }
    int t = W;
    while (t > 0 && P[t - 1] == 0) --t;
    if (t == 0) {  // division by zero: circom's long_div would fail its own asserts; emit zeros, constraints reject
        for (uint32_t i = 0; i < 8 * k; ++i) { q_out[i] = 0; r_out[i] = 0; }
        return 0;
    }
    bd_knuth_div(U, UL, P, t, Q);
    for (int i = t; i < UL + 2; ++i) U[i] = 0;
    for (uint32_t i = 0; i < k; ++i) {
        bd_extract(Q, UL + 1, n * i, n, q_out + 8 * i);
        bd_extract(U, UL + 2, n * i, n, r_out + 8 * i);
    }
    return 0;
}

}  // namespace zke
