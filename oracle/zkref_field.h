/*
 * zkref_field.h - CPU oracle: BN254 Fr / Fq arithmetic (plain C, 4 x 64-bit limbs, unsigned __int128).
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build or call anything under oracle/.  The product (zk-email-verify_b200/) never links this.
 *
 * Restates the field layer of ffjavascript 0.2.56 / wasmcurves 0.2.0 (un-vendored dependencies of the
 * reference's snarkjs fork: /root/reference/packages/helpers/package.json:26, /root/reference/yarn.lock:4646-4652,
 * 8521-8525): Montgomery multiplication with R = 2^256, elements exchanged as 32-byte little-endian integers.
 */
#ifndef ZKREF_FIELD_H
#define ZKREF_FIELD_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } fe;   /* one field element (meaning - standard or Montgomery - by context) */

typedef struct {
    fe p, r, r2;       /* modulus, 2^256 mod p, 2^512 mod p */
    uint64_t inv;      /* -p^-1 mod 2^64 */
} field_t;

extern field_t ZK_FR, ZK_FQ;
void zkref_init(void);   /* idempotent */

static inline int fe_is_zero(const fe* a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) { return a->v[0] == b->v[0] && a->v[1] == b->v[1] && a->v[2] == b->v[2] && a->v[3] == b->v[3]; }
static inline int fe_cmp(const fe* a, const fe* b) {
    for (int i = 3; i >= 0; --i) { if (a->v[i] < b->v[i]) return -1; if (a->v[i] > b->v[i]) return 1; }
    return 0;
}
static inline uint64_t fe_add_raw(fe* r, const fe* a, const fe* b) {
    u128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (u128)a->v[i] + b->v[i]; r->v[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static inline uint64_t fe_sub_raw(fe* r, const fe* a, const fe* b) {
    uint64_t bw = 0;
    for (int i = 0; i < 4; ++i) { u128 d = (u128)a->v[i] - b->v[i] - bw; r->v[i] = (uint64_t)d; bw = (uint64_t)(d >> 64) & 1; }
    return bw;
}
static inline void f_add(const field_t* F, fe* r, const fe* a, const fe* b) {
    uint64_t c = fe_add_raw(r, a, b);
    if (c || fe_cmp(r, &F->p) >= 0) fe_sub_raw(r, r, &F->p);
}
static inline void f_sub(const field_t* F, fe* r, const fe* a, const fe* b) {
    if (fe_sub_raw(r, a, b)) fe_add_raw(r, r, &F->p);
}
static inline void f_neg(const field_t* F, fe* r, const fe* a) {
    if (fe_is_zero(a)) { *r = *a; return; }
    fe_sub_raw(r, &F->p, a);
}
/* Montgomery product a*b/2^256 mod p */
static inline void f_mul(const field_t* F, fe* r, const fe* a, const fe* b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) { c += (u128)a->v[j] * b->v[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * F->inv;
        c = (u128)m * F->p.v[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; ++j) { c += (u128)m * F->p.v[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    fe o = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || fe_cmp(&o, &F->p) >= 0) fe_sub_raw(&o, &o, &F->p);
    *r = o;
}
static inline void f_sqr(const field_t* F, fe* r, const fe* a) { f_mul(F, r, a, a); }
static inline void f_to_mont(const field_t* F, fe* r, const fe* a) { f_mul(F, r, a, &F->r2); }
static inline void f_from_mont(const field_t* F, fe* r, const fe* a) { fe one = {{1, 0, 0, 0}}; f_mul(F, r, a, &one); }
void f_pow(const field_t* F, fe* r, const fe* a_mont, const fe* e);   /* Montgomery in/out */
void f_inv(const field_t* F, fe* r, const fe* a_mont);                /* Fermat; inv(0) = 0 */
void f_batch_inv(const field_t* F, fe* a, size_t n);                  /* Montgomery form, zeros stay zero */

#endif
