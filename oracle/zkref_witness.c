/*
 * zkref_witness.c - CPU oracle: sequential witness calculator and R1CS checker.
 *
 * TEST INFRASTRUCTURE (see oracle/__init__.py).  Plays the role of circom_tester's
 *   calculateWitness(input)  /  checkConstraints(witness)
 * (/root/reference/packages/circuits/tests/email-verifier.test.ts:43-44) for the circuits emitted by the
 * product's front-end: it walks the witness program one op at a time, in program order, with plain
 * big-integer arithmetic, and re-evaluates every R1CS row <A,w> * <B,w> = <C,w>.
 * The hint ops restate the circom `<--` assignments cited next to each case.
 */
#include "zkref.h"
#include <stdlib.h>

field_t ZK_FR, ZK_FQ;
static int g_init = 0;

static void field_setup(field_t* F, const uint64_t p[4]) {
    memcpy(F->p.v, p, 32);
    uint64_t x = 1;
    for (int i = 0; i < 6; ++i) x *= 2 - p[0] * x;
    F->inv = (uint64_t)0 - x;
    fe t = {{1, 0, 0, 0}};
    for (int i = 0; i < 512; ++i) {
        uint64_t c = fe_add_raw(&t, &t, &t);
        if (c || fe_cmp(&t, &F->p) >= 0) fe_sub_raw(&t, &t, &F->p);
        if (i == 255) F->r = t;
    }
    F->r2 = t;
}
void zkref_init(void) {
    if (g_init) return;
    /* r = 21888242871839275222246405745257275088548364400416034343698204186575808495617 */
    static const uint64_t R[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    /* q = 21888242871839275222246405745257275088696311157297823662689037894645226208583 */
    static const uint64_t Q[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    field_setup(&ZK_FR, R);
    field_setup(&ZK_FQ, Q);
    g_init = 1;
}
void f_pow(const field_t* F, fe* r, const fe* a, const fe* e) {
    fe res = F->r, base = *a;
    for (int i = 255; i >= 0; --i) {
        f_sqr(F, &res, &res);
        if ((e->v[i >> 6] >> (i & 63)) & 1) f_mul(F, &res, &res, &base);
    }
    *r = res;
}
void f_inv(const field_t* F, fe* r, const fe* a) {
    fe e, two = {{2, 0, 0, 0}};
    fe_sub_raw(&e, &F->p, &two);
    f_pow(F, r, a, &e);
}
void f_batch_inv(const field_t* F, fe* a, size_t n) {
    fe* pre = (fe*)malloc(sizeof(fe) * (n ? n : 1));
    fe acc = F->r;
    for (size_t i = 0; i < n; ++i) { pre[i] = acc; if (!fe_is_zero(&a[i])) f_mul(F, &acc, &acc, &a[i]); }
    f_inv(F, &acc, &acc);
    for (size_t i = n; i-- > 0;) {
        if (fe_is_zero(&a[i])) continue;
        fe t; f_mul(F, &t, &acc, &pre[i]); f_mul(F, &acc, &acc, &a[i]); a[i] = t;
    }
    free(pre);
}

/* ---- linear combination: sum coef * w, coefficients pre-multiplied by R so one Montgomery product per term */
static void eval_lc(const fe* coef_r, const uint32_t* ptr, const uint32_t* var, const uint32_t* coef, uint32_t id,
                    const fe* w, fe* out) {
    fe acc = {{0, 0, 0, 0}};
    for (uint32_t k = ptr[id]; k < ptr[id + 1]; ++k) {
        fe t;
        f_mul(&ZK_FR, &t, &coef_r[coef[k]], &w[var[k]]);
        f_add(&ZK_FR, &acc, &acc, &t);
    }
    *out = acc;
}

/* ---- multi-precision helpers for the FpMul hint (little-endian 64-bit words) */
#define BIGW 80
static void big_add_shifted(uint64_t* acc, const fe* x, unsigned shift) {  /* acc += x << shift */
    unsigned w = shift >> 6, b = shift & 63;
    uint64_t limbs[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        limbs[i] |= x->v[i] << b;
        if (b) limbs[i + 1] |= x->v[i] >> (64 - b);
    }
    u128 c = 0;
    for (unsigned i = 0; w + i < BIGW; ++i) {
        c += (u128)acc[w + i] + (i < 5 ? limbs[i] : 0);
        acc[w + i] = (uint64_t)c;
        c >>= 64;
        if (i >= 5 && c == 0) break;
    }
}
static int big_ge(const uint64_t* a, const uint64_t* b, int n) {
    for (int i = n - 1; i >= 0; --i) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; }
    return 1;
}
static void big_sub(uint64_t* a, const uint64_t* b, int n) {
    uint64_t bw = 0;
    for (int i = 0; i < n; ++i) { u128 d = (u128)a[i] - b[i] - bw; a[i] = (uint64_t)d; bw = (uint64_t)(d >> 64) & 1; }
}
static void big_extract(const uint64_t* x, unsigned bit, unsigned nbits, fe* out) {  /* nbits <= 128 */
    memset(out, 0, sizeof(*out));
    for (unsigned i = 0; i < nbits; ++i) {
        unsigned pos = bit + i;
        if (pos >= 64 * BIGW) break;
        if ((x[pos >> 6] >> (pos & 63)) & 1) out->v[i >> 6] |= 1ull << (i & 63);
    }
}
/* q[i], r[i] <-- long_div(a*b, p)   (/root/reference/packages/circuits/lib/fp.circom:32-50, hint functions
 * lib/bigint-func.circom:32-53,169-218).  Equivalent integer statement: (q, r) = divmod(A*B, P) on the integers
 * A = sum a_i 2^(n i) etc., re-split into n-bit limbs (unique by the circuit's range checks). */
static void fpmul_hint(uint32_t n, uint32_t k, const uint32_t* av, const uint32_t* bv, const uint32_t* pv, fe* w, uint32_t dst) {
    uint64_t A[BIGW], B[BIGW], P[BIGW], prod[BIGW], Q[BIGW], Rm[BIGW];
    memset(A, 0, sizeof A); memset(B, 0, sizeof B); memset(P, 0, sizeof P);
    memset(prod, 0, sizeof prod); memset(Q, 0, sizeof Q); memset(Rm, 0, sizeof Rm);
    for (uint32_t i = 0; i < k; ++i) {
        big_add_shifted(A, &w[av[i]], n * i);
        big_add_shifted(B, &w[bv[i]], n * i);
        big_add_shifted(P, &w[pv[i]], n * i);
    }
    const int half = BIGW / 2;
    for (int i = 0; i < half; ++i) {
        u128 c = 0;
        for (int j = 0; j < half; ++j) { c += (u128)A[i] * B[j] + prod[i + j]; prod[i + j] = (uint64_t)c; c >>= 64; }
        if (i + half < BIGW) prod[i + half] += (uint64_t)c;
    }
    int pzero = 1;
    for (int i = 0; i < BIGW; ++i) if (P[i]) pzero = 0;
    if (!pzero) {
        for (int bit = 64 * BIGW - 1; bit >= 0; --bit) {   /* restoring shift-subtract division */
            uint64_t carry = (prod[bit >> 6] >> (bit & 63)) & 1;
            for (int i = 0; i < BIGW; ++i) { uint64_t nc = Rm[i] >> 63; Rm[i] = (Rm[i] << 1) | carry; carry = nc; }
            if (big_ge(Rm, P, BIGW)) { big_sub(Rm, P, BIGW); Q[bit >> 6] |= 1ull << (bit & 63); }
        }
    }
    for (uint32_t i = 0; i < k; ++i) {
        big_extract(Q, n * i, n, &w[dst + i]);
        big_extract(Rm, n * i, n, &w[dst + k + i]);
    }
}

int zkref_witness(const zkref_circuit* C, const uint8_t* inputs, uint8_t* w_bytes) {
    zkref_init();
    fe* w = (fe*)w_bytes;
    const uint32_t total = C->n_vars + C->n_temps;
    memset(w, 0, (size_t)total * 32);
    w[0].v[0] = 1;
    memcpy(&w[1 + C->n_outputs], inputs, (size_t)C->n_inputs * 32);
    for (uint32_t i = 0; i < C->n_inputs; ++i)
        if (fe_cmp(&w[1 + C->n_outputs + i], &ZK_FR.p) >= 0) return -2;   /* input not reduced */
    fe* coef_r = (fe*)malloc(sizeof(fe) * C->n_coefs);
    for (uint32_t i = 0; i < C->n_coefs; ++i) f_to_mont(&ZK_FR, &coef_r[i], (const fe*)(C->coefs + 32 * (size_t)i));
    for (uint32_t i = 0; i < C->n_ops; ++i) {
        const uint32_t* op = C->ops + 5 * (size_t)i;
        const uint32_t code = op[0], dst = op[1], a = op[2], b = op[3], c = op[4];
        switch (code) {
            case 0: /* OP_LIN */
                eval_lc(coef_r, C->lc_ptr, C->lc_var, C->lc_coef, a, w, &w[dst]);
                break;
            case 1: { /* OP_QUAD: dst = A*B + C */
                fe x, y, z, t;
                eval_lc(coef_r, C->lc_ptr, C->lc_var, C->lc_coef, a, w, &x);
                eval_lc(coef_r, C->lc_ptr, C->lc_var, C->lc_coef, b, w, &y);
                eval_lc(coef_r, C->lc_ptr, C->lc_var, C->lc_coef, c, w, &z);
                f_mul(&ZK_FR, &t, &x, &y);
                f_mul(&ZK_FR, &t, &t, &ZK_FR.r2);
                f_add(&ZK_FR, &w[dst], &t, &z);
                break;
            }
            case 2:   /* OP_SHRAND: (in >> b) & (2^c - 1) - circomlib Num2Bits `out[i] <-- (in >> i) & 1`,
                         lib/sha.circom:111 `inBlockIndex <-- (paddedInLength >> 9)` */
            case 5: { /* OP_SHRLC: the same hint applied to a linear combination (the front-end's fusion of the
                         preceding partial-sum op, e.g. circomlib BinSum `out[k] <-- (lin >> k) & 1`) */
                fe x, o = {{0, 0, 0, 0}};
                if (code == 5) eval_lc(coef_r, C->lc_ptr, C->lc_var, C->lc_coef, a, w, &x);
                else x = w[a];
                unsigned ws = b >> 6, bs = b & 63;
                for (unsigned j = 0; j + ws < 4; ++j) {
                    o.v[j] = x.v[j + ws] >> bs;
                    if (bs && j + ws + 1 < 4) o.v[j] |= x.v[j + ws + 1] << (64 - bs);
                }
                if (b >= 256) memset(&o, 0, sizeof o);
                if (c && c < 256) {
                    for (unsigned j = 0; j < 4; ++j) {
                        if (64 * j >= c) o.v[j] = 0;
                        else if (64 * (j + 1) > c) o.v[j] &= (1ull << (c - 64 * j)) - 1;
                    }
                }
                w[dst] = o;
                break;
            }
            case 3: { /* OP_INVZ - circomlib IsZero `inv <-- in!=0 ? 1/in : 0` */
                fe m, iv;
                f_to_mont(&ZK_FR, &m, &w[a]);
                f_inv(&ZK_FR, &iv, &m);
                f_from_mont(&ZK_FR, &w[dst], &iv);
                break;
            }
            case 4: { /* OP_FPMUL */
                const uint32_t* ax = C->aux + a;
                uint32_t n = ax[0], k = ax[1];
                if (n > 128 || (uint64_t)n * k + 256 > 64 * (BIGW / 2)) { free(coef_r); return -3; }
                fpmul_hint(n, k, ax + 2, ax + 2 + k, ax + 2 + 2 * k, w, dst);
                break;
            }
            default:
                free(coef_r);
                return -1;
        }
    }
    free(coef_r);
    return 0;
}

int64_t zkref_check_r1cs(const zkref_circuit* C, const uint8_t* w_bytes) {
    zkref_init();
    const fe* w = (const fe*)w_bytes;
    fe* coef_r = (fe*)malloc(sizeof(fe) * C->n_coefs);
    for (uint32_t i = 0; i < C->n_coefs; ++i) f_to_mont(&ZK_FR, &coef_r[i], (const fe*)(C->coefs + 32 * (size_t)i));
    int64_t bad = -1;
    for (uint32_t i = 0; i < C->n_constraints; ++i) {
        fe a, b, c, t;
        eval_lc(coef_r, C->a_ptr, C->a_var, C->a_coef, i, w, &a);
        eval_lc(coef_r, C->b_ptr, C->b_var, C->b_coef, i, w, &b);
        eval_lc(coef_r, C->c_ptr, C->c_var, C->c_coef, i, w, &c);
        f_mul(&ZK_FR, &t, &a, &b);
        f_mul(&ZK_FR, &t, &t, &ZK_FR.r2);
        if (!fe_eq(&t, &c)) { bad = i; break; }
    }
    free(coef_r);
    return bad;
}
