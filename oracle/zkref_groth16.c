/*
 * zkref_groth16.c - CPU oracle: Groth16 setup and prover over BN254 (plain C + pthreads).
 *
 * TEST INFRASTRUCTURE (see oracle/__init__.py); also the timed "port" CPU baseline of bench.py.
 *
 * Restates the algorithm of snarkjs 0.5.0 `groth16.prove` / `zkey new` (un-vendored fork pinned at
 * /root/reference/packages/helpers/package.json:26; call site /root/reference/packages/helpers/src/chunked-zkey.ts:80-84)
 * as recorded in SURVEY.md 3.2 and appendix A.7:
 *   1. a_i = <A_i, w>, b_i = <B_i, w> over the R1CS rows plus the nPublic+1 extra rows a = w_j; c_i = a_i b_i
 *   2. three inverse FFTs, multiplication of coefficient j by g^j (g = primitive 2N-th root), three forward FFTs
 *   3. d_i = a'_i b'_i - c'_i
 *   4. pi_A = alpha + sum w_j A_j + r delta ; pi_B = beta + sum w_j B_j + s delta (G2, and its G1 twin) ;
 *      pi_C = sum_{j > nPublic} w_j C_j + sum d_i H_i + s pi_A + r pi_B1 - r s delta
 * Multi-exponentiations use the bucket method with signed windows (the role of wasmcurves' multiExpAffine).
 * Parity status: prover parity is unpinned in the reference (no proof-producing test or fixture exists); proofs
 * produced here are checked against the fixture-pinned verifier of oracle/bn254.py.
 */
#include "zkref.h"
#include <pthread.h>
#include <stdlib.h>

/* ------------------------------------------------------------------------------------------------ Fq2 */
typedef struct { fe c0, c1; } fe2;
static inline void f2_add(fe2* r, const fe2* a, const fe2* b) { f_add(&ZK_FQ, &r->c0, &a->c0, &b->c0); f_add(&ZK_FQ, &r->c1, &a->c1, &b->c1); }
static inline void f2_sub(fe2* r, const fe2* a, const fe2* b) { f_sub(&ZK_FQ, &r->c0, &a->c0, &b->c0); f_sub(&ZK_FQ, &r->c1, &a->c1, &b->c1); }
static inline void f2_mul(fe2* r, const fe2* a, const fe2* b) {
    fe t0, t1, s0, s1, m;
    f_mul(&ZK_FQ, &t0, &a->c0, &b->c0);
    f_mul(&ZK_FQ, &t1, &a->c1, &b->c1);
    f_add(&ZK_FQ, &s0, &a->c0, &a->c1);
    f_add(&ZK_FQ, &s1, &b->c0, &b->c1);
    f_mul(&ZK_FQ, &m, &s0, &s1);
    f_sub(&ZK_FQ, &m, &m, &t0);
    f_sub(&ZK_FQ, &r->c1, &m, &t1);
    f_sub(&ZK_FQ, &r->c0, &t0, &t1);
}
static inline void f2_sqr(fe2* r, const fe2* a) { f2_mul(r, a, a); }
static inline int f2_is_zero(const fe2* a) { return fe_is_zero(&a->c0) && fe_is_zero(&a->c1); }
static inline int f2_eq(const fe2* a, const fe2* b) { return fe_eq(&a->c0, &b->c0) && fe_eq(&a->c1, &b->c1); }
static void f2_inv(fe2* r, const fe2* a) {
    fe t0, t1, d;
    f_sqr(&ZK_FQ, &t0, &a->c0); f_sqr(&ZK_FQ, &t1, &a->c1);
    f_add(&ZK_FQ, &d, &t0, &t1);
    f_inv(&ZK_FQ, &d, &d);
    f_mul(&ZK_FQ, &r->c0, &a->c0, &d);
    f_mul(&ZK_FQ, &t0, &a->c1, &d);
    f_neg(&ZK_FQ, &r->c1, &t0);
}

/* ------------------------------------------------------------------------------------------------ curves
 * The same Jacobian formulas instantiated for G1 (coordinates in Fq) and G2 (Fq2) through a tiny macro layer. */
#define DEFINE_CURVE(PFX, T, ADD, SUB, MUL, SQR, ISZ, EQ)                                                         \
    typedef struct { T x, y; } PFX##_aff;          /* (0,0) = infinity */                                         \
    typedef struct { T x, y, z; } PFX##_jac;       /* z = 0: infinity */                                          \
    static inline int PFX##_aff_inf(const PFX##_aff* p) { return ISZ(&p->x) && ISZ(&p->y); }                      \
    static inline void PFX##_set_inf(PFX##_jac* p) { memset(p, 0, sizeof *p); }                                   \
    static void PFX##_dbl(PFX##_jac* r, const PFX##_jac* p) {                                                     \
        if (ISZ(&p->z)) { *r = *p; return; }                                                                      \
        T A, B, C, D, E, F, t, X3, Y3, Z3;                                                                        \
        SQR(&A, &p->x); SQR(&B, &p->y); SQR(&C, &B);                                                              \
        ADD(&t, &p->x, &B); SQR(&t, &t); SUB(&t, &t, &A); SUB(&t, &t, &C); ADD(&D, &t, &t);                       \
        ADD(&E, &A, &A); ADD(&E, &E, &A); SQR(&F, &E);                                                            \
        SUB(&X3, &F, &D); SUB(&X3, &X3, &D);                                                                      \
        SUB(&t, &D, &X3); MUL(&Y3, &E, &t);                                                                       \
        ADD(&t, &C, &C); ADD(&t, &t, &t); ADD(&t, &t, &t); SUB(&Y3, &Y3, &t);                                     \
        MUL(&Z3, &p->y, &p->z); ADD(&Z3, &Z3, &Z3);                                                               \
        r->x = X3; r->y = Y3; r->z = Z3;                                                                          \
    }                                                                                                             \
    /* r = p + q (q affine, optionally negated) */                                                                \
    static void PFX##_madd(PFX##_jac* r, const PFX##_jac* p, const PFX##_aff* q, int negate, const T* one) {      \
        if (PFX##_aff_inf(q)) { *r = *p; return; }                                                                \
        T qy = q->y;                                                                                              \
        if (negate) { T z0; memset(&z0, 0, sizeof z0); SUB(&qy, &z0, &q->y); }                                    \
        if (ISZ(&p->z)) { r->x = q->x; r->y = qy; r->z = *one; return; }                                          \
        T Z1Z1, U2, S2, H, HH, I, J, rr, V, t, X3, Y3, Z3;                                                        \
        SQR(&Z1Z1, &p->z); MUL(&U2, &q->x, &Z1Z1); MUL(&S2, &qy, &p->z); MUL(&S2, &S2, &Z1Z1);                    \
        SUB(&H, &U2, &p->x); SUB(&rr, &S2, &p->y);                                                                \
        if (ISZ(&H)) { if (ISZ(&rr)) { PFX##_dbl(r, p); return; } PFX##_set_inf(r); return; }                     \
        SQR(&HH, &H); MUL(&J, &H, &HH); MUL(&V, &p->x, &HH);                                                      \
        SQR(&X3, &rr); SUB(&X3, &X3, &J); SUB(&X3, &X3, &V); SUB(&X3, &X3, &V);                                   \
        SUB(&t, &V, &X3); MUL(&Y3, &rr, &t); MUL(&t, &p->y, &J); SUB(&Y3, &Y3, &t);                               \
        MUL(&Z3, &p->z, &H);                                                                                      \
        (void)I; r->x = X3; r->y = Y3; r->z = Z3;                                                                 \
    }                                                                                                             \
    static void PFX##_add(PFX##_jac* r, const PFX##_jac* p, const PFX##_jac* q) {                                 \
        if (ISZ(&p->z)) { *r = *q; return; }                                                                      \
        if (ISZ(&q->z)) { *r = *p; return; }                                                                      \
        T Z1Z1, Z2Z2, U1, U2, S1, S2, H, HH, J, rr, V, t, X3, Y3, Z3;                                             \
        SQR(&Z1Z1, &p->z); SQR(&Z2Z2, &q->z);                                                                     \
        MUL(&U1, &p->x, &Z2Z2); MUL(&U2, &q->x, &Z1Z1);                                                           \
        MUL(&S1, &p->y, &q->z); MUL(&S1, &S1, &Z2Z2); MUL(&S2, &q->y, &p->z); MUL(&S2, &S2, &Z1Z1);               \
        SUB(&H, &U2, &U1); SUB(&rr, &S2, &S1);                                                                    \
        if (ISZ(&H)) { if (ISZ(&rr)) { PFX##_dbl(r, p); return; } PFX##_set_inf(r); return; }                     \
        SQR(&HH, &H); MUL(&J, &H, &HH); MUL(&V, &U1, &HH);                                                        \
        SQR(&X3, &rr); SUB(&X3, &X3, &J); SUB(&X3, &X3, &V); SUB(&X3, &X3, &V);                                   \
        SUB(&t, &V, &X3); MUL(&Y3, &rr, &t); MUL(&t, &S1, &J); SUB(&Y3, &Y3, &t);                                 \
        MUL(&Z3, &p->z, &q->z); MUL(&Z3, &Z3, &H);                                                                \
        r->x = X3; r->y = Y3; r->z = Z3;                                                                          \
    }

#define FQ_ADD(r, a, b) f_add(&ZK_FQ, r, a, b)
#define FQ_SUB(r, a, b) f_sub(&ZK_FQ, r, a, b)
#define FQ_MUL(r, a, b) f_mul(&ZK_FQ, r, a, b)
#define FQ_SQR(r, a) f_sqr(&ZK_FQ, r, a)
DEFINE_CURVE(g1, fe, FQ_ADD, FQ_SUB, FQ_MUL, FQ_SQR, fe_is_zero, fe_eq)
DEFINE_CURVE(g2, fe2, f2_add, f2_sub, f2_mul, f2_sqr, f2_is_zero, f2_eq)

static fe FQ_ONE_M;       /* Montgomery one of Fq */
static fe2 FQ2_ONE_M;
static void curve_init(void) { zkref_init(); FQ_ONE_M = ZK_FQ.r; memset(&FQ2_ONE_M, 0, sizeof FQ2_ONE_M); FQ2_ONE_M.c0 = ZK_FQ.r; }

static void g1_to_affine(g1_aff* r, const g1_jac* p) {
    if (fe_is_zero(&p->z)) { memset(r, 0, sizeof *r); return; }
    fe zi, zi2, zi3;
    f_inv(&ZK_FQ, &zi, &p->z); f_sqr(&ZK_FQ, &zi2, &zi); f_mul(&ZK_FQ, &zi3, &zi2, &zi);
    f_mul(&ZK_FQ, &r->x, &p->x, &zi2); f_mul(&ZK_FQ, &r->y, &p->y, &zi3);
}
static void g2_to_affine(g2_aff* r, const g2_jac* p) {
    if (f2_is_zero(&p->z)) { memset(r, 0, sizeof *r); return; }
    fe2 zi, zi2, zi3;
    f2_inv(&zi, &p->z); f2_sqr(&zi2, &zi); f2_mul(&zi3, &zi2, &zi);
    f2_mul(&r->x, &p->x, &zi2); f2_mul(&r->y, &p->y, &zi3);
}
static void g1_from_aff(g1_jac* r, const g1_aff* a) { if (g1_aff_inf(a)) { g1_set_inf(r); return; } r->x = a->x; r->y = a->y; r->z = FQ_ONE_M; }
static void g2_from_aff(g2_jac* r, const g2_aff* a) { if (g2_aff_inf(a)) { g2_set_inf(r); return; } r->x = a->x; r->y = a->y; r->z = FQ2_ONE_M; }
static void g1_neg_jac(g1_jac* p) { f_neg(&ZK_FQ, &p->y, &p->y); }

static void g1_scalar_mul(g1_jac* r, const g1_jac* p, const fe* k /* standard form */) {
    g1_jac acc; g1_set_inf(&acc);
    for (int i = 255; i >= 0; --i) {
        g1_dbl(&acc, &acc);
        if ((k->v[i >> 6] >> (i & 63)) & 1) g1_add(&acc, &acc, p);
    }
    *r = acc;
}
static void g2_scalar_mul(g2_jac* r, const g2_jac* p, const fe* k) {
    g2_jac acc; g2_set_inf(&acc);
    for (int i = 255; i >= 0; --i) {
        g2_dbl(&acc, &acc);
        if ((k->v[i >> 6] >> (i & 63)) & 1) g2_add(&acc, &acc, p);
    }
    *r = acc;
}

/* standard-form little-endian byte images <-> Montgomery points */
static void g1_load_std(g1_aff* r, const uint8_t* b) {
    fe x, y; memcpy(&x, b, 32); memcpy(&y, b + 32, 32);
    f_to_mont(&ZK_FQ, &r->x, &x); f_to_mont(&ZK_FQ, &r->y, &y);
}
static void g2_load_std(g2_aff* r, const uint8_t* b) {
    fe t[4]; memcpy(t, b, 128);
    f_to_mont(&ZK_FQ, &r->x.c0, &t[0]); f_to_mont(&ZK_FQ, &r->x.c1, &t[1]);
    f_to_mont(&ZK_FQ, &r->y.c0, &t[2]); f_to_mont(&ZK_FQ, &r->y.c1, &t[3]);
}
static void g1_store_std(uint8_t* b, const g1_aff* p) {
    fe x, y; f_from_mont(&ZK_FQ, &x, &p->x); f_from_mont(&ZK_FQ, &y, &p->y);
    memcpy(b, &x, 32); memcpy(b + 32, &y, 32);
}
static void g2_store_std(uint8_t* b, const g2_aff* p) {
    fe t[4];
    f_from_mont(&ZK_FQ, &t[0], &p->x.c0); f_from_mont(&ZK_FQ, &t[1], &p->x.c1);
    f_from_mont(&ZK_FQ, &t[2], &p->y.c0); f_from_mont(&ZK_FQ, &t[3], &p->y.c1);
    memcpy(b, t, 128);
}

/* ------------------------------------------------------------------------------------------------ MSM
 * Bucket method, signed c-bit windows; a task = (slice of the points, window); tasks are spread over threads. */
#define MSM_C 13
#define MSM_W ((255 + MSM_C - 1) / MSM_C)
#define MSM_HALF (1u << (MSM_C - 1))

typedef struct {
    int is_g2;
    const uint8_t* points_std;   /* n affine points, standard form (64 or 128 bytes each) */
    const fe* scalars;           /* n standard-form scalars */
    size_t n;
    int parts;
    g1_jac* out1;                /* [parts][MSM_W] window sums */
    g2_jac* out2;
    int next_task;
    pthread_mutex_t mu;
} msm_job;

static inline int32_t msm_digit(const fe* s, int window, int* carry_inout) {
    int bit = window * MSM_C;
    int w = bit >> 6, b = bit & 63;
    uint64_t v = s->v[w] >> b;
    if (b + MSM_C > 64 && w + 1 < 4) v |= s->v[w + 1] << (64 - b);
    uint32_t raw = (uint32_t)(v & ((1u << MSM_C) - 1)) + (uint32_t)*carry_inout;
    if (raw > MSM_HALF) { *carry_inout = 1; return (int32_t)raw - (1 << MSM_C); }
    *carry_inout = 0;
    return (int32_t)raw;
}

static void* msm_worker(void* arg) {
    msm_job* J = (msm_job*)arg;
    const int total = J->parts * MSM_W;
    g1_jac* b1 = NULL; g2_jac* b2 = NULL;
    if (J->is_g2) b2 = (g2_jac*)malloc(sizeof(g2_jac) * MSM_HALF); else b1 = (g1_jac*)malloc(sizeof(g1_jac) * MSM_HALF);
    for (;;) {
        pthread_mutex_lock(&J->mu);
        int task = J->next_task++;
        pthread_mutex_unlock(&J->mu);
        if (task >= total) break;
        const int part = task / MSM_W, window = task % MSM_W;
        const size_t beg = J->n * part / J->parts, end = J->n * (part + 1) / J->parts;
        if (J->is_g2) memset(b2, 0, sizeof(g2_jac) * MSM_HALF); else memset(b1, 0, sizeof(g1_jac) * MSM_HALF);
        for (size_t i = beg; i < end; ++i) {
            const fe* s = &J->scalars[i];
            if (fe_is_zero(s)) continue;
            /* digit of this window needs the carries of all lower windows */
            int carry = 0; int32_t d = 0;
            for (int w = 0; w <= window; ++w) d = msm_digit(s, w, &carry);
            if (d == 0) continue;
            uint32_t idx = (uint32_t)(d < 0 ? -d : d) - 1;
            if (J->is_g2) { g2_aff p; g2_load_std(&p, J->points_std + 128 * i); g2_madd(&b2[idx], &b2[idx], &p, d < 0, &FQ2_ONE_M); }
            else { g1_aff p; g1_load_std(&p, J->points_std + 64 * i); g1_madd(&b1[idx], &b1[idx], &p, d < 0, &FQ_ONE_M); }
        }
        if (J->is_g2) {
            g2_jac run, sum; g2_set_inf(&run); g2_set_inf(&sum);
            for (int k = MSM_HALF - 1; k >= 0; --k) { g2_add(&run, &run, &b2[k]); g2_add(&sum, &sum, &run); }
            J->out2[task] = sum;
        } else {
            g1_jac run, sum; g1_set_inf(&run); g1_set_inf(&sum);
            for (int k = MSM_HALF - 1; k >= 0; --k) { g1_add(&run, &run, &b1[k]); g1_add(&sum, &sum, &run); }
            J->out1[task] = sum;
        }
    }
    free(b1); free(b2);
    return NULL;
}

static void msm_run(int is_g2, const uint8_t* points_std, const fe* scalars, size_t n, int threads, g1_jac* r1, g2_jac* r2) {
    msm_job J;
    memset(&J, 0, sizeof J);
    J.is_g2 = is_g2; J.points_std = points_std; J.scalars = scalars; J.n = n;
    J.parts = (threads + MSM_W - 1) / MSM_W;
    if (J.parts < 1) J.parts = 1;
    if (n < 4096) J.parts = 1;
    const int total = J.parts * MSM_W;
    if (is_g2) J.out2 = (g2_jac*)calloc(total, sizeof(g2_jac)); else J.out1 = (g1_jac*)calloc(total, sizeof(g1_jac));
    pthread_mutex_init(&J.mu, NULL);
    int nt = threads < 1 ? 1 : threads;
    if (nt > total) nt = total;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nt);
    for (int t = 0; t < nt; ++t) pthread_create(&th[t], NULL, msm_worker, &J);
    for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
    free(th);
    pthread_mutex_destroy(&J.mu);
    /* combine: sum over parts per window, then Horner over windows */
    if (is_g2) {
        g2_jac acc; g2_set_inf(&acc);
        for (int w = MSM_W - 1; w >= 0; --w) {
            for (int k = 0; k < MSM_C; ++k) g2_dbl(&acc, &acc);
            for (int p = 0; p < J.parts; ++p) g2_add(&acc, &acc, &J.out2[p * MSM_W + w]);
        }
        *r2 = acc; free(J.out2);
    } else {
        g1_jac acc; g1_set_inf(&acc);
        for (int w = MSM_W - 1; w >= 0; --w) {
            for (int k = 0; k < MSM_C; ++k) g1_dbl(&acc, &acc);
            for (int p = 0; p < J.parts; ++p) g1_add(&acc, &acc, &J.out1[p * MSM_W + w]);
        }
        *r1 = acc; free(J.out1);
    }
}

/* ------------------------------------------------------------------------------------------------ FFT */
static void fr_root(fe* w, unsigned log_n) {   /* primitive 2^log_n-th root (generator 5), Montgomery */
    fe five = {{5, 0, 0, 0}}, g, e, one = {{1, 0, 0, 0}};
    f_to_mont(&ZK_FR, &g, &five);
    fe_sub_raw(&e, &ZK_FR.p, &one);
    for (int i = 0; i < 28; ++i) {   /* e = (r-1) >> 28 */
        for (int j = 0; j < 4; ++j) e.v[j] = (e.v[j] >> 1) | (j < 3 ? e.v[j + 1] << 63 : 0);
    }
    f_pow(&ZK_FR, w, &g, &e);
    for (unsigned i = log_n; i < 28; ++i) f_sqr(&ZK_FR, w, w);
}

/* in-place radix-2 FFT of Montgomery values; natural order in and out */
static void fft(fe* a, unsigned log_n, const fe* root) {
    const size_t n = (size_t)1 << log_n;
    for (size_t i = 0, j = 0; i < n; ++i) {
        if (i < j) { fe t = a[i]; a[i] = a[j]; a[j] = t; }
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
    }
    fe* tw = (fe*)malloc(sizeof(fe) * (n / 2 ? n / 2 : 1));
    tw[0] = ZK_FR.r;
    for (size_t i = 1; i < n / 2; ++i) f_mul(&ZK_FR, &tw[i], &tw[i - 1], root);
    for (size_t len = 2; len <= n; len <<= 1) {
        const size_t half = len >> 1, step = n / len;
        for (size_t i = 0; i < n; i += len) {
            for (size_t j = 0; j < half; ++j) {
                fe t, u = a[i + j];
                f_mul(&ZK_FR, &t, &a[i + j + half], &tw[j * step]);
                f_add(&ZK_FR, &a[i + j], &u, &t);
                f_sub(&ZK_FR, &a[i + j + half], &u, &t);
            }
        }
    }
    free(tw);
}

typedef struct { fe* v; unsigned log_n; } coset_job;
/* evaluations on H -> evaluations on the coset g H */
static void* coset_worker(void* arg) {
    coset_job* J = (coset_job*)arg;
    const size_t n = (size_t)1 << J->log_n;
    fe w, wi, g, acc, ninv, nn = {{(uint64_t)n, 0, 0, 0}};
    fr_root(&w, J->log_n);
    f_inv(&ZK_FR, &wi, &w);
    fr_root(&g, J->log_n + 1);
    f_to_mont(&ZK_FR, &ninv, &nn);
    f_inv(&ZK_FR, &ninv, &ninv);
    fft(J->v, J->log_n, &wi);
    acc = ninv;
    for (size_t j = 0; j < n; ++j) { f_mul(&ZK_FR, &J->v[j], &J->v[j], &acc); f_mul(&ZK_FR, &acc, &acc, &g); }
    fft(J->v, J->log_n, &w);
    return NULL;
}

/* ------------------------------------------------------------------------------------------------ prove */
static void lc_dot(const fe* coef_r, const uint32_t* ptr, const uint32_t* var, const uint32_t* coef, uint32_t row, const fe* w, fe* out) {
    fe acc = {{0, 0, 0, 0}};
    for (uint32_t k = ptr[row]; k < ptr[row + 1]; ++k) {
        fe t; f_mul(&ZK_FR, &t, &coef_r[coef[k]], &w[var[k]]); f_add(&ZK_FR, &acc, &acc, &t);
    }
    *out = acc;   /* standard form */
}

int zkref_groth16_prove(const zkref_circuit* C, const zkref_zkey* K, const uint8_t* w_bytes, const uint8_t* r32, const uint8_t* s32,
                        int threads, uint8_t* proof256) {
    curve_init();
    const fe* w = (const fe*)w_bytes;
    const unsigned log_n = K->log_n;
    const size_t N = (size_t)1 << log_n;
    const uint32_t m = K->n_vars, l = K->n_public;
    if (C->n_constraints + l + 1 > N) return -1;
    fe* coef_r = (fe*)malloc(sizeof(fe) * C->n_coefs);
    for (uint32_t i = 0; i < C->n_coefs; ++i) f_to_mont(&ZK_FR, &coef_r[i], (const fe*)(C->coefs + 32 * (size_t)i));
    fe* a = (fe*)calloc(N, sizeof(fe)); fe* b = (fe*)calloc(N, sizeof(fe)); fe* c = (fe*)calloc(N, sizeof(fe));
    for (uint32_t i = 0; i < C->n_constraints; ++i) {
        fe x, y;
        lc_dot(coef_r, C->a_ptr, C->a_var, C->a_coef, i, w, &x);
        lc_dot(coef_r, C->b_ptr, C->b_var, C->b_coef, i, w, &y);
        f_to_mont(&ZK_FR, &a[i], &x); f_to_mont(&ZK_FR, &b[i], &y);
    }
    for (uint32_t j = 0; j <= l; ++j) f_to_mont(&ZK_FR, &a[C->n_constraints + j], &w[j]);
    for (size_t i = 0; i < N; ++i) f_mul(&ZK_FR, &c[i], &a[i], &b[i]);
    coset_job jobs[3] = {{a, log_n}, {b, log_n}, {c, log_n}};
    if (threads >= 3) {
        pthread_t th[3];
        for (int t = 0; t < 3; ++t) pthread_create(&th[t], NULL, coset_worker, &jobs[t]);
        for (int t = 0; t < 3; ++t) pthread_join(th[t], NULL);
    } else for (int t = 0; t < 3; ++t) coset_worker(&jobs[t]);
    fe* d = (fe*)malloc(sizeof(fe) * N);
    for (size_t i = 0; i < N; ++i) { fe t; f_mul(&ZK_FR, &t, &a[i], &b[i]); f_sub(&ZK_FR, &t, &t, &c[i]); f_from_mont(&ZK_FR, &d[i], &t); }
    free(a); free(b); free(c); free(coef_r);

    g1_jac mA, mB1, mC, mH; g2_jac mB2;
    msm_run(0, K->A, w, m, threads, &mA, NULL);
    msm_run(0, K->B1, w, m, threads, &mB1, NULL);
    msm_run(0, K->C, w, m, threads, &mC, NULL);     /* C is infinity for the public signals */
    msm_run(0, K->H, d, N, threads, &mH, NULL);
    msm_run(1, K->B2, w, m, threads, NULL, &mB2);
    free(d);

    fe r, s, rm, sm, rs;
    memcpy(&r, r32, 32); memcpy(&s, s32, 32);
    f_to_mont(&ZK_FR, &rm, &r); f_to_mont(&ZK_FR, &sm, &s);
    f_mul(&ZK_FR, &rs, &rm, &sm); f_from_mont(&ZK_FR, &rs, &rs);
    g1_aff alpha1, beta1, delta1; g2_aff beta2, delta2;
    g1_load_std(&alpha1, K->alpha1); g1_load_std(&beta1, K->beta1); g1_load_std(&delta1, K->delta1);
    g2_load_std(&beta2, K->beta2); g2_load_std(&delta2, K->delta2);
    g1_jac d1, t1, piA, piB1, piC; g2_jac d2, t2, piB2;
    g1_from_aff(&d1, &delta1); g2_from_aff(&d2, &delta2);
    /* pi_A = alpha + A + r delta */
    g1_scalar_mul(&t1, &d1, &r); g1_madd(&piA, &mA, &alpha1, 0, &FQ_ONE_M); g1_add(&piA, &piA, &t1);
    /* pi_B = beta + B + s delta (G2 and G1) */
    g2_scalar_mul(&t2, &d2, &s); g2_madd(&piB2, &mB2, &beta2, 0, &FQ2_ONE_M); g2_add(&piB2, &piB2, &t2);
    g1_scalar_mul(&t1, &d1, &s); g1_madd(&piB1, &mB1, &beta1, 0, &FQ_ONE_M); g1_add(&piB1, &piB1, &t1);
    /* pi_C = C + H + s pi_A + r pi_B1 - r s delta */
    g1_add(&piC, &mC, &mH);
    g1_scalar_mul(&t1, &piA, &s); g1_add(&piC, &piC, &t1);
    g1_scalar_mul(&t1, &piB1, &r); g1_add(&piC, &piC, &t1);
    g1_scalar_mul(&t1, &d1, &rs); g1_neg_jac(&t1); g1_add(&piC, &piC, &t1);
    g1_aff A, Cc; g2_aff B;
    g1_to_affine(&A, &piA); g1_to_affine(&Cc, &piC); g2_to_affine(&B, &piB2);
    g1_store_std(proof256, &A); g2_store_std(proof256 + 64, &B); g1_store_std(proof256 + 192, &Cc);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ setup
 * Independent restatement of the key derivation for SMALL circuits (direct formulas, one inversion per Lagrange
 * value, double-and-add for every point).  toxic = tau, alpha, beta, gamma, delta (standard form, 32 bytes each).
 * Outputs use the same standard-form affine images as the product's zke_zkey_section(). */
static void lagrange_at(fe* out, const fe* x, const fe* numer, unsigned log_n) {   /* out[i] = numer w^i / (x - w^i) */
    const size_t n = (size_t)1 << log_n;
    fe w, wi = ZK_FR.r;
    fr_root(&w, log_n);
    for (size_t i = 0; i < n; ++i) {
        fe den; f_sub(&ZK_FR, &den, x, &wi); f_inv(&ZK_FR, &den, &den);
        f_mul(&ZK_FR, &out[i], &den, &wi); f_mul(&ZK_FR, &out[i], &out[i], numer);
        f_mul(&ZK_FR, &wi, &wi, &w);
    }
}

int zkref_groth16_setup(const zkref_circuit* C, unsigned log_n, uint32_t n_public, const uint8_t* toxic,
                        uint8_t* A, uint8_t* B1, uint8_t* B2, uint8_t* Cs, uint8_t* H, uint8_t* IC,
                        uint8_t* alpha1, uint8_t* beta1, uint8_t* delta1, uint8_t* beta2, uint8_t* gamma2, uint8_t* delta2) {
    curve_init();
    const size_t N = (size_t)1 << log_n;
    const uint32_t m = C->n_vars, l = n_public;
    fe tx[5], tm[5];
    memcpy(tx, toxic, 160);
    for (int i = 0; i < 5; ++i) f_to_mont(&ZK_FR, &tm[i], &tx[i]);
    const fe *tau = &tm[0], *alpha = &tm[1], *beta = &tm[2], *gamma = &tm[3], *delta = &tm[4];
    fe nn = {{(uint64_t)N, 0, 0, 0}}, ninv, tauN, ztau, numer;
    f_to_mont(&ZK_FR, &ninv, &nn); f_inv(&ZK_FR, &ninv, &ninv);
    f_pow(&ZK_FR, &tauN, tau, &nn);
    f_sub(&ZK_FR, &ztau, &tauN, &ZK_FR.r);
    f_mul(&ZK_FR, &numer, &ztau, &ninv);
    fe* lag = (fe*)malloc(sizeof(fe) * N);
    lagrange_at(lag, tau, &numer, log_n);
    fe* coef_m = (fe*)malloc(sizeof(fe) * C->n_coefs);
    for (uint32_t i = 0; i < C->n_coefs; ++i) f_to_mont(&ZK_FR, &coef_m[i], (const fe*)(C->coefs + 32 * (size_t)i));
    fe* av = (fe*)calloc(m, sizeof(fe)); fe* bv = (fe*)calloc(m, sizeof(fe)); fe* cv = (fe*)calloc(m, sizeof(fe));
    for (uint32_t row = 0; row < C->n_constraints; ++row) {
        for (uint32_t k = C->a_ptr[row]; k < C->a_ptr[row + 1]; ++k) { fe t; f_mul(&ZK_FR, &t, &coef_m[C->a_coef[k]], &lag[row]); f_add(&ZK_FR, &av[C->a_var[k]], &av[C->a_var[k]], &t); }
        for (uint32_t k = C->b_ptr[row]; k < C->b_ptr[row + 1]; ++k) { fe t; f_mul(&ZK_FR, &t, &coef_m[C->b_coef[k]], &lag[row]); f_add(&ZK_FR, &bv[C->b_var[k]], &bv[C->b_var[k]], &t); }
        for (uint32_t k = C->c_ptr[row]; k < C->c_ptr[row + 1]; ++k) { fe t; f_mul(&ZK_FR, &t, &coef_m[C->c_coef[k]], &lag[row]); f_add(&ZK_FR, &cv[C->c_var[k]], &cv[C->c_var[k]], &t); }
    }
    for (uint32_t j = 0; j <= l; ++j) f_add(&ZK_FR, &av[j], &av[j], &lag[C->n_constraints + j]);
    fe ginv, dinv;
    f_inv(&ZK_FR, &ginv, gamma); f_inv(&ZK_FR, &dinv, delta);
    g1_aff G1 = {{{1, 0, 0, 0}}, {{2, 0, 0, 0}}};
    f_to_mont(&ZK_FQ, &G1.x, &G1.x); f_to_mont(&ZK_FQ, &G1.y, &G1.y);
    static const uint64_t G2X0[4] = {0x46debd5cd992f6edull, 0x674322d4f75edaddull, 0x426a00665e5c4479ull, 0x1800deef121f1e76ull};
    static const uint64_t G2X1[4] = {0x97e485b7aef312c2ull, 0xf1aa493335a9e712ull, 0x7260bfb731fb5d25ull, 0x198e9393920d483aull};
    static const uint64_t G2Y0[4] = {0x4ce6cc0166fa7daaull, 0xe3d1e7690c43d37bull, 0x4aab71808dcb408full, 0x12c85ea5db8c6debull};
    static const uint64_t G2Y1[4] = {0x55acdadcd122975bull, 0xbc4b313370b38ef3ull, 0xec9e99ad690c3395ull, 0x090689d0585ff075ull};
    g2_aff G2;
    memcpy(&G2.x.c0, G2X0, 32); memcpy(&G2.x.c1, G2X1, 32); memcpy(&G2.y.c0, G2Y0, 32); memcpy(&G2.y.c1, G2Y1, 32);
    f_to_mont(&ZK_FQ, &G2.x.c0, &G2.x.c0); f_to_mont(&ZK_FQ, &G2.x.c1, &G2.x.c1);
    f_to_mont(&ZK_FQ, &G2.y.c0, &G2.y.c0); f_to_mont(&ZK_FQ, &G2.y.c1, &G2.y.c1);
    g1_jac g1j; g2_jac g2j;
    g1_from_aff(&g1j, &G1); g2_from_aff(&g2j, &G2);
#define EMIT_G1(dst, scalar_mont) do { fe s_; f_from_mont(&ZK_FR, &s_, scalar_mont); g1_jac p_; g1_scalar_mul(&p_, &g1j, &s_); g1_aff q_; g1_to_affine(&q_, &p_); g1_store_std(dst, &q_); } while (0)
#define EMIT_G2(dst, scalar_mont) do { fe s_; f_from_mont(&ZK_FR, &s_, scalar_mont); g2_jac p_; g2_scalar_mul(&p_, &g2j, &s_); g2_aff q_; g2_to_affine(&q_, &p_); g2_store_std(dst, &q_); } while (0)
    EMIT_G1(alpha1, alpha); EMIT_G1(beta1, beta); EMIT_G1(delta1, delta);
    EMIT_G2(beta2, beta); EMIT_G2(gamma2, gamma); EMIT_G2(delta2, delta);
    for (uint32_t j = 0; j < m; ++j) {
        EMIT_G1(A + 64 * (size_t)j, &av[j]);
        EMIT_G1(B1 + 64 * (size_t)j, &bv[j]);
        EMIT_G2(B2 + 128 * (size_t)j, &bv[j]);
        fe k1, k2, kc;
        f_mul(&ZK_FR, &k1, beta, &av[j]); f_mul(&ZK_FR, &k2, alpha, &bv[j]);
        f_add(&ZK_FR, &kc, &k1, &k2); f_add(&ZK_FR, &kc, &kc, &cv[j]);
        if (j <= l) { f_mul(&ZK_FR, &kc, &kc, &ginv); EMIT_G1(IC + 64 * (size_t)j, &kc); memset(Cs + 64 * (size_t)j, 0, 64); }
        else { f_mul(&ZK_FR, &kc, &kc, &dinv); EMIT_G1(Cs + 64 * (size_t)j, &kc); }
    }
    /* H_i = -L_i(tau/g) Z(tau) / (2 delta) with (tau/g)^N = -tau^N */
    fe g, gi, taug, zc, two = {{2, 0, 0, 0}}, twoinv, hn, zero = {{0, 0, 0, 0}};
    fr_root(&g, log_n + 1); f_inv(&ZK_FR, &gi, &g); f_mul(&ZK_FR, &taug, tau, &gi);
    f_sub(&ZK_FR, &zc, &zero, &tauN); f_sub(&ZK_FR, &zc, &zc, &ZK_FR.r);
    f_to_mont(&ZK_FR, &twoinv, &two); f_inv(&ZK_FR, &twoinv, &twoinv);
    f_mul(&ZK_FR, &hn, &zc, &ninv); f_mul(&ZK_FR, &hn, &hn, &ztau); f_mul(&ZK_FR, &hn, &hn, &twoinv); f_mul(&ZK_FR, &hn, &hn, &dinv);
    f_sub(&ZK_FR, &hn, &zero, &hn);
    lagrange_at(lag, &taug, &hn, log_n);
    for (size_t i = 0; i < N; ++i) EMIT_G1(H + 64 * i, &lag[i]);
    free(lag); free(coef_m); free(av); free(bv); free(cv);
    return 0;
}
