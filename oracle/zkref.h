/*
 * zkref.h - CPU oracle entry points (plain C).  TEST INFRASTRUCTURE - see oracle/__init__.py.
 * The circuit is passed in as flat read-only arrays obtained from the product's zke_circuit_array().
 */
#ifndef ZKREF_H
#define ZKREF_H
#include "zkref_field.h"
#include <stddef.h>

typedef struct {
    uint32_t n_vars, n_temps, n_outputs, n_inputs, n_constraints, n_ops, n_coefs, pad_;
    const uint8_t* coefs;                     /* [n_coefs][32] standard form LE */
    const uint32_t *a_ptr, *a_var, *a_coef;
    const uint32_t *b_ptr, *b_var, *b_coef;
    const uint32_t *c_ptr, *c_var, *c_coef;
    const uint32_t* ops;                      /* [n_ops][5] */
    const uint32_t *lc_ptr, *lc_var, *lc_coef;
    const uint32_t* aux;
} zkref_circuit;

/* calculateWitness: inputs [n_inputs][32] LE -> w [(n_vars + n_temps)][32] LE (scratch slots follow the witness) */
int zkref_witness(const zkref_circuit* C, const uint8_t* inputs, uint8_t* w);
/* checkConstraints: index of the first violated R1CS row, or -1 */
int64_t zkref_check_r1cs(const zkref_circuit* C, const uint8_t* w);


/* Proving key as flat standard-form affine point arrays (the images returned by the product's zke_zkey_section). */
typedef struct {
    uint32_t n_vars, n_public, log_n, pad_;
    const uint8_t *alpha1, *beta1, *delta1;   /* G1, 64 bytes each */
    const uint8_t *beta2, *delta2;            /* G2, 128 bytes each */
    const uint8_t *A, *B1, *C, *H;            /* G1 arrays: n_vars, n_vars, n_vars, 2^log_n points */
    const uint8_t* B2;                        /* G2 array: n_vars points */
} zkref_zkey;

/* snarkjs groth16.prove: witness w [n_vars][32], blinding r, s (32 bytes LE each) -> proof
 * [A.x A.y B.x.c0 B.x.c1 B.y.c0 B.y.c1 C.x C.y] x 32 bytes LE standard form. */
int zkref_groth16_prove(const zkref_circuit* C, const zkref_zkey* K, const uint8_t* w, const uint8_t* r32, const uint8_t* s32,
                        int threads, uint8_t* proof256);
/* Toy setup for SMALL circuits from explicit toxic waste (tau, alpha, beta, gamma, delta: 5 x 32 bytes). */
int zkref_groth16_setup(const zkref_circuit* C, unsigned log_n, uint32_t n_public, const uint8_t* toxic,
                        uint8_t* A, uint8_t* B1, uint8_t* B2, uint8_t* Cs, uint8_t* H, uint8_t* IC,
                        uint8_t* alpha1, uint8_t* beta1, uint8_t* delta1, uint8_t* beta2, uint8_t* gamma2, uint8_t* delta2);

#endif
