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

#endif
