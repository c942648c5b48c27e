/*
 * zkemail_b200.h - C ABI of the B200-native EmailVerifier witness + Groth16 proving engine.
 *
 * This is the drop-in boundary.  It replaces, for the EmailVerifier path, what the
 * reference reaches through
 *     snarkjs.groth16.fullProve(input, wasm, zkey)   /root/reference/packages/helpers/src/chunked-zkey.ts:80-84
 *     snarkjs.groth16.verify(vkey, publicSignals, proof)   /root/reference/packages/helpers/src/chunked-zkey.ts:101
 * and, at test time, the circom_tester verbs
 *     wasm_tester(circuit) / calculateWitness / checkConstraints / assertOut
 *                                                 /root/reference/packages/circuits/tests/email-verifier.test.ts:21-44,204
 *
 * Conventions
 *   - Field elements cross the ABI as 32-byte little-endian integers in standard (non-Montgomery) form,
 *     the same image as a `.wtns` entry (data formats).
 *   - All buffers are caller-owned host memory; contexts own device memory.  No torch / C++ types.
 *   - Return codes: 0 ok; > 0 per-item failure (e.g. a witness that violates a constraint: the message
 *     contains "Assert Failed", the string the reference's tests match, email-verifier.test.ts:78);
 *     < 0 fatal (bad arguments, CUDA failure, library built without a device).
 *   - `err`/`errcap`: optional message buffer, always NUL-terminated when errcap > 0.
 */
#ifndef ZKEMAIL_B200_H
#define ZKEMAIL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zke_circuit zke_circuit; /* R1CS + levelised witness program (the ".r1cs + .wasm" pair) */
typedef struct zke_zkey zke_zkey;       /* Groth16 proving + verification key (the ".zkey")           */
typedef struct zke_ctx zke_ctx;         /* circuit + zkey resident on one GPU, with work buffers       */

#define ZKE_FR_BYTES 32

/* ---------------------------------------------------------------------------------------------------
 * Circuits.  Replaces "circom file -> r1cs + witness calculator" (wasm_tester(...) in
 * /root/reference/packages/circuits/tests/email-verifier.test.ts:21-31).
 * `template_name` is a template of /root/reference/packages/circuits (e.g. "EmailVerifier", "Sha256Bytes",
 * "RSAVerifier65537", "FpMul", "Base64Lookup", ...); `params` are its circom template parameters in order.
 * For "EmailVerifier": {maxHeadersLength, maxBodyLength, n, k, ignoreBodyHashCheck, enableHeaderMasking,
 * enableBodyMasking, removeSoftLineBreaks, publicPubkey(0/1: `component main { public [pubkey] }`)}.
 * ------------------------------------------------------------------------------------------------- */
zke_circuit* zke_circuit_build(const char* template_name, const int64_t* params, size_t n_params,
                               char* err, size_t errcap);
void zke_circuit_free(zke_circuit* c);

typedef struct zke_circuit_info {
    uint32_t n_vars;        /* witness length m, w[0] = 1 */
    uint32_t n_temps;       /* scratch slots used by the witness program (not part of the witness) */
    uint32_t n_outputs, n_pub_inputs, n_prv_inputs;
    uint32_t n_public;      /* snarkjs nPublic = n_outputs + n_pub_inputs */
    uint32_t n_constraints;
    uint32_t n_levels;      /* dependency depth of the witness program */
    uint32_t n_ops;
    uint32_t n_coefs;
    uint32_t domain_log2;   /* Groth16 evaluation domain: 2^domain_log2 >= n_constraints + n_public + 1 */
    uint32_t n_groups;      /* named signal groups (outputs and inputs) */
    uint64_t nnz_a, nnz_b, nnz_c;
} zke_circuit_info;
int zke_circuit_get_info(const zke_circuit* c, zke_circuit_info* out);

/* Named signals (the .sym role for main's inputs/outputs; circom_tester assertOut / snarkjs input JSON keys).
 * kind: 0 output, 1 public input, 2 private input.  `first` is the witness index of element 0. */
int zke_circuit_group(const zke_circuit* c, uint32_t index, char* name, size_t namecap,
                      uint32_t* first, uint32_t* count, int* kind);
/* Index (into the packed input vector, i.e. witness index - 1 - n_outputs) of a named input; < 0 if absent. */
int64_t zke_circuit_input_offset(const zke_circuit* c, const char* name, uint32_t* count);

/* Raw read-only views of the flat circuit arrays (for exporters and for the test oracle).  `which`: */
enum {
    ZKE_ARR_COEFS = 0,      /* uint8[n_coefs][32]  interned coefficients, standard form, LE            */
    ZKE_ARR_A_PTR = 1, ZKE_ARR_A_VAR = 2, ZKE_ARR_A_COEF = 3,     /* uint32 CSR of matrix A            */
    ZKE_ARR_B_PTR = 4, ZKE_ARR_B_VAR = 5, ZKE_ARR_B_COEF = 6,
    ZKE_ARR_C_PTR = 7, ZKE_ARR_C_VAR = 8, ZKE_ARR_C_COEF = 9,
    ZKE_ARR_OPS = 10,       /* uint32[n_ops][5] = {code, dst, a, b, c}, sorted by level               */
    ZKE_ARR_LEVEL_PTR = 11, /* uint32[n_levels + 1]                                                   */
    ZKE_ARR_LC_PTR = 12, ZKE_ARR_LC_VAR = 13, ZKE_ARR_LC_COEF = 14, /* LC pool of the witness program */
    ZKE_ARR_AUX = 15,       /* uint32 operands of OP_FPMUL                                            */
    ZKE_ARR_SCOPE_OF_CONSTRAINT = 16 /* uint16[n_constraints]                                         */
};
const void* zke_circuit_array(const zke_circuit* c, int which, size_t* n_elems);
const char* zke_circuit_scope_name(const zke_circuit* c, uint32_t scope_index);

/* Library / device introspection.  zke_device_count() returns 0 when no CUDA device is usable;
 * every compute entry point then fails with a negative code (no CPU fallback exists). */
int zke_device_count(void);
const char* zke_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ZKEMAIL_B200_H */
