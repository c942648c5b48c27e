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
/* zk-regex circuit of an arbitrary decomposed regex (the generator behind BodyHashRegex; un-vendored
 * @zk-email/zk-regex-circom, call site /root/reference/packages/circuits/email-verifier.circom:5,126):
 * `parts[i]` is a regex fragment, `is_public[i]` != 0 marks the fragment whose matched bytes are revealed.
 * Signals: input msg[msg_len], outputs out (match flag) and reveal0[msg_len]. */
zke_circuit* zke_circuit_build_regex(const char* const* parts, const uint8_t* is_public, size_t n_parts, uint32_t msg_len,
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
    ZKE_ARR_SCOPE_OF_CONSTRAINT = 16, /* uint16[n_constraints]                                        */
    ZKE_ARR_SHA_BLOCKS = 17, /* uint32: {n_blocks, per block: var_begin, var_end, temp_begin, temp_end, n_desc,
                               inputs[768], desc[n_desc][2] = {signal, quantity << 8 | bit}} - the Sha256compression
                               instances the engine evaluates natively (one compression instead of ~320 levels)  */
    ZKE_ARR_REGEX_SEEDS = 18 /* uint32: {n_seeds, per seed: n_desc, n_bytes, n_states | mode << 31, first_mask lo, hi,
                               bytes[n_bytes] (signal of message byte j), table[n_states * 64] (destination state of
                               (source, byte), 0xff = none, 4 per word), mode 1 (compact shape): group[n_states * 64] (the
                               product that fires), desc[n_desc][2] = {signal, position << 8 | state or product}} - the
                               regex instances whose chained signals the engine seeds with one automaton run      */
};
const void* zke_circuit_array(const zke_circuit* c, int which, size_t* n_elems);
const char* zke_circuit_scope_name(const zke_circuit* c, uint32_t scope_index);


/* ---------------------------------------------------------------------------------------------------
 * Proving key.  zke_setup() is a TOY trusted setup (`snarkjs groth16 setup` role,
 * /root/reference/docs/zk-email-docs/UsageGuide/README.md:139-153): tau, alpha, beta, gamma, delta are derived
 * from `seed`, i.e. the toxic waste is known - benchmark / test keys only.  The field side runs on the host, the
 * ~5m + N fixed-base scalar multiplications on GPU `device`; the key stays resident on that GPU.
 * ------------------------------------------------------------------------------------------------- */
zke_zkey* zke_setup(const zke_circuit* c, uint64_t seed, int device, char* err, size_t errcap);
/* Real proving keys: an iden3 `.zkey` (Groth16, BN254) as produced by `snarkjs groth16 setup` / `zkey contribute` -
 * the third argument of snarkjs.groth16.fullProve(input, wasm, zkey)
 * (/root/reference/packages/helpers/src/chunked-zkey.ts:80-84) and the first of `snarkjs groth16 prove zkey wtns`
 * (/root/reference/docs/zk-email-docs/UsageGuide/README.md:139-195).  Sections 2-9 are validated (field moduli, every
 * point on its curve) and made resident on GPU `device`: points as they are stored (affine, Montgomery), the A / B
 * coefficient matrices of section 4 as CSR, the H points with their fixed-base table.
 * zke_zkey_load_chunks takes the fork's chunked form: chunk i holds section i + 1, i.e. the files `${name}.zkeyb` ..
 * `${name}.zkeyk` (chunked-zkey.ts:9,35-37); n_chunks >= 9 (section 10, the contribution log, is not needed). */
zke_zkey* zke_zkey_load(const void* zkey_bytes, size_t len, int device, char* err, size_t errcap);
zke_zkey* zke_zkey_load_chunks(const void* const* chunks, const size_t* lens, size_t n_chunks, int device, char* err, size_t errcap);
/* Writes the key as a `.zkey` file image.  `c` supplies the coefficient section for keys made by zke_setup (pass NULL
 * for loaded keys, which carry their own).  out == NULL: returns the size needed; < 0 on error. */
int64_t zke_zkey_write(const zke_zkey* z, const zke_circuit* c, uint8_t* out, size_t cap);
/* 1 if the key came from zke_setup (toxic waste known - anyone can forge proofs for it), 0 for a loaded key. */
int zke_zkey_is_toy(const zke_zkey* z);
void zke_zkey_free(zke_zkey* z);
int zke_zkey_info(const zke_zkey* z, uint32_t* n_vars, uint32_t* n_public, uint32_t* domain_log2);
/* zkey sections (iden3 .zkey numbering where one exists: 3 IC, 5 A, 6 B1, 7 B2, 8 C/"L", 9 H; header points apart). */
enum {
    ZKE_SEC_ALPHA1 = 101, ZKE_SEC_BETA1 = 102, ZKE_SEC_DELTA1 = 103,
    ZKE_SEC_BETA2 = 104, ZKE_SEC_GAMMA2 = 105, ZKE_SEC_DELTA2 = 106,
    ZKE_SEC_IC = 3, ZKE_SEC_A = 5, ZKE_SEC_B1 = 6, ZKE_SEC_B2 = 7, ZKE_SEC_C = 8, ZKE_SEC_H = 9
};
/* Copies a section to host memory as affine points, standard-form LE coordinates (G1 64 bytes: x,y;
 * G2 128 bytes: x.c0,x.c1,y.c0,y.c1; infinity = zeros).  Returns the number of points (with out == NULL: just the
 * count), < 0 on error.  Sections A/B1/B2/C have n_vars entries (C is infinity for the public signals). */
int64_t zke_zkey_section(const zke_zkey* z, int section, uint8_t* out, size_t cap);

/* ---------------------------------------------------------------------------------------------------
 * Contexts: circuit (+ optional proving key) resident on one GPU with work buffers for `max_batch` emails.
 * One host thread per context (or external locking).  All calls are synchronous at the ABI.
 * ------------------------------------------------------------------------------------------------- */
/* `c` may be NULL when the key was loaded from a `.zkey`: such a context proves externally computed witnesses
 * (zke_load_witness / zke_wtns_prove + zke_prove) from the key's own coefficient matrices. */
zke_ctx* zke_ctx_open(const zke_circuit* c, const zke_zkey* zkey_or_null, int device, uint32_t max_batch,
                      char* err, size_t errcap);
void zke_ctx_close(zke_ctx* x);
void* zke_ctx_stream(const zke_ctx* x);      /* the cudaStream_t all of the context's work is enqueued on */
uint64_t zke_kernel_launches(void);          /* kernels launched by this library since it was loaded */

/* Copies a batch of packed inputs into the context's device buffer; a later zke_witness / zke_fullprove call with
 * inputs == NULL uses them (lets a benchmark start its timed region with the inputs already resident in HBM). */
int zke_upload_inputs(zke_ctx* x, const uint8_t* inputs, size_t batch, char* err, size_t errcap);

/* Optional per-stage device timing (CUDA events on the context's stream, accumulated over calls). */
enum {
    ZKE_STAGE_WITNESS = 0,        /* witness kernel, whole batch                                   */
    ZKE_STAGE_MATVEC = 1,         /* <A,w>, <B,w>, constraint check, per email                     */
    ZKE_STAGE_NTT = 2,            /* hadamard + 3 inverse + 3 forward NTTs + quotient, per email   */
    ZKE_STAGE_MSM_A = 3, ZKE_STAGE_MSM_B1 = 4, ZKE_STAGE_MSM_C = 5,
    ZKE_STAGE_MSM_H = 6,          /* whole H multi-exponentiation (N full-width scalars)           */
    ZKE_STAGE_MSM_H_BUCKETS = 7,  /* its bucket-accumulation kernel alone (the dominant kernel)    */
    ZKE_STAGE_MSM_B2 = 8,
    ZKE_N_STAGES = 9
};
/* Number of concurrent proving lanes (streams) used by zke_prove, 1..allocated (default min(8, max_batch), or the
 * ZKE_LANES environment variable at zke_ctx_open).  Returns the value in effect.  Profiling forces one lane. */
int zke_ctx_set_lanes(zke_ctx* x, int n);
int zke_ctx_profile(zke_ctx* x, int enable);                       /* enabling also clears the accumulators */
int zke_ctx_profile_get(const zke_ctx* x, double* ms_out, uint64_t* count_out);  /* arrays of ZKE_N_STAGES */

/* calculateWitness + checkConstraints for a batch (circom_tester verbs; witness step of fullProve).
 * inputs: [batch][n_inputs][32] in witness order (see zke_circuit_input_offset), or NULL to use the inputs made
 * resident by zke_upload_inputs.  wtns_out (optional):
 * [batch][n_vars][32], the `.wtns` payload.  status (optional): per email, -1 = satisfied, else the index of
 * the first violated constraint.  Returns the number of failing emails (message: "Assert Failed: ..."). */
int zke_witness(zke_ctx* x, const uint8_t* inputs, size_t batch, uint8_t* wtns_out, int32_t* status,
                char* err, size_t errcap);
/* Makes host witnesses resident (snarkjs `groth16 prove zkey wtns` entry). */
int zke_load_witness(zke_ctx* x, const uint8_t* wtns, size_t batch, char* err, size_t errcap);
/* Groth16 prove for the resident witnesses (snarkjs.groth16.prove).  rs (optional): [batch][2][32] fixed blinding
 * scalars r, s (parity tests); NULL draws them from /dev/urandom.  proofs_out: [batch][8][32] =
 * A.x, A.y, B.x.c0, B.x.c1, B.y.c0, B.y.c1, C.x, C.y (standard form, LE).  publics_out: [batch][n_public][32]. */
int zke_prove(zke_ctx* x, size_t batch, const uint8_t* rs, uint8_t* proofs_out, uint8_t* publics_out,
              int32_t* status, char* err, size_t errcap);
/* witness + prove in one call with host buffers (snarkjs.groth16.fullProve,
 * /root/reference/packages/helpers/src/chunked-zkey.ts:80-84). */
int zke_fullprove(zke_ctx* x, const uint8_t* inputs, size_t batch, const uint8_t* rs, uint8_t* proofs_out,
                  uint8_t* publics_out, int32_t* status, char* err, size_t errcap);

/* Pipelined form of zke_fullprove: _submit enqueues the H2D copy, the witness kernel and all proving kernels of one
 * batch and returns; _collect waits for the oldest submitted batch, finishes its proofs on the host and returns them.
 * Up to two batches may be in flight, so the (latency-bound) witness kernel of batch k + 1 runs under the proving
 * kernels of batch k.  zke_fullprove == submit + collect.  Returns as zke_fullprove. */
int zke_fullprove_submit(zke_ctx* x, const uint8_t* inputs, size_t batch, const uint8_t* rs, char* err, size_t errcap);
int zke_fullprove_collect(zke_ctx* x, uint8_t* proofs_out, uint8_t* publics_out, int32_t* status, char* err, size_t errcap);
/* `snarkjs groth16 prove <zkey> <wtns>` (/root/reference/docs/zk-email-docs/UsageGuide/README.md:139-195): one iden3
 * `.wtns` file image in, proof + public signals out. */
int zke_wtns_prove(zke_ctx* x, const void* wtns_bytes, size_t len, const uint8_t* rs, uint8_t* proof_out, uint8_t* publics_out,
                   char* err, size_t errcap);

/* ---------------------------------------------------------------------------------------------------
 * One proof across 2, 4 or 8 GPUs (BASELINE configs[3] "sharded MSM", configs[4] "NCCL-sharded MSM+NTT").
 * Every GPU (one process each) opens a context with the key and loads the same witness (zke_witness / zke_load_witness,
 * batch 1); then, in lock step:
 *     zke_shard_begin   mat-vec for the rows of this GPU's column range + the cross-block inverse NTT stages
 *     [exchange 1]      all-to-all "columns -> rows" on the three vectors of zke_shard_vector (caller: NCCL / torch.distributed)
 *     zke_shard_mid     block-local inverse stages, coset shift, block-local forward stages on this GPU's row block
 *     [exchange 2]      all-to-all "rows -> columns"
 *     zke_shard_end     cross-block forward stages, a o b - c on the column range (this GPU's H scalars), the five
 *                       multi-exponentiations over this GPU's share of the points -> ZKE_SHARD_PARTIAL_BYTES
 *     [all-gather]      of the partial blocks
 *     zke_shard_combine (host only) adds the partial points and assembles the proof: bit-identical to the 1-GPU proof.
 * Layout of the vectors: element i of the N = 2^domain_log2 evaluation vector at byte offset 32 i; block g = elements
 * [g M, (g + 1) M) with M = N / world; GPU r owns row block r and the columns [r M / world, (r + 1) M / world) of
 * every block.  Exchange 1 sends (block g, columns of r) to GPU g; exchange 2 is its inverse.
 * ------------------------------------------------------------------------------------------------- */
#define ZKE_SHARD_PARTIAL_BYTES 388   /* A, B1, C, H (G1: x, y) + B2 (G2: x.c0, x.c1, y.c0, y.c1), standard form LE, + u32 first bad row */
int zke_shard_begin(zke_ctx* x, int rank, int world, char* err, size_t errcap);
void* zke_shard_vector(zke_ctx* x, int which /* 0 a, 1 b, 2 c */, size_t* n_elems);   /* device pointer, 32 bytes per element */
int zke_shard_mid(zke_ctx* x, char* err, size_t errcap);
int zke_shard_end(zke_ctx* x, uint8_t* partial_out, uint8_t* publics_out, char* err, size_t errcap);
int zke_shard_combine(const zke_zkey* z, const uint8_t* partials, int world, const uint8_t* rs, uint8_t* proof_out, int32_t* status,
                      char* err, size_t errcap);
/* the same with the five key points it needs given directly: alpha_1, beta_1, delta_1 (64 bytes each), beta_2, delta_2
 * (128 bytes each), standard form LE - no GPU-resident key required (a coordinator process can combine) */
int zke_shard_combine_raw(const uint8_t* key_points, const uint8_t* partials, int world, const uint8_t* rs, uint8_t* proof_out,
                          int32_t* status, char* err, size_t errcap);

/* ---------------------------------------------------------------------------------------------------
 * JSON faces of the boundary (snarkjs file formats; shapes as in
 * /root/reference/packages/rust-verifier/tests/data/proof_of_twitter/{vkey,proof,public}.json).
 * String outputs: pass the buffer capacity in *len; on return *len = bytes needed incl. NUL (rc -2 if too small).
 * ------------------------------------------------------------------------------------------------- */
/* snarkjs.groth16.verify(vkey, publicSignals, proof): 1 valid, 0 invalid, < 0 malformed input.  Host only. */
int zke_verify_json(const char* vkey_json, const char* public_json, const char* proof_json, char* err, size_t errcap);
/* The same check for n proofs under ONE verification key ("batch Groth16 verification"; the reference
 * verifies one proof per call: chunked-zkey.ts:93-105, rust-verifier/src/verifier_utils.rs:20): a random linear combination
 * turns the 4n pairings into n + 3 Miller loops and one final exponentiation.  publics_json: array of n public-signal
 * arrays, proofs_json: array of n proof objects; rand16: n x 16 bytes of caller randomness the provers cannot predict
 * (NULL: std::random_device); ok (may be NULL): ok[i] = 1 / 0 - when the combined check fails the proofs are verified one
 * by one to name the offenders.  Returns the number of valid proofs, < 0 on malformed input.  Host only. */
int zke_verify_batch_json(const char* vkey_json, const char* publics_json, const char* proofs_json, const uint8_t* rand16,
                          uint8_t* ok, char* err, size_t errcap);
/* `snarkjs zkey export verificationkey`, incl. vk_alphabeta_12 = e(alpha_1, beta_2) in snarkjs' Fq12 tower layout
 * (/root/reference/packages/rust-verifier/tests/data/proof_of_twitter/vkey.json:43). */
int zke_zkey_vkey_json(const zke_zkey* z, char* out, size_t* len);
/* e(alpha_1, beta_2) as snarkjs exports it: 12 x 32 bytes = vk_alphabeta_12[i][j][k] flattened, standard form LE.
 * alpha: x, y; beta: x.c0, x.c1, y.c0, y.c1 (standard form LE).  Host only. */
int zke_pairing_alphabeta(const uint8_t* alpha64, const uint8_t* beta128, uint8_t* out384);
/* zke_prove output -> proof.json / public.json */
int zke_proof_to_json(const uint8_t* proof256, const uint8_t* publics, uint32_t n_public, char* proof_json, size_t* proof_len,
                      char* public_json, size_t* public_len);
/* input.json ({signal: decimal string | number | nested arrays}) -> packed inputs in witness order */
int zke_pack_inputs_json(const zke_circuit* c, const char* input_json, uint8_t* out, size_t cap, char* err, size_t errcap);
/* snarkjs.groth16.fullProve(input, wasm, zkey) for one email: JSON in, proof.json + public.json out */
int zke_fullprove_json(zke_ctx* x, const zke_circuit* c, const char* input_json, char* proof_json, size_t* proof_len,
                       char* public_json, size_t* public_len, char* err, size_t errcap);
/* Diagnostic: the FpMul big-integer hint evaluated on the host with the same code the witness kernel runs
 * (a, b, p: k limbs of 32 bytes LE; q, r out likewise). */
/* Diagnostic: the toxic waste (tau, alpha, beta, gamma, delta; 5 x 32 bytes LE) zke_setup derives from `seed` -
 * lets a test rebuild the same key independently.  It exists precisely because the setup is a toy. */
int zke_setup_toxic(uint64_t seed, uint8_t* out160);
int zke_selftest_fpmul_hint(uint32_t n, uint32_t k, const uint8_t* a, const uint8_t* b, const uint8_t* p, uint8_t* q, uint8_t* r);

/* Library / device introspection.  zke_device_count() returns 0 when no CUDA device is usable;
 * every compute entry point then fails with a negative code (no CPU fallback exists). */
int zke_device_count(void);
const char* zke_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ZKEMAIL_B200_H */
