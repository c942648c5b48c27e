// TypeScript host shim (UNRUN: node is not present in the build image).  Drop-in for the two snarkjs verbs the
// reference uses in packages/helpers/src/chunked-zkey.ts:80-84,101, backed by the C ABI of include/zkemail_b200.h.
// See INTEGRATION.md.
import koffi from 'koffi';

const lib = koffi.load(process.env.ZKEMAIL_B200_LIB ?? 'libzkemail_b200.so');
const zke_circuit_build = lib.func('void* zke_circuit_build(const char*, const int64_t*, size_t, char*, size_t)');
const zke_setup = lib.func('void* zke_setup(void*, uint64_t, int, char*, size_t)');
const zke_zkey_load = lib.func('void* zke_zkey_load(const void*, size_t, int, char*, size_t)');
const zke_zkey_load_chunks = lib.func('void* zke_zkey_load_chunks(const void**, const size_t*, size_t, int, char*, size_t)');
const zke_zkey_is_toy = lib.func('int zke_zkey_is_toy(void*)');
const zke_ctx_open = lib.func('void* zke_ctx_open(void*, void*, int, uint32_t, char*, size_t)');
const zke_zkey_vkey_json = lib.func('int zke_zkey_vkey_json(void*, char*, size_t*)');
const zke_fullprove_json = lib.func('int zke_fullprove_json(void*, void*, const char*, char*, size_t*, char*, size_t*, char*, size_t)');
const zke_verify_json = lib.func('int zke_verify_json(const char*, const char*, const char*, char*, size_t)');
const zke_verify_batch_json = lib.func('int zke_verify_batch_json(const char*, const char*, const char*, const uint8_t*, uint8_t*, char*, size_t)');

const cstr = (b: Buffer) => b.toString('utf8', 0, b.indexOf(0));
type Entry = { circuit: unknown; zkey: unknown; ctx: unknown };
const registry = new Map<string, Entry>();

/**
 * Plays the role of downloadProofFiles + uncompress (chunked-zkey.ts:35-37, 59-74): makes `${circuitName}.zkey`
 * resident on a GPU.  `zkeyChunks` are the contents of `${circuitName}.zkeyb` .. `.zkeyk` (the fork's per-section
 * files, chunked-zkey.ts:9) or a single whole `.zkey`; they are parsed and validated by zke_zkey_load[_chunks].
 * The circuit's witness program replaces `${circuitName}.wasm`.
 */
export function registerEmailVerifier(circuitName: string, params: number[], zkeyChunks: Buffer[], device = 0): void {
  const err = Buffer.alloc(4096);
  const circuit = zke_circuit_build('EmailVerifier', BigInt64Array.from(params.map(BigInt)), params.length, err, err.length);
  if (!circuit) throw new Error(cstr(err));
  const zkey = zkeyChunks.length === 1
    ? zke_zkey_load(zkeyChunks[0], zkeyChunks[0].length, device, err, err.length)
    : zke_zkey_load_chunks(zkeyChunks, BigUint64Array.from(zkeyChunks.map((c) => BigInt(c.length))), zkeyChunks.length, device, err, err.length);
  if (!zkey) throw new Error(cstr(err));
  const ctx = zke_ctx_open(circuit, zkey, device, 1, err, err.length);
  if (!ctx) throw new Error(cstr(err));
  registry.set(circuitName, { circuit, zkey, ctx });
}

/** Tests and benchmarks only: a key from the seeded TOY setup (known toxic waste - proofs under it are forgeable). */
export function registerEmailVerifierWithToyKey(circuitName: string, params: number[], seed: number, device = 0): void {
  const err = Buffer.alloc(4096);
  const circuit = zke_circuit_build('EmailVerifier', BigInt64Array.from(params.map(BigInt)), params.length, err, err.length);
  if (!circuit) throw new Error(cstr(err));
  const zkey = zke_setup(circuit, BigInt(seed), device, err, err.length);
  if (!zkey || zke_zkey_is_toy(zkey) !== 1) throw new Error(cstr(err));
  const ctx = zke_ctx_open(circuit, zkey, device, 1, err, err.length);
  if (!ctx) throw new Error(cstr(err));
  registry.set(circuitName, { circuit, zkey, ctx });
}

export function exportVerificationKey(circuitName: string): object {
  const e = registry.get(circuitName);
  if (!e) throw new Error(`unknown circuit ${circuitName}`);
  const len = [1 << 20];
  const buf = Buffer.alloc(len[0]);
  if (zke_zkey_vkey_json(e.zkey, buf, len) !== 0) throw new Error('vkey export failed');
  return JSON.parse(cstr(buf));
}

export const groth16 = {
  /** snarkjs.groth16.fullProve(input, wasmFile, zkeyFileName) */
  async fullProve(input: object, _wasmFile: string, zkeyFileName: string) {
    const e = registry.get(zkeyFileName.replace(/\.zkey$/, ''));
    if (!e) throw new Error(`Error downloading ${zkeyFileName} after 3 retries`);
    const proof = Buffer.alloc(4096), pub = Buffer.alloc(1 << 16), err = Buffer.alloc(4096);
    const pl = [proof.length], sl = [pub.length];
    const rc = zke_fullprove_json(e.ctx, e.circuit, JSON.stringify(input), proof, pl, pub, sl, err, err.length);
    if (rc !== 0) throw new Error(cstr(err));
    return { proof: JSON.parse(cstr(proof)), publicSignals: JSON.parse(cstr(pub)) };
  },
  /** snarkjs.groth16.verify(vkey, publicSignals, proof) */
  async verify(vkey: object, publicSignals: string[], proof: object): Promise<boolean> {
    const err = Buffer.alloc(4096);
    const rc = zke_verify_json(JSON.stringify(vkey), JSON.stringify(publicSignals), JSON.stringify(proof), err, err.length);
    if (rc < 0) throw new Error(cstr(err));
    return rc === 1;
  },
  /** n proofs under one key with one randomised product of pairings; per-proof verdicts (not in snarkjs) */
  async verifyBatch(vkey: object, publicSignals: string[][], proofs: object[]): Promise<boolean[]> {
    const err = Buffer.alloc(4096), ok = Buffer.alloc(proofs.length);
    const rand = require('crypto').randomBytes(16 * proofs.length);
    const rc = zke_verify_batch_json(JSON.stringify(vkey), JSON.stringify(publicSignals), JSON.stringify(proofs), rand, ok, err, err.length);
    if (rc < 0) throw new Error(cstr(err));
    return Array.from(ok).map((b) => b === 1);
  },
};
