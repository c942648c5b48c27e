"""Mirror of the prove / verify shim of /root/reference/packages/helpers/src/chunked-zkey.ts:76-105.

The reference resolves `${circuitName}.wasm` / `${circuitName}.zkey` by URL and hands them to snarkjs; here a
circuit name resolves to a registered (Circuit, Zkey, Context) triple resident on a GPU."""
from __future__ import annotations
import json
from .engine import Context, Zkey, proof_to_json, verify
from .circuit import Circuit

_REGISTRY: dict[str, tuple] = {}


class InsecureKeyError(RuntimeError):
    pass


def register_zkey_files(circuit_name: str, circuit: Circuit, chunk_files, device: int = 0, max_batch: int = 1):
    """downloadProofFiles + uncompress (chunked-zkey.ts:35-37, 59-74) for local files: `chunk_files` are the paths of
    `${circuitName}.zkeyb` .. `.zkeyk` (or a single whole `.zkey`)."""
    blobs = [open(p, "rb").read() for p in chunk_files]
    zkey = Zkey.load(blobs[0], device=device, circuit=circuit) if len(blobs) == 1 else Zkey.load_chunks(blobs, device=device, circuit=circuit)
    return register_circuit(circuit_name, circuit, zkey, device=device, max_batch=max_batch)


def register_circuit(circuit_name: str, circuit: Circuit, zkey: Zkey, device: int = 0, max_batch: int = 1,
                     allow_toy_key: bool = False):
    """Plays the role of downloadProofFiles (chunked-zkey.ts:59-74): makes the proving artefacts of `circuitName`
    available to generateProof / verifyProof.  A key made by the seeded toy setup (Zkey(circuit, seed): the toxic waste
    is known, anyone can forge proofs that verify under it) is refused unless `allow_toy_key=True` - benchmarks and
    tests only; production keys come from a ceremony `.zkey` through Zkey.load / Zkey.load_chunks."""
    if zkey.is_toy and not allow_toy_key:
        raise InsecureKeyError("refusing a proving key made by the toy setup (known toxic waste); load a ceremony .zkey with "
                               "Zkey.load(...) or pass allow_toy_key=True for tests")
    ctx = Context(circuit, zkey, device=device, max_batch=max_batch)
    _REGISTRY[circuit_name] = (circuit, zkey, ctx, zkey.vkey())
    return ctx


def generate_proof(input: dict, base_url: str, circuit_name: str):
    """generateProof(input, baseUrl, circuitName) -> {proof, publicSignals}  (chunked-zkey.ts:76-91 ->
    snarkjs.groth16.fullProve)."""
    if circuit_name not in _REGISTRY:
        raise KeyError(f"Error downloading {base_url}{circuit_name}.zkey after 3 retries")   # chunked-zkey.ts:32
    circuit, _, ctx, _ = _REGISTRY[circuit_name]
    packed = circuit.pack_inputs(input)
    proofs, publics, _ = ctx.fullprove(packed, 1)
    proof, public_signals = proof_to_json(proofs[:256], publics, circuit.info.n_public)
    return {"proof": proof, "publicSignals": public_signals}


def verify_proof(proof: dict, public_signals, base_url: str, circuit_name: str) -> bool:
    """verifyProof(proof, publicSignals, baseUrl, circuitName) (chunked-zkey.ts:93-105 -> snarkjs.groth16.verify)."""
    if circuit_name not in _REGISTRY:
        raise KeyError(f"Error downloading {base_url}{circuit_name}.vkey.json after 3 retries")
    vkey = _REGISTRY[circuit_name][3]
    return verify(vkey, public_signals, proof)


# camelCase aliases matching the reference's exports
generateProof = generate_proof
verifyProof = verify_proof
