"""The CPU oracle's Groth16 setup + prover, checked against the fixture-pinned verifier (oracle/bn254.py) on small
circuits.  (Prover parity is unpinned in the reference - no proof-producing test exists - so the anchor is: proofs
must satisfy the textbook verification equation under the verifier that accepts the reference's own fixture.)"""
import random
import pytest
from zkemail_b200 import Circuit, FR_MODULUS
from zkutil import (oracle_witness, oracle_setup, oracle_prove, vkey_from_sections, proof_json_from_bytes)
from oracle import bn254

TOXIC = (0x1234567890ABCDEF1234567, 0x2222222222222222222333, 0x3333333333444, 0x44444444444445555, 0x5555555566666)


@pytest.mark.parametrize("name,params,inputs", [
    ("Multiplier", [], {"a": 3, "b": 5}),
    ("FpMul", [2, 4], {"a": [1, 0, 1, 0], "b": [0, 1, 1, 0], "p": [1, 1, 1, 1]}),
    ("Poseidon", [2], {"inputs": [1, 2]}),
])
def test_oracle_prove_verifies(name, params, inputs):
    c = Circuit(name, params)
    sec = oracle_setup(c, TOXIC)
    w = oracle_witness(c, inputs)
    rnd = random.Random(name)
    r, s = rnd.randrange(FR_MODULUS), rnd.randrange(FR_MODULUS)
    proof = oracle_prove(c, sec, w.raw(), r, s, threads=4)
    npub = c.info.n_public
    vkey = vkey_from_sections(sec, npub)
    pubs = [str(w[1 + i]) for i in range(npub)]
    assert bn254.groth16_verify(vkey, pubs, proof_json_from_bytes(proof))
    if pubs:
        bad = list(pubs)
        bad[0] = str((int(bad[0]) + 1) % FR_MODULUS)
        assert not bn254.groth16_verify(vkey, bad, proof_json_from_bytes(proof))
    # a different (r, s) gives a different but equally valid proof (zero-knowledge blinding)
    proof2 = oracle_prove(c, sec, w.raw(), r + 1, s, threads=1)
    assert proof2 != proof and bn254.groth16_verify(vkey, pubs, proof_json_from_bytes(proof2))
