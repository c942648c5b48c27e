"""Pins the oracle's Groth16 verifier on the reference's only Groth16 known-answer fixture
(/root/reference/packages/rust-verifier/tests/data/proof_of_twitter/*, asserted true at
/root/reference/packages/rust-verifier/tests/verifier_utils.rs:11-18).  Copies live in tests/golden/proof_of_twitter/."""
import copy
import json
import os
import pytest
from oracle import bn254

D = os.path.join(os.path.dirname(__file__), "golden", "proof_of_twitter")


@pytest.fixture(scope="module")
def fixture():
    return tuple(json.load(open(os.path.join(D, f))) for f in ("vkey.json", "public.json", "proof.json"))


def test_fixture_verifies(fixture):
    vk, pub, proof = fixture
    assert bn254.groth16_verify(vk, pub, proof) is True


def test_tampered_public_rejected(fixture):
    vk, pub, proof = fixture
    for i in range(len(pub)):
        bad = list(pub)
        bad[i] = str(int(bad[i]) ^ 1)
        assert bn254.groth16_verify(vk, bad, proof) is False


def test_tampered_proof_rejected(fixture):
    vk, pub, proof = fixture
    bad = copy.deepcopy(proof)
    bad["pi_c"][0], bad["pi_c"][1] = bad["pi_a"][0], bad["pi_a"][1]   # still on the curve, wrong point
    assert bn254.groth16_verify(vk, pub, bad) is False
    bad = copy.deepcopy(proof)
    bad["pi_a"][0] = str(int(bad["pi_a"][0]) + 1)                    # off the curve
    assert bn254.groth16_verify(vk, pub, bad) is False


def test_vk_alphabeta_12_matches(fixture):
    """snarkjs vkeys carry e(alpha, beta) as vk_alphabeta_12 (2 x 3 x 2 tower coefficients).  Check the oracle's
    pairing against it through the verification equation only (tower layout differs from the flat Fq12 used here)."""
    vk, _, _ = fixture
    a = bn254.g1_from_json(vk["vk_alpha_1"])
    b = bn254.g2_from_json(vk["vk_beta_2"])
    e1 = bn254.pairing(b, a)
    e2 = bn254.pairing(bn254.g2_mul(b, 5), a)
    assert e1 ** 5 == e2      # bilinearity
