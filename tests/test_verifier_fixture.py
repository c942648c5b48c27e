"""The product's host verifier (zke_verify_json, pairing_host.cpp) on the reference's only Groth16 KAT:
/root/reference/packages/rust-verifier/tests/data/proof_of_twitter/{proof,public,vkey}.json must verify true
(/root/reference/packages/rust-verifier/tests/verifier_utils.rs:11-18); tampered variants must not.  Also pins
vk_alphabeta_12 (vkey.json:43) = the product's e(alpha_1, beta_2) export, and the G2 subgroup check."""
import ctypes
import json
import os

import zkemail_b200 as z
from zkemail_b200 import _lib as L

GOLD = os.path.join(os.path.dirname(__file__), "golden", "proof_of_twitter")
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583


def _load():
    return tuple(json.load(open(os.path.join(GOLD, n))) for n in ("vkey.json", "public.json", "proof.json"))


def test_product_verifier_accepts_reference_fixture():
    vkey, public, proof = _load()
    assert z.verify(vkey, public, proof) is True


def test_product_verifier_rejects_tampered_fixture():
    vkey, public, proof = _load()
    bad_pub = list(public)
    bad_pub[1] = str(int(bad_pub[1]) + 1)
    assert z.verify(vkey, bad_pub, proof) is False
    bad_proof = json.loads(json.dumps(proof))
    # another point of the curve (the negation of pi_c): well-formed, wrong
    bad_proof["pi_c"][1] = str(Q - int(bad_proof["pi_c"][1]))
    assert z.verify(vkey, public, bad_proof) is False
    swapped = json.loads(json.dumps(proof))
    swapped["pi_a"], swapped["pi_c"] = swapped["pi_c"], swapped["pi_a"]
    assert z.verify(vkey, public, swapped) is False
    assert z.verify(vkey, public[:-1] + [str(int(public[-1]) ^ 1)], proof) is False


def test_product_verifier_rejects_point_off_curve_and_out_of_subgroup():
    vkey, public, proof = _load()
    off = json.loads(json.dumps(proof))
    off["pi_a"][0] = str(int(off["pi_a"][0]) + 1)
    assert z.verify(vkey, public, off) is False
    # a point of the twist curve outside the order-r subgroup: x = 1 + 0u gives one for this curve only if
    # x^3 + b is a square in Fq2 - search a few small x and take the first that lands on the curve
    from oracle import bn254
    found = None
    for x0 in range(1, 60):
        x = (x0, 0)
        rhs = bn254.f2_add(bn254.f2_mul(bn254.f2_sqr(x), x), bn254.B2)
        # square root in Fq2 via the norm trick (q = 3 mod 4)
        a, b = rhs
        n = (a * a + b * b) % Q
        s = pow(n, (Q + 1) // 4, Q)
        if s * s % Q != n:
            continue
        for sgn in (s, Q - s):
            t = (a + sgn) * pow(2, -1, Q) % Q
            y0 = pow(t, (Q + 1) // 4, Q)
            if y0 * y0 % Q != t or y0 == 0:
                continue
            y1 = b * pow(2 * y0, -1, Q) % Q
            y = (y0, y1)
            if bn254.f2_sqr(y) == rhs:
                found = (x, y)
                break
        if found:
            break
    assert found is not None and bn254.g2_is_on_curve(found)
    # not in the order-r subgroup (cofactor > 1): [r - 1] P != -P   (g2_mul reduces its scalar mod r)
    assert bn254.g2_mul(found, bn254.R - 1) != bn254.g2_neg(found)
    rogue = json.loads(json.dumps(proof))
    rogue["pi_b"] = [[str(found[0][0]), str(found[0][1])], [str(found[1][0]), str(found[1][1])], ["1", "0"]]
    assert z.verify(vkey, public, rogue) is False


def test_alphabeta_matches_reference_fixture():
    vkey, _, _ = _load()
    le = lambda v: int(v).to_bytes(32, "little")
    alpha = le(vkey["vk_alpha_1"][0]) + le(vkey["vk_alpha_1"][1])
    beta = le(vkey["vk_beta_2"][0][0]) + le(vkey["vk_beta_2"][0][1]) + le(vkey["vk_beta_2"][1][0]) + le(vkey["vk_beta_2"][1][1])
    out = ctypes.create_string_buffer(384)
    assert L.zke_pairing_alphabeta(alpha, beta, out) == 0
    got = [[[str(int.from_bytes(out.raw[32 * ((i * 3 + j) * 2 + k):32 * ((i * 3 + j) * 2 + k) + 32], "little")) for k in range(2)]
            for j in range(3)] for i in range(2)]
    assert got == vkey["vk_alphabeta_12"]
    # malformed input: alpha off the curve
    assert L.zke_pairing_alphabeta(le(1) + le(1), beta, out) != 0


# ---- batch verification (zke_verify_batch_json) -------------------------------------------------------
R_ORDER = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _rerandomised(proof, k):
    """Another valid proof of the same statement: (k A, k^-1 B, C) - e(kA, B / k) = e(A, B)."""
    from oracle import bn254
    a = bn254.g1_mul(bn254.g1_from_json(proof["pi_a"]), k)
    b = bn254.g2_mul(bn254.g2_from_json(proof["pi_b"]), pow(k, -1, R_ORDER))
    out = json.loads(json.dumps(proof))
    out["pi_a"], out["pi_b"] = bn254.g1_to_json(a), bn254.g2_to_json(b)
    return out


def test_batch_verifier_accepts_distinct_valid_proofs_and_names_offenders():
    vkey, public, proof = _load()
    proofs = [proof, _rerandomised(proof, 7), _rerandomised(proof, 0x1234567890abcdef)]
    assert proofs[1] != proof and z.verify(vkey, public, proofs[1]) is True
    assert z.verify_batch(vkey, [public] * 3, proofs) == [True, True, True]
    assert z.verify_batch(vkey, [public] * 3, proofs, rand=bytes(range(1, 49))) == [True, True, True]   # fixed randomness
    assert z.verify_batch(vkey, [], []) == []
    # one bad member: the combined check fails and the per-proof pass names it
    bad = json.loads(json.dumps(proof))
    bad["pi_c"][1] = str(Q - int(bad["pi_c"][1]))
    assert z.verify_batch(vkey, [public] * 3, [proofs[0], bad, proofs[2]]) == [True, False, True]
    bad_pub = list(public)
    bad_pub[1] = str(int(bad_pub[1]) + 1)
    assert z.verify_batch(vkey, [public, public, bad_pub], proofs) == [True, True, False]
    # errors that cancel under equal weights are caught by the random ones: C_0 + D, C_1 - D
    from oracle import bn254
    d = bn254.g1_mul((1, 2), 5)
    p0, p1 = json.loads(json.dumps(proof)), json.loads(json.dumps(proofs[1]))
    p0["pi_c"] = bn254.g1_to_json(bn254.g1_add(bn254.g1_from_json(proof["pi_c"]), d))
    p1["pi_c"] = bn254.g1_to_json(bn254.g1_add(bn254.g1_from_json(proofs[1]["pi_c"]), bn254.g1_neg(d)))
    assert z.verify_batch(vkey, [public] * 2, [p0, p1]) == [False, False]


def test_batch_verifier_malformed_input():
    import pytest
    vkey, public, proof = _load()
    with pytest.raises(ValueError):
        z.verify_batch(vkey, [public], [proof, proof])
    with pytest.raises(ValueError):
        z.verify_batch(vkey, [public], [proof], rand=b"short")
    off = json.loads(json.dumps(proof))
    off["pi_a"][0] = str(int(off["pi_a"][0]) + 1)          # off the curve: invalid, not an error
    assert z.verify_batch(vkey, [public, public], [proof, off]) == [True, False]
    assert z.verify_batch(vkey, [public[:-1]], [proof]) == [False]    # wrong number of public signals
