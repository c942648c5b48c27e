"""Verifier-argument encoders (zkemail_b200.verifier_args) on the reference's proof fixture
(/root/reference/packages/rust-verifier/tests/data/proof_of_twitter): ark-serialize compressed images
(main.rs:81-104, verifier_template.rs:17-31) round-trip to the JSON points and keep the proof verifying; Solidity
calldata and the CircomUtils packing helpers (contracts/utils/CircomUtils.sol:41-129)."""
import json
import os
import random

import zkemail_b200 as z
from zkemail_b200 import verifier_args as V
from oracle import bn254

GOLD = os.path.join(os.path.dirname(__file__), "golden", "proof_of_twitter")
vkey, public, proof = (json.load(open(os.path.join(GOLD, n))) for n in ("vkey.json", "public.json", "proof.json"))


def test_ark_compressed_proof_round_trip_and_layout():
    blob = V.ark_proof_compressed(proof)
    assert len(blob) == 128
    a, b, c = V.ark_g1_decompress(blob[:32]), V.ark_g2_decompress(blob[32:96]), V.ark_g1_decompress(blob[96:])
    assert a == (int(proof["pi_a"][0]), int(proof["pi_a"][1]))
    assert b == ((int(proof["pi_b"][0][0]), int(proof["pi_b"][0][1])), (int(proof["pi_b"][1][0]), int(proof["pi_b"][1][1])))
    assert c == (int(proof["pi_c"][0]), int(proof["pi_c"][1]))
    # the decompressed proof still verifies under the fixture-pinned oracle and the product verifier
    rebuilt = {"pi_a": [str(a[0]), str(a[1]), "1"], "pi_b": [[str(b[0][0]), str(b[0][1])], [str(b[1][0]), str(b[1][1])], ["1", "0"]],
               "pi_c": [str(c[0]), str(c[1]), "1"], "protocol": "groth16", "curve": "bn128"}
    assert bn254.groth16_verify(vkey, public, rebuilt) and z.verify(vkey, public, rebuilt)
    # flag semantics: the top bit says "y is the larger of (y, -y)"; flipping it yields the negated point
    flipped = bytearray(blob[:32]); flipped[31] ^= V.FLAG_Y_NEGATIVE
    assert V.ark_g1_decompress(bytes(flipped)) == (a[0], V.Q - a[1])
    # known images: the G1 generator (1, 2) has the smaller y; infinity is the 0x40 flag alone
    assert V.ark_g1_compressed((1, 2)) == bytes([1]) + bytes(31)
    assert V.ark_g1_compressed((1, V.Q - 2))[31] == 0x80 | ((1).to_bytes(32, "little")[31])
    assert V.ark_g1_compressed(None) == bytes(31) + bytes([0x40]) and V.ark_g1_decompress(bytes(31) + bytes([0x40])) is None
    g2 = bn254.G2_GEN
    assert V.ark_g2_decompress(V.ark_g2_compressed(g2)) == g2
    assert V.ark_g2_decompress(V.ark_g2_compressed(bn254.g2_neg(g2))) == bn254.g2_neg(g2)


def test_ark_public_inputs_and_vkey():
    pub = V.ark_public_inputs_compressed(public)
    assert len(pub) == 32 * len(public) == 96
    assert [int.from_bytes(pub[32 * i:32 * i + 32], "little") for i in range(3)] == [int(s) for s in public]
    args = V.rust_verifier_arguments(proof, public)
    assert bytes(args["PROOF"]) == V.ark_proof_compressed(proof) and bytes(args["PUBLIC_INPUTS"]) == pub
    vk = V.ark_vkey_compressed(vkey)
    n_ic = len(vkey["IC"])
    assert len(vk) == 32 + 3 * 64 + 8 + 32 * n_ic and int.from_bytes(vk[224:232], "little") == n_ic == vkey["nPublic"] + 1
    assert V.ark_g1_decompress(vk[:32]) == (int(vkey["vk_alpha_1"][0]), int(vkey["vk_alpha_1"][1]))
    for k, name in enumerate(("vk_beta_2", "vk_gamma_2", "vk_delta_2")):
        p = V.ark_g2_decompress(vk[32 + 64 * k:96 + 64 * k])
        assert p == ((int(vkey[name][0][0]), int(vkey[name][0][1])), (int(vkey[name][1][0]), int(vkey[name][1][1])))
    for i in range(n_ic):
        assert V.ark_g1_decompress(vk[232 + 32 * i:264 + 32 * i]) == (int(vkey["IC"][i][0]), int(vkey["IC"][i][1]))
    import pytest
    with pytest.raises(ValueError):
        V.ark_public_inputs_compressed([str(V.R)])


def test_solidity_calldata_and_packing():
    cd = V.solidity_calldata(proof, public)
    words = json.loads("[" + cd + "]")
    assert [int(w, 16) for w in words[0]] == [int(proof["pi_a"][0]), int(proof["pi_a"][1])]
    assert [[int(w, 16) for w in row] for row in words[1]] == [[int(proof["pi_b"][0][1]), int(proof["pi_b"][0][0])],
                                                               [int(proof["pi_b"][1][1]), int(proof["pi_b"][1][0])]]
    assert [int(w, 16) for w in words[3]] == [int(s) for s in public] and all(len(w) == 66 for w in words[0] + words[2] + words[3])
    # public.json[1] of the fixture is PackBytes("zktestemail") - the same 31-byte little-endian packing the contracts use
    assert V.pack_fields_array(b"zktestemail", 21) == [int(public[1])]
    assert V.unpack_fields_array([int(public[1])], 21) == b"zktestemail"
    rnd = random.Random(3)
    for n, pad in ((0, 1), (31, 31), (32, 62), (40, 93), (100, 100)):
        data = bytes(rnd.randrange(1, 256) for _ in range(n))
        f = V.pack_fields_array(data, pad)
        assert len(f) == (pad + 30) // 31 and all(x < (1 << 248) for x in f) and V.unpack_fields_array(f, pad) == data
    assert V.pack_bool(True) == [1] and V.unpack_bool([1]) and not V.unpack_bool(V.pack_bool(False))
    import pytest
    with pytest.raises(ValueError, match="InvalidDataLength"):
        V.pack_fields_array(b"abcd", 3)
