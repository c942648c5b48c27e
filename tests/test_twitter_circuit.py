"""Proof-of-Twitter circuit (BASELINE configs[3]) and the templates it adds on top of EmailVerifier:
PackBytes (/root/reference/packages/circuits/utils/bytes.circom:28-60), PackRegexReveal (utils/regex.circom:61-77) and
the body regex `email was meant for @(\\w+)` (selector: docs/zk-email-docs/UsageGuide/README.md:84).

Reference pin: the second public signal of the reference's only proof fixture
(packages/rust-verifier/tests/data/proof_of_twitter/public.json) is PackBytes("zktestemail") - the test below reproduces
that field element from a synthetic email carrying the same user name, and signals [1:] of the fixture with its address.
Witnesses come from the CPU oracle; every witness is checked against the R1CS."""
import json
import os
import pytest

import zkemail_b200 as z
from zkutil import oracle_witness, AssertFailed

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "proof_of_twitter", "public.json")
FIXTURE_PUBLIC = [int(x) for x in json.load(open(GOLDEN))]
USERNAME = b"zktestemail"


def test_pack_bytes_matches_reference_fixture_and_layout():
    c = z.Circuit("PackBytes", [21])
    w = oracle_witness(c, {"in": list(USERNAME) + [0] * (21 - len(USERNAME))})
    assert w.values("out") == [FIXTURE_PUBLIC[1]] == [int.from_bytes(USERNAME, "little")]
    c = z.Circuit("PackBytes", [40])                      # two chunks: 31 + 9 bytes, little-endian within a chunk
    data = list(range(1, 41))
    w = oracle_witness(c, {"in": data})
    assert w.values("out") == [int.from_bytes(bytes(data[:31]), "little"), int.from_bytes(bytes(data[31:]), "little")]


def test_pack_regex_reveal():
    c = z.Circuit("PackRegexReveal", [64, 21])
    arr = [0] * 64
    arr[17:17 + len(USERNAME)] = list(USERNAME)
    w = oracle_witness(c, {"in": arr, "startIndex": 17})
    assert w.values("out") == [FIXTURE_PUBLIC[1]]
    with pytest.raises(AssertFailed):                     # start index must sit on the first revealed byte
        oracle_witness(c, {"in": arr, "startIndex": 18})


def test_twitter_reset_regex():
    c = z.Circuit("TwitterResetRegex", [64])
    def run(text):
        msg = list(text.encode()) + [0] * (64 - len(text))
        w = oracle_witness(c, {"msg": msg})
        return w.values("out")[0], bytes(v for v in w.values("reveal0") if v)
    assert run("This email was meant for @zktestemail\r\nbye") == (1, USERNAME)
    assert run("xx email was meant for @a_B9. ") == (1, b"a_B9")
    assert run("This email was meant for nobody")[0] == 0
    assert run("email was meant for @")[0] == 0          # \w+ needs at least one character


@pytest.fixture(scope="module")
def twitter_setup():
    c = z.Circuit("TwitterVerifier", [1024, 1536, 121, 17])
    key = z.synthetic.generate_key()
    em = z.synthetic.make_signed_email(7, key, marker="This email was meant for @" + USERNAME.decode())
    dk = z.verify_dkim_signature(em, resolver=lambda n, t: [z.synthetic.key_record(key)])
    inputs = z.generate_twitter_verifier_inputs_from_dkim_result(dk, FIXTURE_PUBLIC[2])
    return c, inputs


def test_twitter_verifier_public_signals(twitter_setup):
    c, inputs = twitter_setup
    assert c.info.n_public == 3 and c.info.n_outputs == 2 and c.info.n_pub_inputs == 1
    assert [g for g, (_, _, kind) in c.groups.items() if kind == 0] == ["pubkeyHash", "twitterUsername"]
    w = oracle_witness(c, inputs)                          # also checks every constraint
    publics = [w[1 + i] for i in range(3)]
    assert publics[1:] == FIXTURE_PUBLIC[1:]               # [twitterUsername, address] as in the reference's fixture
    assert publics[0] == w.values("pubkeyHash")[0] != 0


def test_twitter_verifier_rejects(twitter_setup):
    c, inputs = twitter_setup
    bad = dict(inputs, twitterUsernameIndex=str(int(inputs["twitterUsernameIndex"]) + 1))
    with pytest.raises(AssertFailed):
        oracle_witness(c, bad)
    body = list(inputs["emailBody"])
    at = int(inputs["twitterUsernameIndex"])
    body[at - 1] = str(ord("#"))                           # "@" -> "#": no match (and a body-hash mismatch)
    with pytest.raises(AssertFailed):
        oracle_witness(c, dict(inputs, emailBody=body))


@pytest.mark.gpu
def test_twitter_verifier_gpu_witness_and_proof(twitter_setup):
    from oracle import bn254
    c, inputs = twitter_setup
    zk = z.Zkey(c, seed=11)
    ctx = z.Context(c, zk, device=0, max_batch=1)
    packed = c.pack_inputs(inputs)
    wt, status = ctx.witness(packed, 1)
    assert status == [-1] and wt == oracle_witness(c, inputs).raw()
    proofs, publics, _ = ctx.prove(1)
    proof, pubs = z.proof_to_json(proofs, publics, c.info.n_public)
    assert [int(x) for x in pubs[1:]] == FIXTURE_PUBLIC[1:]
    assert bn254.groth16_verify(zk.vkey(), pubs, proof) and z.verify(zk.vkey(), pubs, proof)
    tampered = list(pubs); tampered[2] = str(int(pubs[2]) + 1)      # the address is bound to the proof
    assert not bn254.groth16_verify(zk.vkey(), tampered, proof)
