"""EmailVerifier integration scenarios of /root/reference/packages/circuits/tests/email-verifier.test.ts:34-207 on the
test circuit EmailVerifier(640, 768, 121, 17, 0, 0, 0, 0) with `public [pubkey]`
(tests/test-circuits/email-verifier-test.circom:5), evaluated by the CPU oracle.  The reference uses a real
icloud.com email whose key comes from DNS; the same scenarios run here on a synthetic self-signed email."""
import hashlib
import pytest
from zkemail_b200 import (Circuit, generate_email_verifier_inputs_from_dkim_result, synthetic, verify_dkim_signature)
from zkutil import oracle_witness, assert_out, AssertFailed
from oracle import poseidon as oposeidon


@pytest.fixture(scope="module")
def setup():
    c = Circuit("EmailVerifier", [640, 768, 121, 17, 0, 0, 0, 0, 1])
    key = synthetic.generate_key()
    email = synthetic.make_signed_email(1, key, body_len=512)
    dk = verify_dkim_signature(email, resolver=lambda n, t: [synthetic.key_record(key)])
    inputs = generate_email_verifier_inputs_from_dkim_result(dk, {"maxHeadersLength": 640, "maxBodyLength": 768})
    return c, dk, inputs


def test_accepts_valid_email(setup):
    c, dk, inputs = setup
    w = oracle_witness(c, inputs)
    digest = hashlib.sha256(dk.headers).digest()
    assert_out(w, {"shaHi": int.from_bytes(digest[:16], "big"), "shaLo": int.from_bytes(digest[16:], "big")})
    # email-verifier.test.ts:188-207 - pubkeyHash == poseidonLarge(publicKey, 9, 242)
    assert_out(w, {"pubkeyHash": oposeidon.poseidon_large(dk.publicKey, 9, 242)})
    # witness order: [1, pubkeyHash, shaHi, shaLo, pubkey[17], ...]
    assert w[4:4 + 17] == [int(x) for x in inputs["pubkey"]]


def test_rejects_bad_signature(setup):          # :61-79
    c, _, inputs = setup
    bad = dict(inputs)
    bad["signature"] = [str(int(bad["signature"][0]) ^ 1)] + list(bad["signature"][1:])
    with pytest.raises(AssertFailed, match="Assert Failed"):
        oracle_witness(c, bad)


def test_rejects_tampered_header(setup):        # :81-102
    c, _, inputs = setup
    bad = dict(inputs)
    hdr = list(bad["emailHeader"])
    hdr[10] = str((int(hdr[10]) + 1) % 128)
    bad["emailHeader"] = hdr
    with pytest.raises(AssertFailed):
        oracle_witness(c, bad)


def test_rejects_nonzero_header_padding(setup):  # :104-121
    c, _, inputs = setup
    bad = dict(inputs)
    hdr = list(bad["emailHeader"])
    hdr[int(bad["emailHeaderLength"]) + 3] = "1"
    bad["emailHeader"] = hdr
    with pytest.raises(AssertFailed, match="AssertZeroPadding"):
        oracle_witness(c, bad)


def test_rejects_tampered_body(setup):          # :123-144
    c, _, inputs = setup
    bad = dict(inputs)
    body = list(bad["emailBody"])
    body[5] = str((int(body[5]) + 1) % 128)
    bad["emailBody"] = body
    with pytest.raises(AssertFailed):
        oracle_witness(c, bad)


def test_rejects_nonzero_body_padding(setup):   # :146-163
    c, _, inputs = setup
    bad = dict(inputs)
    body = list(bad["emailBody"])
    body[int(bad["emailBodyLength"]) + 1] = "1"
    bad["emailBody"] = body
    with pytest.raises(AssertFailed, match="AssertZeroPadding"):
        oracle_witness(c, bad)


def test_rejects_wrong_body_hash_index(setup):  # :165-186
    c, _, inputs = setup
    bad = dict(inputs)
    bad["bodyHashIndex"] = str(int(bad["bodyHashIndex"]) + 1)
    with pytest.raises(AssertFailed):
        oracle_witness(c, bad)


# ---- circuit variants: email-verifier-no-body.test.ts:34-47, -with-header-mask, -with-body-mask.test.ts:33-58 ----------
def _synthetic(maxh=640, maxb=768, **extra):
    key = synthetic.generate_key()
    email = synthetic.make_signed_email(2, key, body_len=300)
    dk = verify_dkim_signature(email, resolver=lambda n, t: [synthetic.key_record(key)])
    params = {"maxHeadersLength": maxh, "maxBodyLength": maxb}
    params.update(extra)
    return dk, generate_email_verifier_inputs_from_dkim_result(dk, params)


def test_no_body_variant():
    c = Circuit("EmailVerifier", [640, 768, 121, 17, 1, 0, 0, 0, 1])
    dk, inputs = _synthetic(ignoreBodyHashCheck=True)
    assert "emailBody" not in inputs
    w = oracle_witness(c, inputs)
    digest = hashlib.sha256(dk.headers).digest()
    assert_out(w, {"shaHi": int.from_bytes(digest[:16], "big"), "shaLo": int.from_bytes(digest[16:], "big")})


def test_header_mask_variant():
    c = Circuit("EmailVerifier", [640, 768, 121, 17, 0, 1, 0, 0, 1])
    mask = [1 if i % 3 == 0 else 0 for i in range(640)]
    dk, inputs = _synthetic(enableHeaderMasking=True, headerMask=mask)
    w = oracle_witness(c, inputs)
    hdr = [int(x) for x in inputs["emailHeader"]]
    assert_out(w, {"maskedHeader": [h * m for h, m in zip(hdr, mask)]})
    bad = dict(inputs)
    bad["headerMask"] = [2] + mask[1:]              # AssertBit
    with pytest.raises(AssertFailed):
        oracle_witness(c, bad)


def test_body_mask_variant():
    c = Circuit("EmailVerifier", [640, 768, 121, 17, 0, 0, 1, 0, 1])
    mask = [1] * 10 + [0] * 758
    dk, inputs = _synthetic(enableBodyMasking=True, bodyMask=mask)
    w = oracle_witness(c, inputs)
    body = [int(x) for x in inputs["emailBody"]]
    assert_out(w, {"maskedBody": body[:10] + [0] * 758})


# ---- the compact regex shape inside EmailVerifier (regex.cpp, template parameter 10 = 1) -------------------------------
def test_compact_regex_shape_variant(setup):
    """Same statement, smaller circuit: identical public outputs on the same email, the tamper scenarios that go through
    the body-hash regex still reject, and the default circuit drops below 2^21 constraints."""
    _, dk, inputs = setup
    c = Circuit("EmailVerifier", [640, 768, 121, 17, 0, 0, 0, 0, 1, 1])
    ref = Circuit("EmailVerifier", [640, 768, 121, 17, 0, 0, 0, 0, 1, 0])
    assert c.info.n_constraints < ref.info.n_constraints - 300_000
    w, w0 = oracle_witness(c, inputs), oracle_witness(ref, inputs)
    n_pub = 1 + c.info.n_public
    assert w[:n_pub] == w0[:n_pub]
    digest = hashlib.sha256(dk.headers).digest()
    assert_out(w, {"shaHi": int.from_bytes(digest[:16], "big"), "shaLo": int.from_bytes(digest[16:], "big")})
    bad = dict(inputs)
    bad["bodyHashIndex"] = str(int(inputs["bodyHashIndex"]) + 1)           # email-verifier.test.ts:165-186
    with pytest.raises(AssertFailed):
        oracle_witness(c, bad)
    hdr = list(inputs["emailHeader"])
    pos = int(inputs["bodyHashIndex"]) - 2                                   # the '=' of "bh=": the regex no longer matches
    hdr[pos] = str(ord("-"))
    with pytest.raises(AssertFailed):
        oracle_witness(c, dict(inputs, emailHeader=hdr))
    full = Circuit("EmailVerifier", [1024, 1536, 121, 17, 0, 0, 0, 0, 0, 1])
    assert full.info.domain_log2 == 21 and Circuit("EmailVerifier", [1024, 1536, 121, 17]).info.domain_log2 == 22
