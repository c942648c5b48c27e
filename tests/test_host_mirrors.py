"""Host-side mirrors of @zk-email/helpers (input side of the boundary), pinned on what the reference's tests and
fixtures pin offline: the canonicalisation / hashing of packages/circuits/tests/test-emails/test.eml, the
selector scenarios of packages/helpers/tests/input-generators.test.ts:39-63 (fed from a locally built DKIM result because
the reference fetches the key from DNS), sha256Pad / generatePartialSHA semantics and their error strings."""
import base64
import hashlib
import os
import re
import pytest

import zkemail_b200 as z
from zkemail_b200 import dkim
from zkemail_b200.dkim import DKIMVerificationResult

EMAILS = os.path.join(os.path.dirname(__file__), "golden", "emails")


def _fake_dkim_result(raw: bytes) -> DKIMVerificationResult:
    parsed, body = dkim.split_message(raw)
    line = [l for k, l in parsed if k == "dkim-signature"][0]
    tags = dkim.parse_tag_list(re.sub(r"\r?\n[ \t]*", " ", line.decode("latin-1").split(":", 1)[1]))
    hc, bc = (tags.get("c", "simple/simple").split("/") + ["simple"])[:2]
    canon_body = dkim.relaxed_body(body) if bc == "relaxed" else dkim.simple_body(body)
    headers = dkim.signed_header_bytes(parsed, line, tags["h"], hc)
    return DKIMVerificationResult(publicKey=(1 << 2047) | 12345, signature=(1 << 2040) | 999, headers=headers, body=canon_body,
                                  bodyHash=re.sub(r"\s+", "", tags["bh"]), signingDomain=tags["d"], selector=tags["s"],
                                  algo="rsa-sha256", format=tags.get("c", ""), modulusLength=2048)


def test_test_eml_canonicalisation_kats():
    """rsa.test.ts:40-43 message limbs and the bh= tag, reproduced from test.eml (no key needed)."""
    dk = _fake_dkim_result(open(os.path.join(EMAILS, "test.eml"), "rb").read())
    assert len(dk.headers) == 472
    d = int.from_bytes(hashlib.sha256(dk.headers).digest(), "big")
    assert [(d >> (121 * i)) & ((1 << 121) - 1) for i in range(3)] == [
        1156466847851242602709362303526378170, 191372789510123109308037416804949834, 7204]
    assert base64.b64encode(hashlib.sha256(dk.body).digest()).decode() == dk.bodyHash == "7xQMDuoVVU4m0W0WRVSrVXMeGSIASsnucK9dJsrc+vU="


def test_body_hashes_of_helper_fixtures():
    for name in ("email-good.eml", "email-good-large.eml"):
        dk = _fake_dkim_result(open(os.path.join(EMAILS, name), "rb").read())
        assert base64.b64encode(hashlib.sha256(dk.body).digest()).decode() == dk.bodyHash, name


def test_inputs_shape_and_ignore_body_hash():       # input-generators.test.ts:9-37
    dk = _fake_dkim_result(open(os.path.join(EMAILS, "email-good.eml"), "rb").read())
    inputs = z.generate_email_verifier_inputs_from_dkim_result(dk)
    for k in ("emailHeader", "pubkey", "signature", "precomputedSHA", "emailBody", "emailBodyLength", "bodyHashIndex"):
        assert k in inputs
    assert len(inputs["emailHeader"]) == 1024 and len(inputs["emailBody"]) == 1536 and len(inputs["pubkey"]) == 17
    assert all(isinstance(v, str) for v in inputs["emailHeader"])
    inputs = z.generate_email_verifier_inputs_from_dkim_result(dk, {"ignoreBodyHashCheck": True})
    for k in ("precomputedSHA", "emailBody", "emailBodyLength", "bodyHashIndex"):
        assert k not in inputs


def test_sha_precompute_selector():                  # input-generators.test.ts:39-53
    dk = _fake_dkim_result(open(os.path.join(EMAILS, "email-good-large.eml"), "rb").read())
    inputs = z.generate_email_verifier_inputs_from_dkim_result(dk, {"shaPrecomputeSelector": "thousands"})
    body = bytes(int(b) for b in inputs["emailBody"]).decode("latin-1")
    assert body.startswith("h hundreds of thousands of blocks.")
    # the midstate really is the SHA-256 state after the cut-off prefix
    pre = bytes(int(b) for b in inputs["precomputedSHA"])
    cut = len(dk.body) - (int(inputs["emailBodyLength"]) - (int(inputs["emailBodyLength"]) - 0)) if False else None
    padded, plen = z.sha256_pad(dk.body, max(1536, ((len(dk.body) + 63 + 65) // 64) * 64))
    cutoff = plen - int(inputs["emailBodyLength"])
    assert cutoff % 64 == 0 and pre == z.partial_sha(padded, cutoff)


def test_invalid_selector_message():                 # input-generators.test.ts:55-63
    dk = _fake_dkim_result(open(os.path.join(EMAILS, "email-good.eml"), "rb").read())
    with pytest.raises(ValueError, match='SHA precompute selector "Bla Bla" not found in cleaned body'):
        z.generate_email_verifier_inputs_from_dkim_result(dk, {"shaPrecomputeSelector": "Bla Bla"})


def test_sha256_pad_and_partial_sha():
    for n in (0, 1, 55, 56, 63, 64, 119, 120, 1000):
        msg = bytes((i * 7 + 1) & 0xFF for i in range(n))
        padded, plen = z.sha256_pad(msg, 1536)
        assert len(padded) == 1536 and plen % 64 == 0 and plen >= n + 9
        assert padded[:n] == msg and padded[n] == 0x80 and set(padded[plen:]) <= {0}
        assert int.from_bytes(padded[plen - 8:plen], "big") == 8 * n
        # hashing the padded blocks from the IV gives the digest: partial_sha over everything == final state
        st = z.partial_sha(padded, plen)
        assert st == hashlib.sha256(msg).digest()
    with pytest.raises(AssertionError, match="Padding to max length did not complete properly"):
        z.sha256_pad(b"x" * 100, 64)


def test_remaining_body_too_long_message():
    body = b"a" * 3000
    padded, plen = z.sha256_pad(body, 3072)
    with pytest.raises(ValueError, match=r"Remaining body 3072 after the selector is longer than max \(1536\)"):
        z.generate_partial_sha(padded, plen, None, 1536)


def test_to_circom_bigint_bytes():
    x = (1 << 2047) + 0xDEADBEEF
    limbs = z.to_circom_bigint_bytes(x)
    assert len(limbs) == 17 and sum(int(l) << (121 * i) for i, l in enumerate(limbs)) == x
    assert all(int(l) < (1 << 121) for l in limbs)


def test_dkim_verify_roundtrip_and_failures():
    key = z.synthetic.generate_key()
    email = z.synthetic.make_signed_email(3, key)
    rec = z.synthetic.key_record(key)
    dk = z.verify_dkim_signature(email, resolver=lambda n, t: [rec])
    assert dk.signingDomain == "example.com" and dk.modulusLength == 2048 and len(dk.body) == 1024
    assert pow(dk.signature, 65537, dk.publicKey) & ((1 << 256) - 1) == int.from_bytes(hashlib.sha256(dk.headers).digest(), "big")
    with pytest.raises(ValueError, match="DKIM signature not found for domain other.org"):
        z.verify_dkim_signature(email, domain="other.org", resolver=lambda n, t: [rec])
    tampered = email.replace(b"Subject: synthetic", b"Subject: Synthetic")
    with pytest.raises(ValueError, match="bad signature"):
        z.verify_dkim_signature(tampered, resolver=lambda n, t: [rec])
    with pytest.raises(ValueError, match="body hash did not verify"):
        z.verify_dkim_signature(email + b"extra line\r\n", resolver=lambda n, t: [rec])
    with pytest.raises(ValueError, match="DNS failure"):
        z.verify_dkim_signature(email)      # offline default resolver
