"""Front-end soundness probe: no internal signal of any gadget may be left unconstrained.

The witness oracle (oracle/zkref_witness.c) and the GPU interpreter both walk the witness program the front end emits, so
"GPU == oracle" cannot see a gadget that computes a value by hint and forgets the constraint.  This test looks at the
front end from the other side - the exported R1CS and one satisfying witness, nothing else (tests/r1cs_probe.py): every
non-input variable is perturbed alone and some constraint must object.  The only tolerated freedom is the inverse hint
of circomlib's IsZero at in = 0 (`inv <-- in != 0 ? 1/in : 0; out <== -in*inv + 1; in*out === 0`: at in = 0 the
constraints read out = 1 whatever inv is), which the reference's circuits share.
Scenarios follow the reference's own circuit tests (packages/circuits/tests/*.test.ts) for the inputs.
"""
import hashlib
import random

import pytest
import zkemail_b200 as z
from zkemail_b200.sha_utils import sha256_pad
from zkutil import oracle_witness
from r1cs_probe import probe, scopes_of


def check(circuit, inputs):
    w = oracle_witness(circuit, inputs)                       # calculateWitness + checkConstraints
    values = [w[i] for i in range(circuit.info.n_vars)]
    free, two, unmentioned = probe(circuit, values)
    assert unmentioned == [], f"variables in no constraint: {unmentioned[:8]}"
    assert two == {}, f"two-valued variables: {scopes_of(circuit, two)}"
    classes = scopes_of(circuit, free)
    assert set(classes) <= {"IsZero"}, f"unconstrained variables outside IsZero: {classes}"
    return classes


def _pad(xs, n):
    return list(xs) + [0] * (n - len(xs))


def test_sha256_templates():
    msg = bytes(range(100))
    padded, plen = sha256_pad(msg, 192)
    check(z.Circuit("Sha256Bytes", [192]), {"paddedIn": list(padded), "paddedInLength": plen})
    # Sha256BytesPartial from the midstate after the first block (sha.test.ts partial scenario)
    from zkemail_b200.sha_utils import partial_sha
    pre = partial_sha(padded, 64)
    check(z.Circuit("Sha256BytesPartial", [128]), {"paddedIn": list(padded[64:192]), "paddedInLength": plen - 64, "preHash": list(pre)})


def test_rsa_and_fpmul():
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric import padding
    key = z.synthetic.generate_key(2048)
    data = b"signed header bytes"
    sig = int.from_bytes(key.sign(data, padding.PKCS1v15(), hashes.SHA256()), "big")
    digest = int.from_bytes(hashlib.sha256(data).digest(), "big")
    n = key.public_key().public_numbers().n
    limbs = z.to_circom_bigint_bytes
    check(z.Circuit("RSAVerifier65537", [121, 17]), {"signature": limbs(sig), "modulus": limbs(n), "message": limbs(digest)})
    rng = random.Random(5)
    p = rng.getrandbits(242) | (1 << 241) | 1
    a, b = rng.randrange(p), rng.randrange(p)
    split = lambda x: [(x >> (121 * i)) & ((1 << 121) - 1) for i in range(2)]
    check(z.Circuit("FpMul", [121, 2]), {"a": split(a), "b": split(b), "p": split(p)})


def test_regex_base64_reveal():
    n = 128
    hdr = b"to:a@b.c\r\ndkim-signature:v=1; a=rsa-sha256; bh=" + b"QUJD" * 10 + b"QUI=; b=xyz"
    bh = check(z.Circuit("BodyHashRegex", [n, 0]), {"msg": _pad(hdr, n)})
    assert bh.get("IsZero", 0) > 0                            # character tests that hit their constant
    check(z.Circuit("TwitterResetRegex", [64, 0]), {"msg": _pad(b"x email was meant for @zk_mail.", 64)})
    # the compact shape (one-hot live sets, nibble one-hots): nothing is free at all on a matching input
    assert check(z.Circuit("BodyHashRegex", [n, 1]), {"msg": _pad(hdr, n)}) == {}
    assert check(z.Circuit("TwitterResetRegex", [64, 1]), {"msg": _pad(b"x email was meant for @zk_mail.", 64)}) == {}
    assert set(check(z.Circuit("BodyHashRegex", [n, 1]), {"msg": _pad(b"subject: nothing here", n)})) <= {"IsZero"}
    b64 = b"QUJD" * 10 + b"QUI="
    check(z.Circuit("Base64Decode", [32]), {"in": list(b64)})
    arr = [0] * 10 + list(b"hello") + [0] * 19
    check(z.Circuit("SelectRegexReveal", [34, 8]), {"in": arr, "startIndex": 10})
    check(z.Circuit("PackRegexReveal", [64, 21]), {"in": [0] * 17 + list(b"zk_mail") + [0] * 40, "startIndex": 17})


def test_hashes_and_small_gadgets():
    rng = random.Random(9)
    check(z.Circuit("PoseidonLarge", [121, 17]), {"in": [rng.getrandbits(121) for _ in range(17)]})
    check(z.Circuit("PoseidonModular", [37]), {"in": [rng.getrandbits(200) for _ in range(37)]})
    check(z.Circuit("PackBits", [256, 128]), {"in": [rng.getrandbits(1) for _ in range(256)]})
    check(z.Circuit("ByteMask", [10]), {"in": list(range(1, 11)), "mask": [1, 0] * 5})
    check(z.Circuit("AssertZeroPadding", [32]), {"in": _pad([7] * 9, 32), "startIndex": 9})
    check(z.Circuit("ItemAtIndex", [16]), {"in": list(range(100, 116)), "index": 5})
    check(z.Circuit("VarShiftLeft", [32, 8]), {"in": list(range(32)), "shift": 11})
    check(z.Circuit("SplitBytesToWords", [256, 121, 17]), {"in": [rng.getrandbits(8) for _ in range(256)]})
    check(z.Circuit("EmailNullifier", [121, 17]), {"signature": [rng.getrandbits(121) for _ in range(17)]})


def test_addon_templates():
    text = b"hello=\r\n world=\r\n!"
    decoded = text.replace(b"=\r\n", b"")
    check(z.Circuit("RemoveSoftLineBreaks", [32]), {"encoded": _pad(text, 32), "decoded": _pad(decoded, 32)})
    check(z.Circuit("CheckSubstringMatch", [32]), {"in": _pad([1, 2, 3, 4, 5], 32), "substring": _pad([1, 2, 3], 32)})
    check(z.Circuit("CountSubstringOccurrences", [64, 8]), {"in": _pad(b"abcabcab", 64), "substring": _pad(b"abc", 8)})
    check(z.Circuit("RevealSubstring", [64, 16, 1]), {"in": list(range(1, 65)), "substringStartIndex": 5, "substringLength": 7})
    check(z.Circuit("SelectSubArray", [64, 16]), {"in": list(range(1, 65)), "startIndex": 5, "length": 7})


def test_email_verifier_all_flags():
    """EmailVerifier(640, 768) with header mask, body mask and soft-line-break removal on: every gadget in its context."""
    c = z.Circuit("EmailVerifier", [640, 768, 121, 17, 0, 1, 1, 1, 1])
    key = z.synthetic.generate_key()
    body = b"A quoted-printable body with a soft line =\r\nbreak in the middle and another one right he=\r\nre.\r\n"
    email = z.synthetic.make_signed_email(9, key, body_len=len(body), body_override=body)
    dk = z.verify_dkim_signature(email, resolver=lambda n, t: [z.synthetic.key_record(key)])
    inputs = z.generate_email_verifier_inputs_from_dkim_result(
        dk, {"maxHeadersLength": 640, "maxBodyLength": 768, "removeSoftLineBreaks": True,
             "enableHeaderMasking": True, "headerMask": [i % 2 for i in range(640)],
             "enableBodyMasking": True, "bodyMask": [1] * 20 + [0] * 748})
    classes = check(c, inputs)
    assert classes.get("IsZero", 0) > 0


def test_probe_finds_a_forgotten_constraint():
    """The probe itself: without the row that ties Num2Bits' bits to its input, a bit can flip unnoticed (its booleanity
    row allows exactly the other value) - the probe reports it as two-valued."""
    c = z.Circuit("Num2Bits", [8])
    w = oracle_witness(c, {"in": 0xa5})
    values = [w[i] for i in range(c.info.n_vars)]
    assert probe(c, values) == ({}, {}, [])
    hits = 0
    for row in range(c.info.n_constraints):
        free, two, unmentioned = probe(c, values, drop_rows={row})
        hits += bool(free or two or unmentioned)
    assert hits >= 1
