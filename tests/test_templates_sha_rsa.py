"""SHA-256 and RSA template tests restated from /root/reference/packages/circuits/tests/{sha,rsa}.test.ts."""
import hashlib
import pytest
from zkemail_b200 import Circuit, sha256_pad, partial_sha, to_circom_bigint_bytes
from zkutil import oracle_witness, assert_out, AssertFailed


def _bits_msb(digest: bytes):
    return [(b >> (7 - j)) & 1 for b in digest for j in range(8)]


@pytest.fixture(scope="module")
def sha640():
    return Circuit("Sha256Bytes", [640])


# sha.test.ts:25-43 - Sha256Bytes(640) on "0", "hello world", ""
@pytest.mark.parametrize("msg", [b"0", b"hello world", b""])
def test_sha256_bytes_640(sha640, msg):
    padded, padded_len = sha256_pad(msg, 640)
    w = oracle_witness(sha640, {"paddedIn": list(padded), "paddedInLength": padded_len})
    assert_out(w, {"out": _bits_msb(hashlib.sha256(msg).digest())})


def test_sha256_bytes_multiblock(sha640):
    msg = bytes(range(256)) + b"x" * 200          # 8 blocks after padding
    padded, padded_len = sha256_pad(msg, 640)
    assert padded_len == 512
    w = oracle_witness(sha640, {"paddedIn": list(padded), "paddedInLength": padded_len})
    assert_out(w, {"out": _bits_msb(hashlib.sha256(msg).digest())})


def test_sha256_bad_length_rejected(sha640):
    padded, padded_len = sha256_pad(b"hello world", 640)
    with pytest.raises(AssertFailed):            # not a multiple of 64 bytes
        oracle_witness(sha640, {"paddedIn": list(padded), "paddedInLength": padded_len + 1})
    with pytest.raises(AssertFailed):            # zero blocks: ItemAtIndex finds no index
        oracle_witness(sha640, {"paddedIn": list(padded), "paddedInLength": 0})


def test_sha256_bytes_partial():
    c = Circuit("Sha256BytesPartial", [192])
    msg = bytes((7 * i + 3) & 0xFF for i in range(300))
    padded, padded_len = sha256_pad(msg, 384)      # 320 bytes of SHA-padded data
    cut = 128
    pre = partial_sha(padded, cut)
    rest = padded[cut:cut + 192]
    w = oracle_witness(c, {"paddedIn": list(rest), "paddedInLength": padded_len - cut, "preHash": list(pre)})
    assert_out(w, {"out": _bits_msb(hashlib.sha256(msg).digest())})


# ---- RSA ------------------------------------------------------------------------------------------
MESSAGE = ["1156466847851242602709362303526378170", "191372789510123109308037416804949834", "7204"] + ["0"] * 14
SIG_1024 = 102386562682221859025549328916727857389789009840935140645361501981959969535413501251999442013082353139290537518086128904993091119534674934202202277050635907008004079788691412782712147797487593510040249832242022835902734939817209358184800954336078838331094308355388211284440290335887813714894626653613586546719
PUB_1024 = 106773687078109007595028366084970322147907086635176067918161636756354740353674098686965493426431314019237945536387044259034050617425729739578628872957481830432099721612688699974185290306098360072264136606623400336518126533605711223527682187548332314997606381158951535480830524587400401856271050333371205030999


@pytest.fixture(scope="module")
def rsa_circuit():
    return Circuit("RSAVerifier65537", [121, 17])


# rsa.test.ts:64-103
def test_rsa_1024_kat(rsa_circuit):
    oracle_witness(rsa_circuit, {"signature": to_circom_bigint_bytes(SIG_1024), "modulus": to_circom_bigint_bytes(PUB_1024),
                                 "message": MESSAGE})


# rsa.test.ts:105-143 (message limb 0 incremented)
def test_rsa_bad_message_rejected(rsa_circuit):
    bad = list(MESSAGE)
    bad[0] = str(int(bad[0]) + 1)
    with pytest.raises(AssertFailed, match="Assert Failed"):
        oracle_witness(rsa_circuit, {"signature": to_circom_bigint_bytes(SIG_1024),
                                     "modulus": to_circom_bigint_bytes(PUB_1024), "message": bad})


# rsa.test.ts:27-62 uses the icloud key (DNS, not in the repo): same scenario with a locally generated 2048-bit key
def test_rsa_2048_synthetic(rsa_circuit):
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric import padding
    from zkemail_b200 import synthetic
    key = synthetic.generate_key(2048)
    data = b"signed header bytes"
    sig = int.from_bytes(key.sign(data, padding.PKCS1v15(), hashes.SHA256()), "big")
    digest = int.from_bytes(hashlib.sha256(data).digest(), "big")
    n = key.public_key().public_numbers().n
    oracle_witness(rsa_circuit, {"signature": to_circom_bigint_bytes(sig), "modulus": to_circom_bigint_bytes(n),
                                 "message": to_circom_bigint_bytes(digest)})
    with pytest.raises(AssertFailed):
        oracle_witness(rsa_circuit, {"signature": to_circom_bigint_bytes(sig ^ 2), "modulus": to_circom_bigint_bytes(n),
                                     "message": to_circom_bigint_bytes(digest)})
    with pytest.raises(AssertFailed):            # signature >= modulus
        oracle_witness(rsa_circuit, {"signature": to_circom_bigint_bytes(sig + n), "modulus": to_circom_bigint_bytes(n),
                                     "message": to_circom_bigint_bytes(digest)})
