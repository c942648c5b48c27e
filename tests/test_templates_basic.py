"""Template known-answer tests restated from the reference's circuit unit tests:
base64.test.ts, fp-mul.test.ts, pack-bits.test.ts, byte-mask.test.ts, select-regex-reveal.test.ts."""
import pytest
from zkemail_b200 import Circuit
from zkutil import oracle_witness, assert_out, AssertFailed


# /root/reference/packages/circuits/tests/base64.test.ts:20-39
@pytest.mark.parametrize("ch,val", [(65, 0), (90, 25), (97, 26), (122, 51), (48, 52), (57, 61), (43, 62), (47, 63), (61, 0)])
def test_base64_lookup(ch, val):
    c = Circuit("Base64Lookup")
    w = oracle_witness(c, {"in": ch})
    assert_out(w, {"out": val})


# base64.test.ts:42-57
@pytest.mark.parametrize("ch", [64, 91, 96, 123])
def test_base64_lookup_rejects(ch):
    c = Circuit("Base64Lookup")
    with pytest.raises(AssertFailed, match="Assert Failed"):
        oracle_witness(c, {"in": ch})


def test_base64_decode():
    import base64, hashlib
    digest = hashlib.sha256(b"hello").digest()
    enc = base64.b64encode(digest)
    c = Circuit("Base64Decode", [32])
    w = oracle_witness(c, {"in": list(enc)})
    assert_out(w, {"out": list(digest)})


# /root/reference/packages/circuits/tests/fp-mul.test.ts:33-46
def test_fpmul_2_4():
    c = Circuit("FpMul", [2, 4])
    w = oracle_witness(c, {"a": [1, 0, 1, 0], "b": [0, 1, 1, 0], "p": [1, 1, 1, 1]})
    assert_out(w, {"out": [0, 0, 0, 0]})


def test_fpmul_random_121_17():
    import random
    rnd = random.Random(7)
    n, k = 121, 17
    p = rnd.getrandbits(2048) | (1 << 2047) | 1
    a, b = rnd.randrange(p), rnd.randrange(p)
    limbs = lambda x: [(x >> (n * i)) & ((1 << n) - 1) for i in range(k)]
    c = Circuit("FpMul", [n, k])
    w = oracle_witness(c, {"a": limbs(a), "b": limbs(b), "p": limbs(p)})
    assert_out(w, {"out": limbs(a * b % p)})


# /root/reference/packages/circuits/tests/pack-bits.test.ts
def test_pack_bits():
    bits = [1, 0, 1, 1, 0, 0, 1, 0, 1, 1]
    c = Circuit("PackBits", [10, 4])
    w = oracle_witness(c, {"in": bits})
    assert_out(w, {"out": [0b1011, 0b0010, 0b1100]})
    c = Circuit("PackBits", [256, 128])
    import random
    rnd = random.Random(3)
    bits = [rnd.randint(0, 1) for _ in range(256)]
    w = oracle_witness(c, {"in": bits})
    hi = int("".join(map(str, bits[:128])), 2)
    lo = int("".join(map(str, bits[128:])), 2)
    assert_out(w, {"out": [hi, lo]})


# /root/reference/packages/circuits/tests/byte-mask.test.ts:18-29
def test_byte_mask():
    c = Circuit("ByteMask", [10])
    w = oracle_witness(c, {"in": [1, 2, 3, 4, 5, 6, 7, 8, 9, 10], "mask": [1, 0, 1, 0, 1, 0, 1, 0, 1, 0]})
    assert_out(w, {"out": [1, 0, 3, 0, 5, 0, 7, 0, 9, 0]})
    with pytest.raises(AssertFailed):
        oracle_witness(c, {"in": [1] * 10, "mask": [1, 2, 1, 0, 1, 0, 1, 0, 1, 0]})


# /root/reference/packages/circuits/tests/select-regex-reveal.test.ts:22-115 (SelectRegexReveal(34, 8))
def _srr_input(text: bytes, start: int, total=34):
    arr = [0] * total
    arr[start:start + len(text)] = list(text)
    return arr


def test_select_regex_reveal():
    c = Circuit("SelectRegexReveal", [34, 8])
    text = b"zk email"
    w = oracle_witness(c, {"in": _srr_input(text, 5), "startIndex": 5})
    assert_out(w, {"out": list(text)})
    # startIndex pointing at a zero byte
    with pytest.raises(AssertFailed):
        oracle_witness(c, {"in": _srr_input(text, 5), "startIndex": 4})
    # startIndex in the middle of the reveal (byte before it is non-zero)
    with pytest.raises(AssertFailed):
        oracle_witness(c, {"in": _srr_input(text, 5), "startIndex": 6})
    # non-zero data beyond startIndex + maxRevealLen
    with pytest.raises(AssertFailed):
        oracle_witness(c, {"in": _srr_input(b"zk email!", 5), "startIndex": 5})
