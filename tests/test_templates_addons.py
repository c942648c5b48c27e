"""Add-on templates of the reference, each on the reference's own test vectors; witnesses come
from the CPU oracle, which also checks every constraint (circom_tester's calculateWitness + checkConstraints):

  CheckSubstringMatch(32)              tests/check-substring-match.test.ts:21-108      utils/array.circom:193-217
  CountSubstringOccurrences(1024,128)  tests/count-substring-occurrences.test.ts:23-105 utils/array.circom:226-253
  RevealSubstring(256,16,1)            tests/reveal-substring.test.ts:21-187           helpers/reveal-substring.circom
  CleanEmailAddress(32)                tests/clean-email-address.test.ts:24-122        utils/email.circom:16-139
  SplitBytesToWords(256,121,17)        tests/split-bytes-to-words.test.ts:22-33        utils/bytes.circom:125-149
  EmailNullifier(121,17)               helpers/email-nullifier.circom (no reference test: checked against the Poseidon oracle)
(paths relative to /root/reference/packages/circuits)."""
import random
import pytest

import zkemail_b200 as z
from zkutil import oracle_witness, AssertFailed
from oracle import poseidon as pos


def pad(xs, n):
    return list(xs) + [0] * (n - len(xs))


def ascii_arr(s, n):
    return pad(list(s.encode()), n)


# ------------------------------------------------------------------ CheckSubstringMatch(32)
@pytest.fixture(scope="module")
def csm():
    return z.Circuit("CheckSubstringMatch", [32])


@pytest.mark.parametrize("inp,sub,expected", [
    ([1, 2, 3, 4, 5], [1, 2, 3], 1),                       # substring at the beginning
    ([1, 2, 3, 4, 5], [1, 2, 4], 0),                       # different
    (list(range(1, 33)), list(range(1, 33)), 1),           # full length
    ([9, 1, 2, 3, 4], [1, 2, 3], 0),                       # not at the beginning
    ([1, 2, 3, 4, 5], [1], 1),                             # single element
    ([], [1, 2, 3], 0),                                    # all-zero input
    ([1, 2, 3, 4, 5], [1, 2, 3, 4, 5, 6], 0),              # longer than the non-zero input
])
def test_check_substring_match(csm, inp, sub, expected):
    w = oracle_witness(csm, {"in": pad(inp, 32), "substring": pad(sub, 32)})
    assert w.values("isMatch") == [expected]


def test_check_substring_match_rejects_zero_first_element(csm):
    with pytest.raises(AssertFailed):
        oracle_witness(csm, {"in": pad([1, 2, 3, 4, 5], 32), "substring": pad([0, 2, 3], 32)})


# ------------------------------------------------------------------ CountSubstringOccurrences(1024, 128)
@pytest.fixture(scope="module")
def cso():
    return z.Circuit("CountSubstringOccurrences", [1024, 128])


@pytest.mark.parametrize("inp,sub,expected", [
    (pad([1, 2, 3, 4, 5], 1024), [1, 2, 3], 1),
    (pad([1, 2, 3, 4, 1, 2, 3, 5, 1, 2, 3], 1024), [1, 2, 3], 3),
    (pad([1, 2, 4, 5, 6], 1024), [1, 2, 3], 0),
    (pad([1, 1, 1, 2, 1, 1], 1024), [1, 1], 3),            # overlapping
    ([1, 2, 3, 4] * 256, [1, 2, 3, 4], 256),               # the whole input is the pattern repeated
    (pad([1, 2, 1, 3, 1, 4, 1], 1024), [1], 4),
    ([0] * 1021 + [1, 2, 3], [1, 2, 3], 1),                # at the end of the input
])
def test_count_substring_occurrences(cso, inp, sub, expected):
    w = oracle_witness(cso, {"in": inp, "substring": pad(sub, 128)})
    assert w.values("count") == [expected]


def test_count_substring_occurrences_rejects_empty_substring(cso):
    with pytest.raises(AssertFailed):
        oracle_witness(cso, {"in": pad([1, 2, 3, 4, 5], 1024), "substring": [0] * 128})


# ------------------------------------------------------------------ RevealSubstring(256, 16, 1)
@pytest.fixture(scope="module")
def rs():
    return z.Circuit("RevealSubstring", [256, 16, 1])


RS_IN = pad([(i % 255) + 1 for i in range(100)], 256)


@pytest.mark.parametrize("start,length,expected", [
    (50, 5, [51, 52, 53, 54, 55]),
    (0, 5, [1, 2, 3, 4, 5]),
    (95, 5, [96, 97, 98, 99, 100]),
    (50, 1, [51]),
    (0, 16, list(range(1, 17))),
    (0, 3, [1, 2, 3]),
])
def test_reveal_substring(rs, start, length, expected):
    w = oracle_witness(rs, {"in": RS_IN, "substringStartIndex": start, "substringLength": length})
    assert w.values("substring") == pad(expected, 16)
    assert [w[1 + i] for i in range(16)] == pad(expected, 16)      # witness.slice(1, 17) of the reference test


@pytest.mark.parametrize("start,length", [(256, 1), (257, 1), (0, 16), (0, 17), (250, 7)])
def test_reveal_substring_rejects(rs, start, length):
    # out-of-range index / length, and (0, 16) on an all-ones input: the substring is not unique
    with pytest.raises(AssertFailed):
        oracle_witness(rs, {"in": [1] * 256, "substringStartIndex": start, "substringLength": length})


# ------------------------------------------------------------------ CleanEmailAddress(32)
@pytest.mark.parametrize("encoded,decoded,valid", [
    ("shreyas.londhe+alias@gmail.com", "shreyaslondhe@gmail.com", 1),
    ("shreyas.londhe+alias@gmail.com", "shreyaslondhe@yahoo.com", 0),
    ("shreyas.londhe@gmail.com", "shreyaslondhe@gmail.com", 1),
    ("shreyaslondhe+test@gmail.com", "shreyaslondhe@gmail.com", 1),
    ("shreyas.londhe.test@gmail.com", "shreyaslondhetest@gmail.com", 1),
    ("shs.loe+test.alias+123@gmail.com", "shsloe@gmail.com", 1),
    ("shreyaslondhe@gmail.com", "shreyaslondhe@gmail.com", 1),
])
def test_clean_email_address(encoded, decoded, valid):
    c = z.Circuit("CleanEmailAddress", [32])
    w = oracle_witness(c, {"encoded": ascii_arr(encoded, 32), "decoded": ascii_arr(decoded, 32)})
    assert w.values("isValid") == [valid]


# ------------------------------------------------------------------ SplitBytesToWords(256, 121, 17)
def test_split_bytes_to_words_matches_bigint_to_chunked_bytes():
    c = z.Circuit("SplitBytesToWords", [256, 121, 17])
    rng = random.Random(5)
    for _ in range(3):
        data = bytes(rng.randrange(256) for _ in range(256))
        words = z.bigint_to_chunked_bytes(int.from_bytes(data, "big"), 121, 17)
        w = oracle_witness(c, {"in": list(data)})
        assert w.values("out") == [int(x) for x in words]


# ------------------------------------------------------------------ EmailNullifier(121, 17)
def test_email_nullifier_is_poseidon_of_poseidon_large():
    c = z.Circuit("EmailNullifier", [121, 17])
    rng = random.Random(9)
    sig = [rng.randrange(1 << 121) for _ in range(17)]
    w = oracle_witness(c, {"signature": sig})
    merged = [sig[2 * i] + (sig[2 * i + 1] << 121) for i in range(8)] + [sig[16]]   # PoseidonLarge (utils/hash.circom:30-36)
    assert w.values("out") == [pos.poseidon([pos.poseidon(merged)])]
