"""BodyHashRegex at the template boundary (call site
/root/reference/packages/circuits/email-verifier.circom:126-131, contract in SURVEY A.6): the circuit generated from the
decomposed regex
    (\\r\\n|^)dkim-signature:   ([a-z]+=[^;]+; )+bh=   [a-zA-Z0-9+/=]+ (public)   ;
must set `out` exactly when the regex matches somewhere in the (zero-padded) message and reveal exactly the bytes of the
bh= value.  The zk-regex package is un-vendored, so the pin is the regex itself: Python's `re` is the independent
matcher the circuit's `out` is compared with - on curated headers and on a few hundred random strings over an alphabet
that exercises every part of the automaton.  Every test runs on both circuit shapes of the generator (regex.cpp): style 0,
the zk-regex shape, and style 1, the compact shape (same function of the input, ~4x fewer constraints)."""
import random
import re

import pytest

import zkemail_b200 as z
from zkutil import oracle_witness

PATTERN = re.compile(rb"(\r\n|^)dkim-signature:([a-z]+=[^;]+; )+bh=([a-zA-Z0-9+/=]+);")
N = 128


@pytest.fixture(scope="module", params=[0, 1], ids=["zkregex_shape", "compact_shape"])
def circuit(request):
    return z.Circuit("BodyHashRegex", [N, request.param])


def _run(circuit, msg: bytes, n=N):
    assert len(msg) <= n
    padded = list(msg) + [0] * (n - len(msg))
    w = oracle_witness(circuit, {"msg": padded})
    return w.values("out")[0], w.values("reveal0"), padded


def _expected_reveal(msg: bytes, n=N):
    m = PATTERN.search(msg)
    out = [0] * n
    if m:
        for i in range(m.start(3), m.end(3)):
            out[i] = msg[i]
    return out


BH = b"7xQMDuoVVU4m0W0WRVSrVXMeGSIASsnucK9dJsrc+vU="        # the bh= value of the reference's test.eml
CASES = {
    # name: (message, matches)
    "after_crlf": (b"to:a@b.c\r\ndkim-signature:v=1; a=rsa-sha256; bh=" + BH + b"; h=from:to; b=", True),
    "at_start_of_input": (b"dkim-signature:v=1; d=example.com; bh=" + BH[:20] + b"; b=", True),
    "no_line_start_before_name": (b"x-dkim-signature:v=1; a=rsa-sha256; bh=" + BH + b"; b=", False),
    "other_header_only": (b"subject:bh=abcd; hello\r\nfrom:a@b.c\r\n", False),
    "tag_list_without_bh": (b"from:x\r\ndkim-signature:v=1; a=rsa-sha256; d=example.com; h=from; b=abcd", False),
    "bh_is_the_first_tag": (b"\r\ndkim-signature:bh=" + BH + b"; v=1", False),          # ([a-z]+=[^;]+; )+ needs one tag before bh=
    "bh_value_not_terminated": (b"\r\ndkim-signature:v=1; bh=" + BH, False),
    "empty_bh_value": (b"\r\ndkim-signature:v=1; bh=; b=x", False),
    "utf8_inside_a_tag_value": ("\r\ndkim-signature:v=1; d=exämple€.org; bh=".encode() + BH[:12] + b"; b=", True),
    "upper_case_name_does_not_match": (b"\r\nDKIM-Signature:v=1; a=rsa-sha256; bh=" + BH + b"; b=", False),
    "missing_space_after_semicolon": (b"\r\ndkim-signature:v=1;bh=" + BH + b"; b=", False),
    "second_signature_header_matches": (b"dkim-signature:v=1; b=q\r\ndkim-signature:v=1; a=x; bh=QUJD; b=", True),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_body_hash_regex_cases(circuit, name):
    msg, matches = CASES[name]
    assert (PATTERN.search(msg) is not None) == matches, "the case itself is mislabelled"
    out, reveal, _ = _run(circuit, msg)
    assert out == (1 if matches else 0)
    if matches:
        assert reveal == _expected_reveal(msg), "reveal0 must be the bh= value at its own positions and 0 elsewhere"
        start = msg.index(b"bh=") + 3
        got = bytes(b for b in reveal if b)
        assert got == PATTERN.search(msg).group(3) and reveal[start] == msg[start] and reveal[start - 1] == 0


def test_reveal_feeds_select_regex_reveal(circuit):
    """email-verifier.circom:126-131: bhReveal + bodyHashIndex -> SelectRegexReveal(H, 44) -> the 44 base64 characters."""
    msg, _ = CASES["after_crlf"]
    out, reveal, _ = _run(circuit, msg)
    idx = msg.index(BH)
    sel = z.Circuit("SelectRegexReveal", [N, 44])
    w = oracle_witness(sel, {"in": reveal, "startIndex": idx})
    assert bytes(w.values("out")) == BH


@pytest.mark.parametrize("style", [0, 1])
def test_body_hash_regex_out_agrees_with_python_re_on_random_strings(style):
    n = 48
    c = z.Circuit("BodyHashRegex", [n, style])
    rnd = random.Random(20260923)
    pieces = [b"dkim-signature:", b"\r\n", b"bh=", b"; ", b";", b"v=1", b"a=b", b"d=x.y", b"QUJD", b"+/=", b"=", b" ", b"Z", b"k", b":"]
    agree = hits = 0
    for _ in range(300):
        if rnd.random() < 0.5:
            # a well-formed header, then (half of the time) one random edit: near-matches on both sides of the boundary
            msg = rnd.choice([b"", b"\r\n", b"k:Z\r\n"]) + b"dkim-signature:" + b"".join(rnd.choice([b"v=1; ", b"a=b; ", b"d=x.y; "]) for _ in range(rnd.randint(1, 2)))
            msg += b"bh=" + rnd.choice([b"QUJD", b"+/=", b"Z", b"k9"]) + b";"
            if rnd.random() < 0.5:
                pos = rnd.randrange(len(msg))
                edit = rnd.choice(pieces + [b""])
                msg = msg[:pos] + edit + msg[pos + 1:]
            msg = msg[:n]
        else:
            msg = b""
            while True:
                nxt = rnd.choice(pieces)
                if len(msg) + len(nxt) > n:
                    break
                msg += nxt
                if rnd.random() < 0.08:
                    break
        out, reveal, padded = _run(c, msg, n)
        m = PATTERN.search(bytes(padded))
        assert out == (1 if m else 0), msg
        if m:
            assert reveal == _expected_reveal(bytes(padded), n), msg
        hits += m is not None
        agree += 1
    assert agree == 300 and 10 < hits < 290, "the random strings must exercise both outcomes"


def test_generic_regex_entry_point_matches_named_template():
    """zke_circuit_build_regex (the generic zk-regex generator) given the body-hash parts builds the
    same constraint system as the named template, and another regex works through the same door."""
    parts = [("(\r\n|^)dkim-signature:", False), ("([a-z]+=[^;]+; )+bh=", False), ("[a-zA-Z0-9+/=]+", True), (";", False)]
    g = z.Circuit.from_regex(parts, 64)
    named = z.Circuit("BodyHashRegex", [64])
    assert (g.info.n_constraints, g.info.n_vars) == (named.info.n_constraints, named.info.n_vars)
    msg = b"\r\ndkim-signature:v=1; bh=QUJD; b="
    padded = list(msg) + [0] * (64 - len(msg))
    assert oracle_witness(g, {"msg": padded}).raw() == oracle_witness(named, {"msg": padded}).raw()
    t = z.Circuit.from_regex([("to:", False), ("[a-z0-9.@]+", True), ("\r\n", False)], 32)
    msg = b"from:x\r\nto:bob@mail.io\r\n"
    w = oracle_witness(t, {"msg": list(msg) + [0] * (32 - len(msg))})
    assert w.values("out") == [1] and bytes(b for b in w.values("reveal0") if b) == b"bob@mail.io"
    with pytest.raises(z._lib.ZkeError):
        z.Circuit.from_regex([("a*", True)], 8)          # matches the empty string


def test_generic_entry_point_in_the_compact_shape(monkeypatch):
    """ZKE_REGEX_STYLE=1 selects the compact shape for zke_circuit_build_regex; alternation, optional parts, nested repeats
    and a negated class behave as in the zk-regex shape and as Python's `re` says."""
    cases = [([("to:", False), ("[a-z0-9.@]+", True), ("\r\n", False)], rb"to:([a-z0-9.@]+)\r\n",
              [b"from:x\r\nto:bob@mail.io\r\n", b"to:\r\n", b"to:a\r\nto:bc\r\n", b"cc:bob\r\n"]),
             ([("(ab|cd)+x?", False), ("[^ ]+", True), (" ", False)], rb"(?:ab|cd)+x?([^ ]+) ",
              [b"abcdx12 ", b"zzabab9 z", b"ab ", b"cdxx y", b"ba x "]),
             ([("id=", False), ("(0|[1-9][0-9]*)", True), (";", False)], rb"id=(0|[1-9][0-9]*);",
              [b"id=0;", b"id=007;", b"xid=120;id=", b"id=;", b"id=12"])]
    for parts, pattern, msgs in cases:
        monkeypatch.setenv("ZKE_REGEX_STYLE", "0")
        c0 = z.Circuit.from_regex(parts, 24)
        monkeypatch.setenv("ZKE_REGEX_STYLE", "1")
        c1 = z.Circuit.from_regex(parts, 24)
        assert c1.info.n_constraints < c0.info.n_constraints
        for msg in msgs:
            padded = list(msg) + [0] * (24 - len(msg))
            w0, w1 = oracle_witness(c0, {"msg": padded}), oracle_witness(c1, {"msg": padded})
            m = re.search(pattern, bytes(padded))
            assert w0.values("out") == w1.values("out") == [1 if m else 0], (parts, msg)
            assert w0.values("reveal0") == w1.values("reveal0"), (parts, msg)


def test_compact_shape_is_smaller_and_reveals_the_same():
    """The two shapes are two circuits for one function: identical `out` / `reveal0` on inputs that include bytes >= 128,
    the 255 marker value inside the message, overlapping partial matches and several matches."""
    n = 96
    zk, compact = z.Circuit("BodyHashRegex", [n, 0]), z.Circuit("BodyHashRegex", [n, 1])
    assert compact.info.n_constraints * 3 < zk.info.n_constraints
    rnd = random.Random(7)
    base = b"\r\ndkim-signature:v=1; a=rsa-sha256; bh=QUJDREVG; b=x"
    msgs = [base, base[2:], b"\xff" + base, base[:20] + b"\xff" + base[20:], base + base, b"dkim-signature:" + base,
            "\r\ndkim-signature:v=1; d=\u00e9\u20ac\U0001f600; bh=QQ==; ".encode(), b"\r\ndkim-signature:v=1; d=\xc3\x28; bh=QQ==; ",
            base.replace(b"; bh", b";  bh"), b"bh=QUJD;" * 8]
    for _ in range(40):
        m = bytearray(base)
        for _k in range(rnd.randint(1, 4)):
            m[rnd.randrange(len(m))] = rnd.choice([rnd.randrange(256), ord(";"), ord("="), ord(" "), 13, 10])
        msgs.append(bytes(m))
    for msg in msgs:
        padded = list(msg[:n]) + [0] * (n - len(msg[:n]))
        a, c2 = oracle_witness(zk, {"msg": padded}), oracle_witness(compact, {"msg": padded})
        assert a.values("out") == c2.values("out") and a.values("reveal0") == c2.values("reveal0"), msg
    tw0, tw1 = z.Circuit("TwitterResetRegex", [64, 0]), z.Circuit("TwitterResetRegex", [64, 1])
    for msg in [b"x email was meant for @zk_mail.", b"email was meant for @", b"email was meant for @a email was meant for @bc!"]:
        padded = list(msg) + [0] * (64 - len(msg))
        a, c2 = oracle_witness(tw0, {"msg": padded}), oracle_witness(tw1, {"msg": padded})
        assert a.values("out") == c2.values("out") and a.values("reveal0") == c2.values("reveal0"), msg


@pytest.mark.gpu
def test_body_hash_regex_gpu_matches_oracle(circuit):
    ctx = z.Context(circuit, None, device=0, max_batch=len(CASES))
    names = sorted(CASES)
    packed = b"".join(circuit.pack_inputs({"msg": list(CASES[k][0]) + [0] * (N - len(CASES[k][0]))}) for k in names)
    wt, status = ctx.witness(packed, len(names))
    m = circuit.info.n_vars
    assert status == [-1] * len(names)
    for i, k in enumerate(names):
        msg = CASES[k][0]
        ref = oracle_witness(circuit, {"msg": list(msg) + [0] * (N - len(msg))})
        assert wt[32 * m * i:32 * m * (i + 1)] == ref.raw(), k
        assert int.from_bytes(wt[32 * (m * i + 1):32 * (m * i + 2)], "little") == (1 if CASES[k][1] else 0)
    ctx.close()
