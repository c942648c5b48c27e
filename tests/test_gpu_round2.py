"""GPU parity tests added in round 2 (run with `-m gpu`), all through the C ABI:

  * EmailVerifier with removeSoftLineBreaks = 1 and the qp-encoded-selector variant on the GPU witness kernel
    (/root/reference/packages/circuits/tests/email-verifier-with-soft-line-breaks.test.ts,
     email-verifier-with-qp-encoded-sha-precompute-selector.test.ts:34-49);
  * EmailVerifier(640, 768) proof bit-exact against the CPU oracle prover at fixed (r, s);
  * the `.zkey` path: write -> zke_zkey_load (whole file and the fork's b..k chunks) -> prove gives the bit-identical
    proof; a context opened from the key alone proves a `.wtns` (snarkjs `groth16 prove zkey wtns`);
  * the pipelined submit / collect form of fullprove returns what the synchronous call returns;
  * externally supplied witnesses are validated;
  * config 5 (EmailVerifier(1024, 16384), 2^24 domain) on one email - marked slow;
  * the compact regex shape (regex.cpp) inside EmailVerifier: witness and proof bit-exact against the oracle.
"""
import ctypes
import json
import random

import pytest

import zkemail_b200 as z
from zkemail_b200 import iden3_binfile as B
from zkemail_b200 import _lib as L
from zkutil import oracle_witness, oracle_prove, product_sections
from oracle import bn254

pytestmark = pytest.mark.gpu


def _resolver(key):
    return lambda n, t: [z.synthetic.key_record(key)]


# ------------------------------------------------------------------------------------------------ a15 on the GPU
def test_email_verifier_soft_line_breaks_gpu():
    from zkemail_b200.synthetic import make_signed_email
    c = z.Circuit("EmailVerifier", [640, 768, 121, 17, 0, 0, 0, 1, 1])
    key = z.synthetic.generate_key()
    body = b"This is a quoted-printable body with a soft line =\r\nbreak in the middle and another one right he=\r\nre.\r\n"
    email = make_signed_email(9, key, body_len=len(body), body_override=body)
    dk = z.verify_dkim_signature(email, resolver=_resolver(key))
    inputs = z.generate_email_verifier_inputs_from_dkim_result(
        dk, {"maxHeadersLength": 640, "maxBodyLength": 768, "removeSoftLineBreaks": True})
    bad = dict(inputs)
    dec = list(bad["decodedEmailBodyIn"])
    dec[3] = str((int(dec[3]) + 1) % 128)
    bad["decodedEmailBodyIn"] = dec
    ctx = z.Context(c, None, device=0, max_batch=2)
    wt, status = ctx.witness(c.pack_inputs(inputs) + c.pack_inputs(bad), 2, raise_on_fail=False)
    m = c.info.n_vars
    assert status[0] == -1 and status[1] >= 0
    assert wt[:32 * m] == oracle_witness(c, inputs).raw(), "GPU witness differs from the CPU oracle (removeSoftLineBreaks = 1)"
    with pytest.raises(z.AssertFailed, match="Assert Failed"):
        ctx.witness(c.pack_inputs(bad), 1)
    ctx.close()


def test_email_verifier_qp_encoded_selector_gpu():
    from zkemail_b200.synthetic import make_signed_email
    c = z.Circuit("EmailVerifier", [640, 1408, 121, 17, 0, 0, 0, 1, 1])
    key = z.synthetic.generate_key()
    filler = b"".join(b"Lorem ipsum dolor sit amet, consectetur adipiscing elit %03d.\r\n" % i for i in range(14))
    tail = b"Sed id imperdiet ne=\r\nque. Vivamus vel turpis non elit placerat feugiat ac a =\r\nmassa, and the rest of the body follows here.\r\n"
    email = make_signed_email(11, key, body_len=len(filler + tail), body_override=filler + tail)
    dk = z.verify_dkim_signature(email, resolver=_resolver(key))
    inputs = z.generate_email_verifier_inputs_from_dkim_result(
        dk, {"maxHeadersLength": 640, "maxBodyLength": 1408, "ignoreBodyHashCheck": False, "removeSoftLineBreaks": True,
             "shaPrecomputeSelector": "imperdiet neque."})
    ctx = z.Context(c, None, device=0, max_batch=1)
    wt, status = ctx.witness(c.pack_inputs(inputs), 1)
    assert status == [-1] and wt == oracle_witness(c, inputs).raw()
    ctx.close()


# ------------------------------------------------------------------------------------------------ EmailVerifier-scale prover parity
@pytest.fixture(scope="module")
def ev():
    c = z.Circuit("EmailVerifier", [640, 768, 121, 17, 0, 0, 0, 0, 1])
    key = z.synthetic.generate_key()
    inputs = []
    for i in range(3):
        email = z.synthetic.make_signed_email(20 + i, key, body_len=512)
        dk = z.verify_dkim_signature(email, resolver=_resolver(key))
        inputs.append(z.generate_email_verifier_inputs_from_dkim_result(dk, {"maxHeadersLength": 640, "maxBodyLength": 768}))
    zk = z.Zkey(c, seed=77, device=0)
    return c, zk, inputs


R_S = (0x1234567890abcdef1234567890abcdef1234567890abcdef, 0xfedcba0987654321fedcba0987654321fedcba09876543)


def _rs(batch):
    return b"".join((R_S[0] + k).to_bytes(32, "little") + (R_S[1] + 7 * k).to_bytes(32, "little") for k in range(batch))


def test_email_verifier_proof_bit_exact_vs_oracle(ev):
    """The whole proving pipeline at EmailVerifier scale (N = 2^21: fixed-base H table, three-pass NTTs, every witness
    MSM path) against the independent CPU prover, bit for bit."""
    c, zk, inputs = ev
    ctx = z.Context(c, zk, device=0, max_batch=1)
    wt, status = ctx.witness(c.pack_inputs(inputs[0]), 1)
    assert status == [-1]
    proofs, publics, _ = ctx.prove(1, _rs(1))
    want = oracle_prove(c, product_sections(zk), wt, R_S[0], R_S[1], threads=16)
    assert proofs == want, "GPU proof differs from the CPU oracle's proof at fixed (r, s)"
    proof, pubs = z.proof_to_json(proofs, publics, c.info.n_public)
    assert z.verify(zk.vkey(), pubs, proof) and bn254.groth16_verify(zk.vkey(), pubs, proof)
    ctx.close()


def test_compact_regex_shape_on_gpu(ev):
    """EmailVerifier with the compact regex shape (template parameter 10 = 1): GPU witness == oracle witness, the proof
    at fixed (r, s) == the CPU oracle's proof, same public signals as the zk-regex-shaped circuit, a header whose bh= tag
    is broken is rejected; TwitterVerifier and BodyHashRegex alone in the same shape."""
    c0, _, inputs = ev
    c = z.Circuit("EmailVerifier", [640, 768, 121, 17, 0, 0, 0, 0, 1, 1])
    assert c.info.n_constraints < c0.info.n_constraints - 300_000
    zk = z.Zkey(c, seed=78, device=0)
    ctx = z.Context(c, zk, device=0, max_batch=2)
    hdr = list(inputs[1]["emailHeader"])
    hdr[int(inputs[1]["bodyHashIndex"]) - 2] = str(ord("-"))
    packed = c.pack_inputs(inputs[0]) + c.pack_inputs(dict(inputs[1], emailHeader=hdr))
    wt, status = ctx.witness(packed, 2, raise_on_fail=False)
    assert status[0] == -1 and status[1] >= 0
    m = c.info.n_vars
    ref = oracle_witness(c, inputs[0])
    assert wt[:32 * m] == ref.raw()
    proofs, publics, st = ctx.fullprove(c.pack_inputs(inputs[0]), 1, _rs(1))
    assert st == [-1]
    want = oracle_prove(c, product_sections(zk), wt[:32 * m], R_S[0], R_S[1], threads=16)
    assert proofs == want
    ref0 = oracle_witness(c0, inputs[0])
    n_pub = c.info.n_public
    assert publics == ref0.raw()[32:32 * (1 + n_pub)]
    proof, pubs = z.proof_to_json(proofs, publics, n_pub)
    assert z.verify(zk.vkey(), pubs, proof) and bn254.groth16_verify(zk.vkey(), pubs, proof)
    ctx.close()
    for name, params, msg in (("BodyHashRegex", [128, 1], b"\r\ndkim-signature:v=1; a=rsa-sha256; bh=QUJDREVG; b=x"),
                              ("TwitterResetRegex", [64, 1], b"x email was meant for @zk_mail.")):
        r = z.Circuit(name, params)
        rctx = z.Context(r, None, device=0, max_batch=1)
        padded = list(msg) + [0] * (params[0] - len(msg))
        got, stt = rctx.witness(r.pack_inputs({"msg": padded}), 1)
        assert stt == [-1] and got == oracle_witness(r, {"msg": padded}).raw()
        rctx.close()


# ------------------------------------------------------------------------------------------------ .zkey / .wtns drop-in
def test_zkey_load_gives_bit_identical_proofs(ev):
    c, zk, inputs = ev
    packed = c.pack_inputs(inputs[1])
    ctx = z.Context(c, zk, device=0, max_batch=1)
    wt, _ = ctx.witness(packed, 1)
    want, want_pub, _ = ctx.prove(1, _rs(1))
    ctx.close()
    blob = B.write_zkey(zk)
    assert blob[:4] == b"zkey" and zk.is_toy

    # 1. whole file, circuit + loaded key: witness on the GPU, proving from the loaded points
    zk2 = z.Zkey.load(blob, device=0, circuit=c)
    assert not zk2.is_toy and zk2.info == zk.info
    assert zk2.vkey() == zk.vkey() and "vk_alphabeta_12" in zk2.vkey()
    ctx2 = z.Context(c, zk2, device=0, max_batch=1)
    got, got_pub, _ = ctx2.fullprove(packed, 1, _rs(1))
    assert got == want and got_pub == want_pub
    ctx2.close()

    # 2. the key alone (no circuit): `snarkjs groth16 prove zkey wtns` - coefficient matrices from section 4
    ctx3 = z.Context(None, zk2, device=0, max_batch=1)
    p3, pub3 = ctx3.wtns_prove(B.write_wtns(wt), _rs(1))
    assert p3 == want and pub3 == want_pub
    ctx3.load_witness(wt, 1)
    p3b, _, _ = ctx3.prove(1, _rs(1))
    assert p3b == want
    with pytest.raises(L.ZkeError, match="no witness program"):
        ctx3.witness(packed, 1)
    ctx3.close()
    # re-export of the loaded key reproduces the file; section 4 is regenerated from the CSR, i.e. the same records
    # in row order (the extra public rows of A now precede the B rows)
    _, sec_a = B.read_container(blob, b"zkey")
    _, sec_b = B.read_container(zk2.write(), b"zkey")
    assert sorted(sec_a) == sorted(sec_b) == list(range(1, 11))
    for s_ in (1, 2, 3, 5, 6, 7, 8, 9):
        assert sec_a[s_] == sec_b[s_], f"section {s_}"
    recs = lambda b: sorted(b[4 + 44 * k:48 + 44 * k] for k in range((len(b) - 4) // 44))
    assert sec_a[4][:4] == sec_b[4][:4] and recs(sec_a[4]) == recs(sec_b[4])
    del zk2, sec_b

    # 3. the fork's chunked form: files b..k = sections 1..10 (chunked-zkey.ts:9)
    sec = sec_a
    chunks = [sec[i] for i in range(1, 11)]
    zk4 = z.Zkey.load_chunks(chunks, device=0)
    ctx4 = z.Context(None, zk4, device=0, max_batch=1)
    p4, pub4 = ctx4.wtns_prove(B.write_wtns(wt), _rs(1))
    assert p4 == want and pub4 == want_pub
    ctx4.close()
    del zk4

    # malformed keys are refused: a point moved off the curve, a truncated file, a wrong field
    bad = bytearray(blob)
    off = 12 + sum(12 + len(sec[i]) for i in range(1, 5)) + 12 + 64 * 5      # sixth point of section 5
    bad[off] ^= 1
    with pytest.raises(L.ZkeError, match="not on the curve"):
        z.Zkey.load(bytes(bad), device=0)
    with pytest.raises(L.ZkeError, match="truncated"):
        z.Zkey.load(blob[:len(blob) // 2], device=0)
    with pytest.raises(L.ZkeError, match="magic"):
        z.Zkey.load(b"r1cs" + blob[4:], device=0)


def test_native_zkey_writer_matches_python_layout():
    c = z.Circuit("FpMul", [2, 4])
    zk = z.Zkey(c, seed=9)
    native = zk.write()
    python = B._write_container(b"zkey", 1, sorted(B.zkey_sections(zk).items()))
    assert native == python
    zk2 = z.Zkey.load(native, device=0)
    for sec in (L.SEC_ALPHA1, L.SEC_BETA1, L.SEC_DELTA1, L.SEC_BETA2, L.SEC_GAMMA2, L.SEC_DELTA2, L.SEC_IC, L.SEC_A, L.SEC_B1,
                L.SEC_B2, L.SEC_C, L.SEC_H):
        assert zk2.section(sec) == zk.section(sec)


def test_load_witness_is_validated(ev):
    c, zk, inputs = ev
    ctx = z.Context(c, zk, device=0, max_batch=1)
    wt, _ = ctx.witness(c.pack_inputs(inputs[2]), 1)
    bad = bytearray(wt)
    bad[32 * 5:32 * 6] = (z.FR_MODULUS + 3).to_bytes(32, "little")            # not reduced mod r
    with pytest.raises(L.ZkeError, match="not reduced"):
        ctx.load_witness(bytes(bad), 1)
    bad = bytearray(wt)
    bad[0] = 2                                                                   # w[0] != 1
    with pytest.raises(L.ZkeError, match=r"witness\[0\]"):
        ctx.load_witness(bytes(bad), 1)
    with pytest.raises(L.ZkeError, match="no witness loaded"):
        ctx.prove(1, _rs(1))
    rs_bad = (z.FR_MODULUS).to_bytes(32, "little") + (1).to_bytes(32, "little")
    ctx.load_witness(wt, 1)
    with pytest.raises(L.ZkeError, match="not reduced"):
        ctx.prove(1, rs_bad)
    # unreduced inputs are reduced the way snarkjs' witness calculator reduces them
    packed = bytearray(c.pack_inputs(inputs[2]))
    v = int.from_bytes(packed[:32], "little")
    packed[:32] = (v + z.FR_MODULUS).to_bytes(32, "little")
    wt2, status = ctx.witness(bytes(packed), 1)
    assert status == [-1] and wt2 == wt
    ctx.close()


def test_submit_collect_matches_synchronous_fullprove(ev):
    c, zk, inputs = ev
    batch = 3
    packed = b"".join(c.pack_inputs(i) for i in inputs)
    ctx = z.Context(c, zk, device=0, max_batch=batch)
    want = ctx.fullprove(packed, batch, _rs(batch))
    # three batches through the pipeline, two in flight at a time; the middle one carries a tampered email
    bad = dict(inputs[1])
    body = list(bad["emailBody"]); body[9] = str((int(body[9]) + 1) % 128); bad["emailBody"] = body
    packed_bad = c.pack_inputs(inputs[0]) + c.pack_inputs(bad) + c.pack_inputs(inputs[2])
    ctx.submit(packed, batch, _rs(batch))
    ctx.submit(packed_bad, batch, _rs(batch))
    with pytest.raises(L.ZkeError, match="in flight"):
        ctx.submit(packed, batch, _rs(batch))
    with pytest.raises(L.ZkeError, match="in flight"):
        ctx.fullprove(packed, batch, _rs(batch))
    first = ctx.collect()
    ctx.submit(packed, batch, _rs(batch))
    second = ctx.collect(raise_on_fail=False)
    third = ctx.collect()
    assert first == want and third == want
    assert second[2][0] == -1 and second[2][1] >= 0 and second[2][2] == -1
    assert second[0][:256] == want[0][:256] and second[0][512:] == want[0][512:] and second[0][256:512] == bytes(256)
    err = ctypes.create_string_buffer(256)
    assert L.zke_fullprove_collect(ctx.handle, ctypes.create_string_buffer(256), None, None, err, 256) < 0
    assert b"nothing was submitted" in err.value
    assert ctx.fullprove(packed, batch, _rs(batch)) == want      # the synchronous path still works afterwards
    ctx.close()


def test_context_on_every_visible_device():
    """The witness kernel's > 48 KB shared-memory opt-in is per device: a second context on another device must work."""
    n = z.device_count()
    c = z.Circuit("Sha256Bytes", [64])
    padded, plen = z.sha256_pad(b"dev", 64)
    inp = {"paddedIn": list(padded), "paddedInLength": plen}
    ref = oracle_witness(c, inp).raw()
    for d in range(min(n, 2)):
        ctx = z.Context(c, None, device=d, max_batch=1)
        wt, status = ctx.witness(c.pack_inputs(inp), 1)
        assert status == [-1] and wt == ref
        ctx.close()


# ------------------------------------------------------------------------------------------------ config 5 (slow)
@pytest.mark.slow
def test_config5_one_email_2pow24():
    """BASELINE configs[4] circuit: EmailVerifier(1024, 16384, 121, 17) = 10.2 M constraints, domain 2^24, one email with
    a 12 KB body: GPU witness == CPU oracle over all signals, the proof verifies, a flipped body byte is rejected."""
    c = z.Circuit("EmailVerifier", [1024, 16384, 121, 17])
    assert c.info.domain_log2 == 24
    key = z.synthetic.generate_key()
    email = z.synthetic.make_signed_email(5, key, body_len=12288)
    dk = z.verify_dkim_signature(email, resolver=_resolver(key))
    inputs = z.generate_email_verifier_inputs_from_dkim_result(dk, {"maxHeadersLength": 1024, "maxBodyLength": 16384})
    zk = z.Zkey(c, seed=5, device=0)
    ctx = z.Context(c, zk, device=0, max_batch=1)
    packed = c.pack_inputs(inputs)
    wt, status = ctx.witness(packed, 1)
    assert status == [-1]
    assert wt == oracle_witness(c, inputs).raw()
    proofs, publics, _ = ctx.prove(1)
    proof, pubs = z.proof_to_json(proofs, publics, c.info.n_public)
    assert z.verify(zk.vkey(), pubs, proof)
    bad = dict(inputs)
    body = list(bad["emailBody"]); body[4000] = str((int(body[4000]) + 1) % 128); bad["emailBody"] = body
    with pytest.raises(z.AssertFailed, match="Assert Failed"):
        ctx.fullprove(c.pack_inputs(bad), 1)
    ctx.close()


def test_registry_refuses_toy_keys_and_serves_loaded_ones(tmp_path):
    """generateProof / verifyProof mirror (chunked-zkey.ts:76-105): a key from the seeded toy setup is refused unless
    explicitly allowed; the same key written as the fork's chunk files and loaded back is served."""
    import hashlib
    c = z.Circuit("Sha256Bytes", [64])
    toy = z.Zkey(c, seed=11, device=0)
    with pytest.raises(z.InsecureKeyError):
        z.register_circuit("sha-toy", c, toy)
    chunks = B.write_zkey_chunks(toy)
    paths = []
    for suffix in "bcdefghijk":
        p = tmp_path / ("sha.zkey" + suffix)
        p.write_bytes(chunks["zkey" + suffix])
        paths.append(str(p))
    z.register_zkey_files("sha", c, paths)
    padded, plen = z.sha256_pad(b"registry", 64)
    out = z.generate_proof({"paddedIn": list(padded), "paddedInLength": plen}, "https://example.invalid/", "sha")
    digest = hashlib.sha256(b"registry").digest()
    assert out["publicSignals"] == [str((b >> (7 - j)) & 1) for b in digest for j in range(8)]
    assert z.verify_proof(out["proof"], out["publicSignals"], "https://example.invalid/", "sha")
    with pytest.raises(KeyError, match="after 3 retries"):
        z.generate_proof({}, "https://example.invalid/", "unknown-circuit")
