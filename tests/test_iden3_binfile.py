"""iden3 `.r1cs` / `.wtns` / `.zkey` containers (zkemail_b200.iden3_binfile): layout checks and
write -> read round trips against the engine's own arrays.  The reference only names these artefacts
(/root/reference/packages/helpers/src/chunked-zkey.ts:9,35-37; docs/zk-email-docs/UsageGuide/README.md:139-195) - the
formats live in un-vendored iden3 packages, so interop with snarkjs itself is not exercised here."""
import struct
import pytest

import zkemail_b200 as z
from zkemail_b200 import iden3_binfile as B
from zkemail_b200 import _lib as L
from zkutil import oracle_witness

P = z.FR_MODULUS


def test_r1cs_and_wtns_round_trip_and_satisfaction():
    c = z.Circuit("FpMul", [2, 4])
    blob = B.write_r1cs(c)
    assert blob[:4] == b"r1cs" and struct.unpack_from("<II", blob, 4) == (1, 3)
    d = B.read_r1cs(blob)
    i = c.info
    assert (d["n8"], d["prime"]) == (32, P)
    assert (d["nWires"], d["nPubOut"], d["nPubIn"], d["nPrvIn"], d["nConstraints"]) == (i.n_vars, i.n_outputs, i.n_pub_inputs, i.n_prv_inputs, i.n_constraints)
    assert d["map"] == list(range(i.n_vars)) and d["nLabels"] == i.n_vars
    w = oracle_witness(c, {"a": [1, 0, 1, 0], "b": [0, 1, 1, 0], "p": [1, 1, 1, 1]})
    wb = B.write_wtns(w.raw())
    assert wb[:4] == b"wtns" and struct.unpack_from("<II", wb, 4) == (2, 2)
    raw = B.read_wtns(wb)
    assert raw == w.raw() and raw[:32] == (1).to_bytes(32, "little")      # w[0] = 1
    vals = [int.from_bytes(raw[32 * k:32 * k + 32], "little") for k in range(i.n_vars)]
    ev = lambda lc: sum(v * vals[k] for k, v in lc.items()) % P
    assert all(ev(a) * ev(b) % P == ev(cc) for a, b, cc in d["constraints"])
    vals[i.n_vars - 1] = (vals[i.n_vars - 1] + 1) % P                      # a tampered witness must violate some row
    assert not all(ev(a) * ev(b) % P == ev(cc) for a, b, cc in d["constraints"])


def test_container_errors():
    with pytest.raises(ValueError):
        B.read_container(b"zkey" + bytes(8), b"r1cs")
    with pytest.raises(ValueError):
        B.read_wtns(B.write_wtns(bytes(64))[:-8])
    with pytest.raises(ValueError):
        B.write_wtns(bytes(33))


@pytest.mark.gpu
def test_zkey_round_trip_against_engine_sections():
    c = z.Circuit("FpMul", [2, 4])
    zk = z.Zkey(c, seed=9)
    blob = B.write_zkey(zk)
    assert blob[:4] == b"zkey" and struct.unpack_from("<II", blob, 4) == (1, 10)
    d = B.read_zkey(blob)
    i = c.info
    assert (d["q"], d["r"], d["nVars"], d["nPublic"], d["domainSize"]) == (B.FQ_MODULUS, P, i.n_vars, i.n_public, 1 << i.domain_log2)
    for name, sec in (("vk_alpha_1", L.SEC_ALPHA1), ("vk_beta_1", L.SEC_BETA1), ("vk_beta_2", L.SEC_BETA2), ("vk_gamma_2", L.SEC_GAMMA2),
                      ("vk_delta_1", L.SEC_DELTA1), ("vk_delta_2", L.SEC_DELTA2), ("IC", L.SEC_IC), ("A", L.SEC_A), ("B1", L.SEC_B1),
                      ("B2", L.SEC_B2), ("H", L.SEC_H)):
        assert d[name] == zk.section(sec), name
    assert d["C"] == zk.section(L.SEC_C)[64 * (i.n_public + 1):]
    assert len(d["IC"]) == 64 * (i.n_public + 1) and len(d["H"]) == 64 * (1 << i.domain_log2)
    # coefficient records: every A / B entry of the R1CS plus the n_public + 1 extra A rows, values back in standard form
    r1 = B.read_r1cs(B.write_r1cs(c))
    want = set()
    for row, (a, b, _c) in enumerate(r1["constraints"]):
        want |= {(0, row, s, v) for s, v in a.items()} | {(1, row, s, v) for s, v in b.items()}
    want |= {(0, i.n_constraints + j, j, 1) for j in range(i.n_public + 1)}
    assert set(d["coeffs"]) == want and len(d["coeffs"]) == len(want)
    chunks = B.write_zkey_chunks(zk)
    assert sorted(chunks) == ["zkey" + s for s in "bcdefghijk"] and chunks["zkeyj"] == B.zkey_sections(zk)[9]
