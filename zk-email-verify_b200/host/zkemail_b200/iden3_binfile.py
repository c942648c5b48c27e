"""iden3 binary containers of the circom / snarkjs tool chain: `.r1cs`, `.wtns`, `.zkey` (Groth16).

Role in the reference: the artefacts `generateProof` hands to snarkjs - `${circuitName}.wasm` / `.zkey`, the latter
split by the fork into per-section files with suffixes b..k (/root/reference/packages/helpers/src/chunked-zkey.ts:9,
35-37, 59-74) - and the `.r1cs` / `.wtns` files of the documented CLI flow
(/root/reference/docs/zk-email-docs/UsageGuide/README.md:139-195).  The file formats themselves live in the
un-vendored @iden3/binfileutils 0.0.11 / r1csfile 0.0.41 / snarkjs 0.5.0; the layouts below restate them as recorded
in SURVEY.md section 8(b):

    container : magic[4], u32 version, u32 nSections, then per section {u32 type, u64 size, payload}
    r1cs  v1  : 1 header {u32 n8, prime[n8], u32 nWires, nPubOut, nPubIn, nPrvIn, u64 nLabels, u32 nConstraints}
                2 constraints: 3 x {u32 nTerms, (u32 wire, coeff[n8]) x}     3 wire -> label map (u64 x nWires)
    wtns  v2  : 1 {u32 n8, q[n8], u32 nWitness}     2 nWitness x n8 bytes, standard form
    zkey  v1  : 1 {u32 protocol = 1}   2 header {n8q, q, n8r, r, nVars, nPublic, domainSize, alpha1, beta1, beta2,
                gamma2, delta1, delta2}   3 IC   4 coefficients {u32 n, (u32 matrix, u32 constraint, u32 signal,
                value[n8r]) x}   5 A   6 B1   7 B2   8 C (private signals only)   9 H   10 contributions
                points affine, coordinates in Montgomery form, little-endian, G2 as x.c0, x.c1, y.c0, y.c1;
                the point at infinity is all-zero bytes; coefficient values are stored multiplied by R^2.

Everything is little-endian.  Interoperability with snarkjs itself cannot be exercised offline (no node in this
image): what the tests pin is the layout above and write -> read round trips against the engine's own arrays.
"""
from __future__ import annotations
import ctypes
import io
import struct

from . import _lib as L
from .circuit import Circuit, FR_MODULUS

FQ_MODULUS = 21888242871839275222246405745257275088696311157297823662689037894645226208583
_R = 1 << 256
ZKEY_CHUNK_SUFFIXES = "bcdefghijk"   # chunked-zkey.ts:9 - one file per section 1..10


# ------------------------------------------------------------------------------------------------ container
def _write_container(magic: bytes, version: int, sections: list[tuple[int, bytes]]) -> bytes:
    out = io.BytesIO()
    out.write(magic)
    out.write(struct.pack("<II", version, len(sections)))
    for typ, payload in sections:
        out.write(struct.pack("<IQ", typ, len(payload)))
        out.write(payload)
    return out.getvalue()


def read_container(blob: bytes, magic: bytes) -> tuple[int, dict[int, bytes]]:
    if blob[:4] != magic:
        raise ValueError(f"bad magic {blob[:4]!r}, expected {magic!r}")
    version, n_sections = struct.unpack_from("<II", blob, 4)
    pos, sections = 12, {}
    for _ in range(n_sections):
        typ, size = struct.unpack_from("<IQ", blob, pos)
        pos += 12
        if pos + size > len(blob):
            raise ValueError("truncated section")
        sections[typ] = blob[pos:pos + size]
        pos += size
    return version, sections


def _le32(x: int) -> bytes:
    return int(x).to_bytes(32, "little")


# ------------------------------------------------------------------------------------------------ r1cs
def _u32_array(circuit: Circuit, which: int):
    p, n = circuit.array(which, None)
    return (ctypes.c_uint32 * n).from_address(p) if n else []


def _coefs(circuit: Circuit) -> list[bytes]:
    p, n = circuit.array(L.ARR_COEFS, None)
    raw = ctypes.string_at(p, 32 * n)
    return [raw[32 * i:32 * i + 32] for i in range(n)]


def write_r1cs(circuit: Circuit) -> bytes:
    """The circuit's constraint system as an iden3 `.r1cs` file (what `circom --r1cs` writes)."""
    i = circuit.info
    coefs = _coefs(circuit)
    header = struct.pack("<I", 32) + _le32(FR_MODULUS) + struct.pack("<IIIIQI", i.n_vars, i.n_outputs, i.n_pub_inputs,
                                                                      i.n_prv_inputs, i.n_vars, i.n_constraints)
    mats = [(_u32_array(circuit, a), _u32_array(circuit, b), _u32_array(circuit, c)) for a, b, c in
            ((L.ARR_A_PTR, L.ARR_A_VAR, L.ARR_A_COEF), (L.ARR_B_PTR, L.ARR_B_VAR, L.ARR_B_COEF), (L.ARR_C_PTR, L.ARR_C_VAR, L.ARR_C_COEF))]
    body = io.BytesIO()
    for row in range(i.n_constraints):
        for ptr, var, coef in mats:
            beg, end = ptr[row], ptr[row + 1]
            body.write(struct.pack("<I", end - beg))
            for k in range(beg, end):
                body.write(struct.pack("<I", var[k]))
                body.write(coefs[coef[k]])
    labels = b"".join(struct.pack("<Q", w) for w in range(i.n_vars))
    return _write_container(b"r1cs", 1, [(1, header), (2, body.getvalue()), (3, labels)])


def read_r1cs(blob: bytes) -> dict:
    version, sec = read_container(blob, b"r1cs")
    n8 = struct.unpack_from("<I", sec[1], 0)[0]
    prime = int.from_bytes(sec[1][4:4 + n8], "little")
    n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_constraints = struct.unpack_from("<IIIIQI", sec[1], 4 + n8)
    constraints, pos, body = [], 0, sec[2]
    for _ in range(n_constraints):
        row = []
        for _m in range(3):
            n_terms = struct.unpack_from("<I", body, pos)[0]
            pos += 4
            lc = {}
            for _t in range(n_terms):
                wire = struct.unpack_from("<I", body, pos)[0]
                lc[wire] = int.from_bytes(body[pos + 4:pos + 4 + n8], "little")
                pos += 4 + n8
            row.append(lc)
        constraints.append(tuple(row))
    return {"version": version, "n8": n8, "prime": prime, "nWires": n_wires, "nPubOut": n_pub_out, "nPubIn": n_pub_in,
            "nPrvIn": n_prv_in, "nLabels": n_labels, "nConstraints": n_constraints, "constraints": constraints,
            "map": list(struct.unpack("<%dQ" % n_wires, sec[3]))}


# ------------------------------------------------------------------------------------------------ wtns
def write_wtns(witness: bytes) -> bytes:
    """witness: nWitness x 32 bytes little-endian, standard form (what Context.witness returns for one email)."""
    if len(witness) % 32:
        raise ValueError("witness length is not a multiple of 32")
    header = struct.pack("<I", 32) + _le32(FR_MODULUS) + struct.pack("<I", len(witness) // 32)
    return _write_container(b"wtns", 2, [(1, header), (2, bytes(witness))])


def read_wtns(blob: bytes) -> bytes:
    _, sec = read_container(blob, b"wtns")
    n8 = struct.unpack_from("<I", sec[1], 0)[0]
    q = int.from_bytes(sec[1][4:4 + n8], "little")
    n = struct.unpack_from("<I", sec[1], 4 + n8)[0]
    if n8 != 32 or q != FR_MODULUS or len(sec[2]) != n * n8:
        raise ValueError("unsupported .wtns header")
    return sec[2]


# ------------------------------------------------------------------------------------------------ zkey
def _mont_coords(std: bytes, n_coords: int) -> bytes:
    """standard-form little-endian Fq coordinates -> Montgomery form; all-zero points (infinity) stay all-zero"""
    out = bytearray(len(std))
    for k in range(len(std) // 32):
        v = int.from_bytes(std[32 * k:32 * k + 32], "little")
        if v:
            out[32 * k:32 * k + 32] = (v * _R % FQ_MODULUS).to_bytes(32, "little")
    return bytes(out)


def _std_coords(mont: bytes) -> bytes:
    rinv = pow(_R, -1, FQ_MODULUS)
    out = bytearray(len(mont))
    for k in range(len(mont) // 32):
        v = int.from_bytes(mont[32 * k:32 * k + 32], "little")
        if v:
            out[32 * k:32 * k + 32] = (v * rinv % FQ_MODULUS).to_bytes(32, "little")
    return bytes(out)


def zkey_sections(zkey) -> dict[int, bytes]:
    """Sections 1..10 of the `.zkey` for a product proving key (engine.Zkey)."""
    circuit = zkey.circuit
    i = circuit.info
    n_public = i.n_public
    domain = 1 << i.domain_log2
    g1 = lambda sec: _mont_coords(zkey.section(sec), 2)
    g2 = lambda sec: _mont_coords(zkey.section(sec), 4)
    header = (struct.pack("<I", 32) + _le32(FQ_MODULUS) + struct.pack("<I", 32) + _le32(FR_MODULUS) +
              struct.pack("<III", i.n_vars, n_public, domain) +
              g1(L.SEC_ALPHA1) + g1(L.SEC_BETA1) + g2(L.SEC_BETA2) + g2(L.SEC_GAMMA2) + g1(L.SEC_DELTA1) + g2(L.SEC_DELTA2))
    # coefficients of A and B (C is not stored: snarkjs recomputes c = a o b), values times R^2, plus the n_public + 1
    # extra rows of A that make the public-input polynomials independent
    coefs = [int.from_bytes(c, "little") * _R * _R % FR_MODULUS for c in _coefs(circuit)]
    coef_bytes = [c.to_bytes(32, "little") for c in coefs]
    recs = io.BytesIO()
    n_recs = 0
    for m, (pw, vw, cw) in enumerate(((L.ARR_A_PTR, L.ARR_A_VAR, L.ARR_A_COEF), (L.ARR_B_PTR, L.ARR_B_VAR, L.ARR_B_COEF))):
        ptr, var, coef = _u32_array(circuit, pw), _u32_array(circuit, vw), _u32_array(circuit, cw)
        for row in range(i.n_constraints):
            for k in range(ptr[row], ptr[row + 1]):
                recs.write(struct.pack("<III", m, row, var[k]))
                recs.write(coef_bytes[coef[k]])
                n_recs += 1
    one_r2 = (_R * _R % FR_MODULUS).to_bytes(32, "little")
    for j in range(n_public + 1):
        recs.write(struct.pack("<III", 0, i.n_constraints + j, j))
        recs.write(one_r2)
        n_recs += 1
    c_all = g1(L.SEC_C)
    return {1: struct.pack("<I", 1), 2: header, 3: g1(L.SEC_IC), 4: struct.pack("<I", n_recs) + recs.getvalue(),
            5: g1(L.SEC_A), 6: g1(L.SEC_B1), 7: g2(L.SEC_B2), 8: c_all[64 * (n_public + 1):], 9: g1(L.SEC_H),
            10: bytes(64) + struct.pack("<I", 0)}   # circuit hash placeholder, no contributions (toy setup)


def write_zkey(zkey) -> bytes:
    """`.zkey` file image, written natively by the engine (zke_zkey_write); zkey_sections() above is the Python
    restatement of the same layout, kept for small circuits and as a cross-check of the native writer."""
    return zkey.write()


def write_zkey_chunks(zkey) -> dict[str, bytes]:
    """The fork's chunked layout: `${name}.zkey{b..k}` holds section 1..10 (chunked-zkey.ts:9)."""
    _, sec = read_container(zkey.write(), b"zkey")
    return {"zkey" + ZKEY_CHUNK_SUFFIXES[s - 1]: payload for s, payload in sec.items()}


def read_zkey(blob: bytes) -> dict:
    version, sec = read_container(blob, b"zkey")
    if struct.unpack("<I", sec[1])[0] != 1:
        raise ValueError("not a Groth16 zkey")
    h, pos = sec[2], 0
    n8q = struct.unpack_from("<I", h, pos)[0]; pos += 4
    q = int.from_bytes(h[pos:pos + n8q], "little"); pos += n8q
    n8r = struct.unpack_from("<I", h, pos)[0]; pos += 4
    r = int.from_bytes(h[pos:pos + n8r], "little"); pos += n8r
    n_vars, n_public, domain = struct.unpack_from("<III", h, pos); pos += 12
    names = (("vk_alpha_1", 64), ("vk_beta_1", 64), ("vk_beta_2", 128), ("vk_gamma_2", 128), ("vk_delta_1", 64), ("vk_delta_2", 128))
    out = {"version": version, "q": q, "r": r, "nVars": n_vars, "nPublic": n_public, "domainSize": domain}
    for name, size in names:
        out[name] = _std_coords(h[pos:pos + size]); pos += size
    n_recs = struct.unpack_from("<I", sec[4], 0)[0]
    rr_inv = pow(_R * _R, -1, FR_MODULUS)
    coeffs = []
    for k in range(n_recs):
        m, row, sig = struct.unpack_from("<III", sec[4], 4 + 44 * k)
        v = int.from_bytes(sec[4][4 + 44 * k + 12:4 + 44 * k + 44], "little") * rr_inv % FR_MODULUS
        coeffs.append((m, row, sig, v))
    out.update({"IC": _std_coords(sec[3]), "coeffs": coeffs, "A": _std_coords(sec[5]), "B1": _std_coords(sec[6]),
                "B2": _std_coords(sec[7]), "C": _std_coords(sec[8]), "H": _std_coords(sec[9])})
    return out
