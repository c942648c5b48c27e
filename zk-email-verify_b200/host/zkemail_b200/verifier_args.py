"""Verifier-argument encoders: the step after the proving path.

  * ark-serialize `serialize_compressed` images of a Groth16 proof, its public inputs and the verification key - what
    the reference's Rust CLI prints / embeds (/root/reference/packages/rust-verifier/src/main.rs:81-104 "PROOF" /
    "PUBLIC_INPUTS", :40-66 "[COMPRESSED_VKEY]"; consumed by src/verifier_template.rs:17-31).  Layout of ark-ec /
    ark-serialize 0.4 (un-vendored, Cargo.toml of the rust-verifier): a short-Weierstrass point is its x coordinate,
    little-endian, with two flag bits in the top of the last byte - 0x80 "y is the larger of (y, -y)", 0x40 infinity;
    Fq2 is c0 then c1 (flags on c1, and y is compared c1 first); field elements are 32 bytes little-endian; a Vec is
    prefixed with its u64 length; Proof = A, B, C; VerifyingKey = alpha_g1, beta_g2, gamma_g2, delta_g2, gamma_abc_g1.
    JSON field mapping as in src/verifier_utils.rs:64-172 (pi_b[0] = (x.c0, x.c1)).
  * `snarkjs zkey export soliditycalldata` / `groth16 exportSolidityCallData` argument string, and the field packing
    helpers of /root/reference/packages/contracts/utils/CircomUtils.sol:41-129 (packFieldsArray / unpackFieldsArray /
    packBool / unpackBool: 31 bytes per field element, little-endian within the element).
Host-side byte shuffling; no GPU involved.
"""
from __future__ import annotations

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FLAG_Y_NEGATIVE, FLAG_INFINITY = 0x80, 0x40


# ------------------------------------------------------------------------------------------------ field helpers
def _f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def _f2_sqrt(a):
    """square root in Fq2 = Fq[u]/(u^2 + 1), q = 3 mod 4; None if `a` is not a square"""
    if a == (0, 0):
        return (0, 0)
    n = (a[0] * a[0] + a[1] * a[1]) % Q
    s = pow(n, (Q + 1) // 4, Q)
    if s * s % Q != n:
        return None
    inv2 = pow(2, -1, Q)
    for sg in (s, Q - s):
        t = (a[0] + sg) * inv2 % Q
        y0 = pow(t, (Q + 1) // 4, Q)
        if y0 * y0 % Q != t:
            continue
        if y0 == 0:
            y1 = pow((-a[0]) % Q, (Q + 1) // 4, Q)
            if _f2_mul((0, y1), (0, y1)) == a:
                return (0, y1)
            continue
        y1 = a[1] * pow(2 * y0, -1, Q) % Q
        if _f2_mul((y0, y1), (y0, y1)) == (a[0] % Q, a[1] % Q):
            return (y0, y1)
    return None


_B2 = _f2_mul((3, 0), (lambda d: (9 * pow(d, -1, Q) % Q, (-1) * pow(d, -1, Q) % Q))(82))   # 3 / (9 + u)


def _g1_of(v):
    return None if int(v[2]) == 0 else (int(v[0]) % Q, int(v[1]) % Q)


def _g2_of(v):
    if int(v[2][0]) == 0 and int(v[2][1]) == 0:
        return None
    return ((int(v[0][0]) % Q, int(v[0][1]) % Q), (int(v[1][0]) % Q, int(v[1][1]) % Q))


# ------------------------------------------------------------------------------------------------ ark-serialize (compressed)
def ark_g1_compressed(pt) -> bytes:
    if pt is None:
        out = bytearray(32)
        out[31] |= FLAG_INFINITY
        return bytes(out)
    x, y = pt
    out = bytearray(x.to_bytes(32, "little"))
    if y > (Q - y) % Q:                       # SWFlags::from_y_coordinate: y <= -y is "positive"
        out[31] |= FLAG_Y_NEGATIVE
    return bytes(out)


def ark_g2_compressed(pt) -> bytes:
    if pt is None:
        out = bytearray(64)
        out[63] |= FLAG_INFINITY
        return bytes(out)
    (x0, x1), (y0, y1) = pt
    out = bytearray(x0.to_bytes(32, "little") + x1.to_bytes(32, "little"))
    neg = ((Q - y0) % Q, (Q - y1) % Q)
    if (y1, y0) > (neg[1], neg[0]):           # Fq2 ordering: c1 first, then c0
        out[63] |= FLAG_Y_NEGATIVE
    return bytes(out)


def ark_g1_decompress(b: bytes):
    flags = b[31] & 0xC0
    if flags & FLAG_INFINITY:
        return None
    x = int.from_bytes(bytes(b[:31]) + bytes([b[31] & 0x3F]), "little")
    if x >= Q:
        raise ValueError("x coordinate not reduced")
    y2 = (x * x * x + 3) % Q
    y = pow(y2, (Q + 1) // 4, Q)
    if y * y % Q != y2:
        raise ValueError("x is not the abscissa of a curve point")
    if (y > Q - y) != bool(flags & FLAG_Y_NEGATIVE):
        y = Q - y
    return (x, y)


def ark_g2_decompress(b: bytes):
    flags = b[63] & 0xC0
    if flags & FLAG_INFINITY:
        return None
    x0 = int.from_bytes(b[:32], "little")
    x1 = int.from_bytes(bytes(b[32:63]) + bytes([b[63] & 0x3F]), "little")
    if x0 >= Q or x1 >= Q:
        raise ValueError("x coordinate not reduced")
    x = (x0, x1)
    x3 = _f2_mul(_f2_mul(x, x), x)
    y = _f2_sqrt(((x3[0] + _B2[0]) % Q, (x3[1] + _B2[1]) % Q))
    if y is None:
        raise ValueError("x is not the abscissa of a curve point")
    neg = ((Q - y[0]) % Q, (Q - y[1]) % Q)
    if ((y[1], y[0]) > (neg[1], neg[0])) != bool(flags & FLAG_Y_NEGATIVE):
        y = neg
    return (x, y)


def ark_proof_compressed(proof: dict) -> bytes:
    """Proof::<Bn254>::serialize_compressed: A (32) || B (64) || C (32)  (main.rs:98-100)."""
    return ark_g1_compressed(_g1_of(proof["pi_a"])) + ark_g2_compressed(_g2_of(proof["pi_b"])) + ark_g1_compressed(_g1_of(proof["pi_c"]))


def ark_public_inputs_compressed(public_signals) -> bytes:
    """[Fr; N]::serialize_compressed (main.rs:95-96): N x 32 bytes little-endian, no length prefix (fixed-size array)."""
    out = b""
    for s in public_signals:
        v = int(s)
        if not 0 <= v < R:
            raise ValueError("public input not reduced mod r")
        out += v.to_bytes(32, "little")
    return out


def ark_vkey_compressed(vkey: dict) -> bytes:
    """VerifyingKey::<Bn254>::serialize_compressed, the [COMPRESSED_VKEY] of verifier_template.rs:19."""
    out = ark_g1_compressed(_g1_of(vkey["vk_alpha_1"]))
    for name in ("vk_beta_2", "vk_gamma_2", "vk_delta_2"):
        out += ark_g2_compressed(_g2_of(vkey[name]))
    out += len(vkey["IC"]).to_bytes(8, "little")
    for p in vkey["IC"]:
        out += ark_g1_compressed(_g1_of(p))
    return out


def rust_verifier_arguments(proof: dict, public_signals) -> dict:
    """What `GenerateVerifierArguments` prints (main.rs:81-104): the two byte vectors as lists of ints."""
    return {"PROOF": list(ark_proof_compressed(proof)), "PUBLIC_INPUTS": list(ark_public_inputs_compressed(public_signals))}


# ------------------------------------------------------------------------------------------------ Solidity
def _hex32(v) -> str:
    return "0x" + format(int(v), "064x")


def solidity_calldata(proof: dict, public_signals) -> str:
    """snarkjs `exportSolidityCallData`: pA, pB (each Fq2 as [c1, c0]), pC, pubSignals as 0x-prefixed 32-byte words."""
    a, b, c = proof["pi_a"], proof["pi_b"], proof["pi_c"]
    q = lambda s: '"' + _hex32(s) + '"'
    pa = "[%s, %s]" % (q(a[0]), q(a[1]))
    pb = "[[%s, %s],[%s, %s]]" % (q(b[0][1]), q(b[0][0]), q(b[1][1]), q(b[1][0]))
    pc = "[%s, %s]" % (q(c[0]), q(c[1]))
    pub = "[" + ",".join(q(s) for s in public_signals) + "]"
    return ",".join((pa, pb, pc, pub))


def pack_fields_array(data: bytes, padded_size: int) -> list[int]:
    """CircomUtils.packFieldsArray (CircomUtils.sol:41-71): 31 bytes per field element, byte j of a chunk at bit 8 j."""
    if len(data) > padded_size:
        raise ValueError("InvalidDataLength")
    n = (padded_size + 30) // 31
    fields = []
    for i in range(n):
        v = 0
        for j in range(31):
            idx = 31 * i + j
            if idx >= padded_size:
                break
            v += (data[idx] if idx < len(data) else 0) << (8 * j)
        fields.append(v)
    return fields


def unpack_fields_array(fields, padded_size: int) -> bytes:
    """CircomUtils.unpackFieldsArray (CircomUtils.sol:92-121): inverse of the above with trailing zeros trimmed."""
    n = (padded_size + 30) // 31
    out = bytearray()
    for i in range(n):
        f = int(fields[i])
        for _ in range(31):
            if len(out) >= padded_size:
                break
            out.append(f & 0xFF)
            f >>= 8
    while out and out[-1] == 0:
        out.pop()
    return bytes(out)


def pack_bool(v: bool) -> list[int]:       # CircomUtils.sol:78-82
    return [1 if v else 0]


def unpack_bool(fields) -> bool:           # CircomUtils.sol:128-130
    return int(fields[0]) == 1
