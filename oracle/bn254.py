"""BN254 (alt_bn128 / circom "bn128") big-int arithmetic, pairing and Groth16 verifier.

TEST INFRASTRUCTURE (oracle) - never imported by the product path.

Restates, in plain Python integers, the arithmetic that the reference obtains from un-vendored
dependencies: snarkjs 0.5.0 fork -> ffjavascript 0.2.56 -> wasmcurves 0.2.0
(/root/reference/packages/helpers/package.json:26, yarn.lock:4646-4652) for prove/verify, and
ark-bn254 0.4.0 / ark-groth16 0.4.0 (/root/reference/packages/rust-verifier/Cargo.toml:7-13) for
the Rust verifier.  The JSON <-> curve point mapping follows
/root/reference/packages/rust-verifier/src/verifier_utils.rs:65-172 (pi_b[0][0] -> x.c0, pi_b[0][1] -> x.c1).

Pinned by tests/test_oracle_groth16_fixture.py against the reference's proof_of_twitter fixture.
"""
from __future__ import annotations

P = 21888242871839275222246405745257275088696311157297823662689037894645226208583  # Fq
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # Fr

ATE_LOOP_COUNT = 29793968203157093288
LOG_ATE_LOOP_COUNT = 63


def inv_mod(a: int, m: int) -> int:
    return pow(a % m, -1, m)


# --------------------------------------------------------------------------------------------
# Fq2 = Fq[u]/(u^2+1), elements are (c0, c1) tuples
# --------------------------------------------------------------------------------------------
def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def f2_neg(a):
    return ((-a[0]) % P, (-a[1]) % P)


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_sqr(a):
    return f2_mul(a, a)


def f2_scalar(a, k):
    return ((a[0] * k) % P, (a[1] * k) % P)


def f2_inv(a):
    d = inv_mod(a[0] * a[0] + a[1] * a[1], P)
    return ((a[0] * d) % P, (-a[1] * d) % P)


F2_ZERO = (0, 0)
F2_ONE = (1, 0)

# --------------------------------------------------------------------------------------------
# Fq12 as Fq[w]/(w^12 - 18 w^6 + 82)  (w^6 = 9 + u); dense 12-coefficient lists
# --------------------------------------------------------------------------------------------
_F12_MOD = [82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0]


class F12:
    __slots__ = ("c",)

    def __init__(self, c):
        self.c = [x % P for x in c]

    @staticmethod
    def one():
        return F12([1] + [0] * 11)

    @staticmethod
    def zero():
        return F12([0] * 12)

    def __add__(self, o):
        return F12([a + b for a, b in zip(self.c, o.c)])

    def __sub__(self, o):
        return F12([a - b for a, b in zip(self.c, o.c)])

    def __neg__(self):
        return F12([-a for a in self.c])

    def __eq__(self, o):
        return self.c == o.c

    def scale(self, k: int):
        return F12([a * k for a in self.c])

    def __mul__(self, o):
        a, b = self.c, o.c
        t = [0] * 23
        for i in range(12):
            ai = a[i]
            if ai:
                for j in range(12):
                    t[i + j] += ai * b[j]
        # reduce: w^12 = 18 w^6 - 82
        for i in range(22, 11, -1):
            top = t[i]
            if top:
                t[i - 6] += 18 * top
                t[i - 12] -= 82 * top
        return F12(t[:12])

    def __pow__(self, e: int):
        res = F12.one()
        base = self
        while e:
            if e & 1:
                res = res * base
            base = base * base
            e >>= 1
        return res

    def inv(self):
        # extended Euclid on polynomials over Fq (same method py_ecc uses)
        lm, hm = [1] + [0] * 12, [0] * 13
        low, high = self.c + [0], [x % P for x in _F12_MOD] + [1]

        def deg(p):
            d = len(p) - 1
            while d and p[d] == 0:
                d -= 1
            return d

        def poly_rounded_div(a, b):
            dega, degb = deg(a), deg(b)
            temp = list(a)
            o = [0] * len(a)
            for i in range(dega - degb, -1, -1):
                o[i] = (o[i] + temp[degb + i] * inv_mod(b[degb], P)) % P
                for c in range(degb + 1):
                    temp[c + i] = (temp[c + i] - o[i] * b[c]) % P
            return o[: deg(o) + 1]

        while deg(low):
            r = poly_rounded_div(high, low)
            r += [0] * (13 - len(r))
            nm, new = list(hm), list(high)
            for i in range(13):
                for j in range(13 - i):
                    nm[i + j] = (nm[i + j] - lm[i] * r[j]) % P
                    new[i + j] = (new[i + j] - low[i] * r[j]) % P
            lm, low, hm, high = nm, new, lm, low
        k = inv_mod(low[0], P)
        return F12([x * k for x in lm[:12]])


# --------------------------------------------------------------------------------------------
# Curves.  G1: y^2 = x^3 + 3 over Fq.  G2: y^2 = x^3 + 3/(9+u) over Fq2.  Affine points, None = infinity.
# --------------------------------------------------------------------------------------------
G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)
B1 = 3
B2 = f2_mul((3, 0), f2_inv((9, 1)))


def g1_is_on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B1) % P == 0


def g2_is_on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return f2_sub(f2_sqr(y), f2_add(f2_mul(f2_sqr(x), x), B2)) == F2_ZERO


def g1_neg(pt):
    return None if pt is None else (pt[0], (-pt[1]) % P)


def g1_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        m = 3 * x1 * x1 * inv_mod(2 * y1, P) % P
    else:
        m = (y2 - y1) * inv_mod(x2 - x1, P) % P
    x3 = (m * m - x1 - x2) % P
    return (x3, (m * (x1 - x3) - y1) % P)


def g1_mul(pt, k: int):
    k %= R
    # Jacobian double-and-add for speed
    if pt is None or k == 0:
        return None
    X, Y, Z = pt[0], pt[1], 1
    rx, ry, rz = 0, 1, 0
    for bit in bin(k)[2:]:
        rx, ry, rz = _jac_double(rx, ry, rz)
        if bit == "1":
            rx, ry, rz = _jac_add(rx, ry, rz, X, Y, Z)
    return _jac_to_affine(rx, ry, rz)


def _jac_double(X, Y, Z):
    if Z == 0:
        return X, Y, Z
    A = X * X % P
    B = Y * Y % P
    C = B * B % P
    D = 2 * ((X + B) * (X + B) - A - C) % P
    E = 3 * A % P
    F = E * E % P
    X3 = (F - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y * Z % P
    return X3, Y3, Z3


def _jac_add(X1, Y1, Z1, X2, Y2, Z2):
    if Z1 == 0:
        return X2, Y2, Z2
    if Z2 == 0:
        return X1, Y1, Z1
    Z1Z1 = Z1 * Z1 % P
    Z2Z2 = Z2 * Z2 % P
    U1 = X1 * Z2Z2 % P
    U2 = X2 * Z1Z1 % P
    S1 = Y1 * Z2 * Z2Z2 % P
    S2 = Y2 * Z1 * Z1Z1 % P
    if U1 == U2:
        if S1 == S2:
            return _jac_double(X1, Y1, Z1)
        return 0, 1, 0
    H = (U2 - U1) % P
    Rr = (S2 - S1) % P
    HH = H * H % P
    HHH = H * HH % P
    V = U1 * HH % P
    X3 = (Rr * Rr - HHH - 2 * V) % P
    Y3 = (Rr * (V - X3) - S1 * HHH) % P
    Z3 = Z1 * Z2 * H % P
    return X3, Y3, Z3


def _jac_to_affine(X, Y, Z):
    if Z == 0:
        return None
    zi = inv_mod(Z, P)
    zi2 = zi * zi % P
    return (X * zi2 % P, Y * zi2 * zi % P)


def g2_neg(pt):
    return None if pt is None else (pt[0], f2_neg(pt[1]))


def g2_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if f2_add(y1, y2) == F2_ZERO:
            return None
        m = f2_mul(f2_scalar(f2_sqr(x1), 3), f2_inv(f2_scalar(y1, 2)))
    else:
        m = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_sqr(m), x1), x2)
    return (x3, f2_sub(f2_mul(m, f2_sub(x1, x3)), y1))


def g2_mul(pt, k: int):
    k %= R
    res = None
    addend = pt
    while k:
        if k & 1:
            res = g2_add(res, addend)
        addend = g2_add(addend, addend)
        k >>= 1
    return res


# --------------------------------------------------------------------------------------------
# Optimal ate pairing (Miller loop over the twist embedded in Fq12, as in the textbook construction)
# --------------------------------------------------------------------------------------------
def _cast_g1(pt):
    return (F12([pt[0]] + [0] * 11), F12([pt[1]] + [0] * 11))


_W = F12([0, 1] + [0] * 10)
_W2 = _W * _W
_W3 = _W2 * _W


def _twist(pt):
    (x0, x1), (y0, y1) = pt
    # field isomorphism from Fq[u]/(u^2+1) to Fq[w^6]: u = w^6 - 9
    nx = F12([x0 - 9 * x1] + [0] * 5 + [x1] + [0] * 5)
    ny = F12([y0 - 9 * y1] + [0] * 5 + [y1] + [0] * 5)
    return (nx * _W2, ny * _W3)


def _f12_div(a: F12, b: F12) -> F12:
    return a * b.inv()


def _pt12_double(pt):
    x, y = pt
    m = _f12_div((x * x).scale(3), y.scale(2))
    nx = m * m - x.scale(2)
    ny = m * (x - nx) - y
    return (nx, ny)


def _pt12_add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2 and y1 == y2:
        return _pt12_double(a)
    if x1 == x2:
        return None
    m = _f12_div(y2 - y1, x2 - x1)
    nx = m * m - x1 - x2
    ny = m * (x1 - nx) - y1
    return (nx, ny)


def _linefunc(p1, p2, t):
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if not (x1 == x2):
        m = _f12_div(y2 - y1, x2 - x1)
        return m * (xt - x1) - (yt - y1)
    if y1 == y2:
        m = _f12_div((x1 * x1).scale(3), y1.scale(2))
        return m * (xt - x1) - (yt - y1)
    return xt - x1


def miller_loop(q2, p1) -> F12:
    """Miller loop of the optimal ate pairing, WITHOUT the final exponentiation."""
    if q2 is None or p1 is None:
        return F12.one()
    Q = _twist(q2)
    Pp = _cast_g1(p1)
    Rr = Q
    f = F12.one()
    for i in range(LOG_ATE_LOOP_COUNT, -1, -1):
        f = f * f * _linefunc(Rr, Rr, Pp)
        Rr = _pt12_double(Rr)
        if ATE_LOOP_COUNT & (1 << i):
            f = f * _linefunc(Rr, Q, Pp)
            Rr = _pt12_add(Rr, Q)
    Q1 = (Q[0] ** P, Q[1] ** P)
    nQ2 = (Q1[0] ** P, -(Q1[1] ** P))
    f = f * _linefunc(Rr, Q1, Pp)
    Rr = _pt12_add(Rr, Q1)
    f = f * _linefunc(Rr, nQ2, Pp)
    return f


def final_exponentiate(f: F12) -> F12:
    return f ** ((P ** 12 - 1) // R)


def pairing(q2, p1) -> F12:
    return final_exponentiate(miller_loop(q2, p1))


# --------------------------------------------------------------------------------------------
# snarkjs JSON <-> points, Groth16 verify
# --------------------------------------------------------------------------------------------
def g1_from_json(v):
    x, y, z = int(v[0]), int(v[1]), int(v[2])
    if z == 0:
        return None
    assert z == 1
    return (x, y)


def g2_from_json(v):
    x = (int(v[0][0]), int(v[0][1]))
    y = (int(v[1][0]), int(v[1][1]))
    z = (int(v[2][0]), int(v[2][1]))
    if z == (0, 0):
        return None
    assert z == (1, 0)
    return (x, y)


def g1_to_json(pt):
    if pt is None:
        return ["0", "1", "0"]
    return [str(pt[0]), str(pt[1]), "1"]


def g2_to_json(pt):
    if pt is None:
        return [["0", "0"], ["1", "0"], ["0", "0"]]
    return [[str(pt[0][0]), str(pt[0][1])], [str(pt[1][0]), str(pt[1][1])], ["1", "0"]]


def groth16_verify(vkey: dict, public_signals, proof: dict) -> bool:
    """snarkjs.groth16.verify(vkey, publicSignals, proof) semantics
    (call site /root/reference/packages/helpers/src/chunked-zkey.ts:101; Rust twin
    /root/reference/packages/rust-verifier/src/verifier_utils.rs:20 ``GrothBn::verify``).

    e(A, B) == e(alpha, beta) * e(sum_i pub_i * IC_i, gamma) * e(C, delta)   with pub_0 = 1.
    """
    try:
        if vkey.get("protocol", "groth16") != "groth16" or proof.get("protocol", "groth16") != "groth16":
            return False
        ic = [g1_from_json(p) for p in vkey["IC"]]
        pubs = [int(s) for s in public_signals]
        if len(pubs) + 1 != len(ic):
            return False
        if any(not (0 <= s < R) for s in pubs):
            return False
        a = g1_from_json(proof["pi_a"])
        b = g2_from_json(proof["pi_b"])
        c = g1_from_json(proof["pi_c"])
        alpha = g1_from_json(vkey["vk_alpha_1"])
        beta = g2_from_json(vkey["vk_beta_2"])
        gamma = g2_from_json(vkey["vk_gamma_2"])
        delta = g2_from_json(vkey["vk_delta_2"])
    except (KeyError, ValueError, AssertionError, IndexError, TypeError):
        return False
    for pt in (a, c, alpha, *ic):
        if not g1_is_on_curve(pt):
            return False
    for pt in (b, beta, gamma, delta):
        if not g2_is_on_curve(pt):
            return False
    vk_x = ic[0]
    for s, pt in zip(pubs, ic[1:]):
        vk_x = g1_add(vk_x, g1_mul(pt, s))
    f = miller_loop(b, g1_neg(a))
    f = f * miller_loop(beta, alpha)
    f = f * miller_loop(gamma, vk_x)
    f = f * miller_loop(delta, c)
    return final_exponentiate(f) == F12.one()
