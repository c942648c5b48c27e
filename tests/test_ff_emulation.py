"""Device field arithmetic (zk-email-verify_b200/csrc/ff.cuh) checked on the CPU: the header is compiled with g++ under
ZKE_FF_EMULATE, which swaps the PTX carry-chain primitives for C equivalents, and every product / square / reduction
routine - the interleaved (CIOS) product the kernels use and the separated Karatsuba / half-product forms - is compared
with Python integers for both BN254 fields.  Role in the reference: wasmcurves' Fr / Fq multiplication (un-vendored)."""
import ctypes, os, random, subprocess, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    out = os.path.join(tempfile.gettempdir(), "libzke_ff_emulation_%d.so" % os.getuid())
    src = os.path.join(ROOT, "tests", "ff_emulation.cpp")
    inc = os.path.join(ROOT, "zk-email-verify_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", inc, "-Wno-unknown-pragmas", src, "-o", out])
    return ctypes.CDLL(out)


lib = _build()
P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
def limbs(x, n=8): return [(x >> (32*i)) & 0xffffffff for i in range(n)]
def arr(vals, n=8):
    flat = []
    for v in vals: flat += limbs(v, n)
    return (ctypes.c_uint32 * len(flat))(*flat)
def unarr(a, n, cnt): return [sum(a[i*n+j] << (32*j) for j in range(n)) for i in range(cnt)]
def _check(mod, name):
    R = 1 << 256
    inv = (-pow(mod, -1, 1 << 32)) % (1 << 32)
    lib.ff_set_consts(arr([mod]), arr([R % mod]), arr([R*R % mod]), ctypes.c_uint32(inv))
    random.seed(1)
    edge = [0, 1, 2, mod-1, mod-2, (1<<128)-1, 1<<128, (1<<128)+1, mod >> 1, 0xffffffff, (1<<253), ((1<<256)-1) % mod]
    A = edge + [random.randrange(mod) for _ in range(3000)]
    B = edge[::-1] + [random.randrange(mod) for _ in range(3000)]
    # make some with equal halves etc.
    for _ in range(200):
        h = random.randrange(1<<128); A.append((h | (h << 128)) % mod); B.append(random.randrange(mod))
        A.append(random.randrange(mod)); l = random.randrange(1<<125); B.append(l | (l<<128))
    A += [mod - 1, mod - 1, mod - 2, mod - 1]; B += [mod - 1, mod - 2, mod - 1, (mod - 1) // 2]   # largest sums of products
    n = len(A)
    a, b = arr(A), arr(B)
    Rinv = pow(R, -1, mod)
    for which, nm in [(0,'mul_cios'),(1,'mul_sos(karatsuba)'),(2,'mul_sos_plain'),(3,'sqr'),(8,'sqr_cios')]:
        out = (ctypes.c_uint32 * (8*n))()
        lib.ff_op(which, a, b, out, n)
        got = unarr(out, 8, n)
        bad = 0
        for i in range(n):
            exp = (A[i]*B[i]*Rinv) % mod if which not in (3, 8) else (A[i]*A[i]*Rinv) % mod
            if got[i] != exp:
                bad += 1
                if bad < 3: print(name, nm, 'MISMATCH', i, hex(A[i]), hex(B[i]), hex(got[i]), hex(exp))
        assert not bad, "%s %s: %d/%d wrong" % (name, nm, bad, n)
    for which, nm, f in [(4, 'add', lambda x, y: (x + y) % mod), (5, 'sub', lambda x, y: (x - y) % mod),
                         (6, 'mul_add2', lambda x, y: (x * y + y * ((x + y) % mod)) * Rinv % mod),
                         (7, 'mul_sub2', lambda x, y: (x * y - y * ((x + y) % mod)) * Rinv % mod)]:
        out = (ctypes.c_uint32 * (8*n))()
        lib.ff_op(which, a, b, out, n)
        got = unarr(out, 8, n)
        bad = sum(1 for i in range(n) if got[i] != f(A[i], B[i]))
        assert not bad, "%s %s: %d/%d wrong" % (name, nm, bad, n)
    for which, nm in [(0,'mul_wide8'),(1,'karatsuba'),(2,'sqr_wide')]:
        out = (ctypes.c_uint32 * (16*n))()
        # wide products on arbitrary 256-bit inputs too
        A2 = A[:1000] + [random.randrange(1<<256) for _ in range(1000)] + [(1<<256)-1, (1<<256)-1, 0, (1<<256)-1]
        B2 = B[:1000] + [random.randrange(1<<256) for _ in range(1000)] + [(1<<256)-1, 0, (1<<256)-1, 1]
        n2 = len(A2)
        out = (ctypes.c_uint32 * (16*n2))()
        lib.ff_wide(which, arr(A2), arr(B2), out, n2)
        got = unarr(out, 16, n2)
        bad = sum(1 for i in range(n2) if got[i] != (A2[i]*B2[i] if which != 2 else A2[i]*A2[i]))
        assert not bad, "%s %s: %d/%d wrong" % (name, nm, bad, n2)


def _check_shoup(mod, name):
    """mul_shoup(a, w, floor(w 2^256 / p)) == a * w mod p for every a < 2^256 (not only reduced ones) and w < p."""
    R = 1 << 256
    inv = (-pow(mod, -1, 1 << 32)) % (1 << 32)
    lib.ff_set_consts(arr([mod]), arr([R % mod]), arr([R*R % mod]), ctypes.c_uint32(inv))
    rnd = random.Random(7)
    edge_a = [0, 1, 2, mod - 1, mod, mod + 1, 2 * mod, 3 * mod, R - 1, R - 2, (1 << 255), (1 << 224) - 1, (1 << 192), 0xffffffff, (1 << 32)]
    edge_w = [0, 1, 2, mod - 1, mod - 2, mod >> 1, (1 << 253), (1 << 128) - 1, 0xffffffff, 3]
    A, W = [], []
    for a in edge_a:
        for w in edge_w:
            A.append(a); W.append(w)
    for _ in range(6000):
        A.append(rnd.randrange(R)); W.append(rnd.randrange(mod))
    for _ in range(500):      # operands with long runs of ones / zeros: worst cases for the truncated quotient estimate
        a = rnd.choice([R - 1, R - 1 - rnd.getrandbits(rnd.randrange(1, 200)), rnd.getrandbits(rnd.randrange(1, 256))])
        w = rnd.choice([mod - 1 - rnd.getrandbits(rnd.randrange(1, 200)), rnd.getrandbits(rnd.randrange(1, 253))]) % mod
        A.append(a % R); W.append(w)
    n = len(A)
    WQ = [(w << 256) // mod for w in W]
    assert max(WQ) < R
    out = (ctypes.c_uint32 * (8 * n))()
    lib.ff_shoup(arr(A), arr(W), arr(WQ), out, n)
    got = unarr(out, 8, n)
    bad = [i for i in range(n) if got[i] != A[i] * W[i] % mod]
    assert not bad, "%s mul_shoup: %d/%d wrong, first a=%x w=%x got=%x" % (name, len(bad), n, A[bad[0]], W[bad[0]], got[bad[0]])


def test_fr_shoup_product():
    _check_shoup(P, "Fr")
    _check_shoup(Q, "Fq")


def test_fr_arithmetic():
    _check(P, "Fr")


def test_fq_arithmetic():
    _check(Q, "Fq")

