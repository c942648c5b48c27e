"""Batched-affine accumulation (zk-email-verify_b200/csrc/ba.cuh) checked on the CPU under ZKE_FF_EMULATE against Python
elliptic-curve arithmetic: random points, repeated points (doubling), opposite points (cancellation), points at infinity,
odd / tiny list lengths, 0-3 tree levels.  Role in the reference: the bucket accumulation inside wasmcurves'
multiExpAffine (un-vendored), which snarkjs' groth16 prover calls for the H multi-exponentiation."""
import ctypes, os, random, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bn254

Q = bn254.P
RM = 1 << 256


def _build():
    out = os.path.join(tempfile.gettempdir(), "libzke_ba_emulation_%d.so" % os.getuid())
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "zk-email-verify_b200", "csrc"),
                           "-Wno-unknown-pragmas", os.path.join(ROOT, "tests", "ba_emulation.cpp"), "-o", out])
    lib = ctypes.CDLL(out)
    limbs = lambda x: [(x >> (32 * i)) & 0xffffffff for i in range(8)]
    arr = lambda x: (ctypes.c_uint32 * 8)(*limbs(x))
    lib.ba_set_consts(arr(Q), arr(RM % Q), arr(RM * RM % Q), ctypes.c_uint32((-pow(Q, -1, 1 << 32)) % (1 << 32)))
    return lib


lib = _build()


def _mont(x):
    return (x * RM % Q).to_bytes(32, "little")


def _reduce(points, entries, levels):
    blob = b"".join((_mont(p[0]) + _mont(p[1])) if p else bytes(64) for p in points)
    ent = (ctypes.c_uint32 * max(1, len(entries)))(*entries)
    out = ctypes.create_string_buffer(128)
    lib.ba_reduce(blob, ent, len(entries), levels, out)
    rinv = pow(RM, -1, Q)
    x, y, zz, zzz = [int.from_bytes(out.raw[32 * i:32 * i + 32], "little") * rinv % Q for i in range(4)]
    if zz == 0:
        return None
    return (x * pow(zz, -1, Q) % Q, y * pow(zzz, -1, Q) % Q)


def _expected(points, entries):
    acc = None
    for e in entries:
        p = points[e & 0x7fffffff]
        if p is not None and e >> 31:
            p = bn254.g1_neg(p)
        acc = bn254.g1_add(acc, p)
    return acc


def test_batched_affine_matches_python_ec():
    rng = random.Random(7)
    g = (1, 2)
    pts = [bn254.g1_mul(g, rng.randrange(1, bn254.R)) for _ in range(40)] + [None]
    inf_idx = len(pts) - 1
    cases = []
    for cnt in [0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 31, 64, 105]:
        cases.append([rng.randrange(40) | (rng.randrange(2) << 31) for _ in range(cnt)])
    cases.append([3, 3, 3, 3, 3, 3, 3, 3])                       # doublings at every level
    cases.append([5, 5 | (1 << 31), 6, 6 | (1 << 31), 7])        # cancellations
    cases.append([inf_idx, 1, 2, inf_idx, inf_idx, inf_idx, 4])  # points at infinity
    cases.append([8, 9, 8 | (1 << 31), 9 | (1 << 31)])           # sums that cancel one level up
    cases.append([1, 2, 1, 2, 1, 2, 1, 2, 1, 2])                 # equal partial sums -> doubling one level up
    for entries in cases:
        exp = _expected(pts, entries)
        for levels in (0, 1, 2, 3):
            got = _reduce(pts, entries, levels)
            assert got == exp, (entries, levels, got, exp)
