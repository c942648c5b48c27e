"""Test-side bridge between the product's circuit arrays and the CPU oracle (oracle/_ref/libzkref.so).

Mirrors circom_tester's three verbs used by the reference's circuit tests
(/root/reference/packages/circuits/tests/email-verifier.test.ts:43-44,204):
    calculateWitness -> oracle_witness, checkConstraints -> oracle_check, assertOut -> assert_out.
"""
from __future__ import annotations
import ctypes
import os
import subprocess

from zkemail_b200 import Circuit, FR_MODULUS
from zkemail_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(ROOT, "oracle")
_ORACLE_SO = os.path.join(_ORACLE_DIR, "_ref", "libzkref.so")


def _load_oracle():
    srcs = [os.path.join(_ORACLE_DIR, f) for f in os.listdir(_ORACLE_DIR) if f.endswith((".c", ".h"))]
    if (not os.path.exists(_ORACLE_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_ORACLE_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _ORACLE_DIR], stdout=subprocess.DEVNULL)
    return ctypes.CDLL(_ORACLE_SO)


ref = _load_oracle()


class RefCircuit(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in ("n_vars", "n_temps", "n_outputs", "n_inputs", "n_constraints", "n_ops",
                                               "n_coefs", "pad_")] + \
               [(n, ctypes.c_void_p) for n in ("coefs", "a_ptr", "a_var", "a_coef", "b_ptr", "b_var", "b_coef",
                                               "c_ptr", "c_var", "c_coef", "ops", "lc_ptr", "lc_var", "lc_coef", "aux")]


ref.zkref_witness.restype = ctypes.c_int
ref.zkref_witness.argtypes = [ctypes.POINTER(RefCircuit), ctypes.c_char_p, ctypes.c_void_p]
ref.zkref_check_r1cs.restype = ctypes.c_int64
ref.zkref_check_r1cs.argtypes = [ctypes.POINTER(RefCircuit), ctypes.c_void_p]


def ref_view(c: Circuit) -> RefCircuit:
    i = c.info
    rc = RefCircuit()
    rc.n_vars, rc.n_temps, rc.n_outputs = i.n_vars, i.n_temps, i.n_outputs
    rc.n_inputs, rc.n_constraints, rc.n_ops, rc.n_coefs = c.n_inputs, i.n_constraints, i.n_ops, i.n_coefs
    for name, which in (("coefs", L.ARR_COEFS), ("a_ptr", L.ARR_A_PTR), ("a_var", L.ARR_A_VAR), ("a_coef", L.ARR_A_COEF),
                        ("b_ptr", L.ARR_B_PTR), ("b_var", L.ARR_B_VAR), ("b_coef", L.ARR_B_COEF),
                        ("c_ptr", L.ARR_C_PTR), ("c_var", L.ARR_C_VAR), ("c_coef", L.ARR_C_COEF),
                        ("ops", L.ARR_OPS), ("lc_ptr", L.ARR_LC_PTR), ("lc_var", L.ARR_LC_VAR),
                        ("lc_coef", L.ARR_LC_COEF), ("aux", L.ARR_AUX)):
        p, _ = c.array(which, None)
        setattr(rc, name, p)
    return rc


class AssertFailed(Exception):
    """The reference's tests match the message against "Assert Failed" (email-verifier.test.ts:78)."""


class Witness:
    def __init__(self, circuit: Circuit, buf):
        self.circuit, self.buf = circuit, buf

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return [self[i] for i in range(*idx.indices(self.circuit.info.n_vars))]
        return int.from_bytes(self.buf[32 * idx:32 * idx + 32], "little")

    def values(self, name: str):
        first, count, _ = self.circuit.groups[name]
        return [self[first + i] for i in range(count)]

    def raw(self) -> bytes:
        return bytes(self.buf[: 32 * self.circuit.info.n_vars])


def oracle_witness(c: Circuit, inputs: dict, check: bool = True) -> Witness:
    """calculateWitness (+ the `===` asserts circom evaluates during witness generation)."""
    packed = c.pack_inputs(inputs)
    total = c.info.n_vars + c.info.n_temps
    buf = ctypes.create_string_buffer(32 * total)
    rc = ref_view(c)
    rcode = ref.zkref_witness(ctypes.byref(rc), packed, buf)
    if rcode != 0:
        raise RuntimeError(f"zkref_witness failed: {rcode}")
    w = Witness(c, buf)
    if check:
        oracle_check(c, w)
    return w


def oracle_check(c: Circuit, w: Witness):
    rc = ref_view(c)
    bad = ref.zkref_check_r1cs(ctypes.byref(rc), w.buf)
    if bad >= 0:
        p, n = c.array(L.ARR_SCOPE_OF_CONSTRAINT, None)
        scopes = (ctypes.c_uint16 * n).from_address(p)
        raise AssertFailed(f"Assert Failed: constraint {bad} in {c.scope_name(scopes[bad])}")


def assert_out(w: Witness, expected: dict):
    for name, val in expected.items():
        got = w.values(name)
        exp = [int(x) % FR_MODULUS for x in (val if isinstance(val, (list, tuple)) else [val])]
        assert got == exp, f"{name}: {got} != {exp}"


# ---------------------------------------------------------------------------------------------------------
# Groth16 oracle bridge (oracle/zkref_groth16.c)
# ---------------------------------------------------------------------------------------------------------
class RefZkey(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in ("n_vars", "n_public", "log_n", "pad_")] + \
               [(n, ctypes.c_char_p) for n in ("alpha1", "beta1", "delta1", "beta2", "delta2", "A", "B1", "C", "H", "B2")]


ref.zkref_groth16_prove.restype = ctypes.c_int
ref.zkref_groth16_prove.argtypes = [ctypes.POINTER(RefCircuit), ctypes.POINTER(RefZkey), ctypes.c_char_p, ctypes.c_char_p,
                                    ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p]
ref.zkref_groth16_setup.restype = ctypes.c_int
ref.zkref_groth16_setup.argtypes = [ctypes.POINTER(RefCircuit), ctypes.c_uint, ctypes.c_uint32, ctypes.c_char_p] + [ctypes.c_char_p] * 12

SECTIONS = ("alpha1", "beta1", "delta1", "beta2", "gamma2", "delta2", "IC", "A", "B1", "B2", "C", "H")


def oracle_setup(c: Circuit, toxic) -> dict:
    """Toy setup by the oracle for a SMALL circuit; toxic = (tau, alpha, beta, gamma, delta) integers."""
    i = c.info
    n, m, l = 1 << i.domain_log2, i.n_vars, i.n_public
    sizes = {"A": 64 * m, "B1": 64 * m, "B2": 128 * m, "C": 64 * m, "H": 64 * n, "IC": 64 * (l + 1),
             "alpha1": 64, "beta1": 64, "delta1": 64, "beta2": 128, "gamma2": 128, "delta2": 128}
    bufs = {k: ctypes.create_string_buffer(v) for k, v in sizes.items()}
    tox = b"".join(int(t).to_bytes(32, "little") for t in toxic)
    rc = ref_view(c)
    r = ref.zkref_groth16_setup(ctypes.byref(rc), i.domain_log2, l, tox, bufs["A"], bufs["B1"], bufs["B2"], bufs["C"], bufs["H"],
                                bufs["IC"], bufs["alpha1"], bufs["beta1"], bufs["delta1"], bufs["beta2"], bufs["gamma2"], bufs["delta2"])
    assert r == 0
    return {k: v.raw for k, v in bufs.items()}


def product_sections(zk) -> dict:
    from zkemail_b200 import _lib as L2
    ids = {"alpha1": L2.SEC_ALPHA1, "beta1": L2.SEC_BETA1, "delta1": L2.SEC_DELTA1, "beta2": L2.SEC_BETA2,
           "gamma2": L2.SEC_GAMMA2, "delta2": L2.SEC_DELTA2, "IC": L2.SEC_IC, "A": L2.SEC_A, "B1": L2.SEC_B1,
           "B2": L2.SEC_B2, "C": L2.SEC_C, "H": L2.SEC_H}
    return {k: zk.section(v) for k, v in ids.items()}


def _g1(b):
    x, y = int.from_bytes(b[:32], "little"), int.from_bytes(b[32:64], "little")
    return ["0", "1", "0"] if x == 0 and y == 0 else [str(x), str(y), "1"]


def _g2(b):
    v = [int.from_bytes(b[32 * i:32 * i + 32], "little") for i in range(4)]
    if not any(v):
        return [["0", "0"], ["1", "0"], ["0", "0"]]
    return [[str(v[0]), str(v[1])], [str(v[2]), str(v[3])], ["1", "0"]]


def vkey_from_sections(sec: dict, n_public: int) -> dict:
    return {"protocol": "groth16", "curve": "bn128", "nPublic": n_public, "vk_alpha_1": _g1(sec["alpha1"]),
            "vk_beta_2": _g2(sec["beta2"]), "vk_gamma_2": _g2(sec["gamma2"]), "vk_delta_2": _g2(sec["delta2"]),
            "IC": [_g1(sec["IC"][64 * i:64 * i + 64]) for i in range(n_public + 1)]}


def proof_json_from_bytes(p: bytes) -> dict:
    return {"pi_a": _g1(p[:64]), "pi_b": _g2(p[64:192]), "pi_c": _g1(p[192:256]), "protocol": "groth16", "curve": "bn128"}


def oracle_prove(c: Circuit, sec: dict, witness_bytes: bytes, r: int, s: int, threads: int = 8) -> bytes:
    i = c.info
    k = RefZkey()
    k.n_vars, k.n_public, k.log_n = i.n_vars, i.n_public, i.domain_log2
    for name in ("alpha1", "beta1", "delta1", "beta2", "delta2", "A", "B1", "C", "H", "B2"):
        setattr(k, name, sec[name])
    out = ctypes.create_string_buffer(256)
    rc = ref_view(c)
    rr = ref.zkref_groth16_prove(ctypes.byref(rc), ctypes.byref(k), witness_bytes, int(r).to_bytes(32, "little"),
                                 int(s).to_bytes(32, "little"), threads, out)
    assert rr == 0
    return out.raw
