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
