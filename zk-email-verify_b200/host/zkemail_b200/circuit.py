"""Circuit handle: the role `wasm_tester(circuitFile)` / the `.r1cs + .wasm` pair play in the reference
(/root/reference/packages/circuits/tests/email-verifier.test.ts:21-31)."""
from __future__ import annotations
import ctypes
from . import _lib as L

FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617


class Circuit:
    def __init__(self, template: str, params=(), _handle=None):
        if _handle is None:
            arr = (L.c_i64 * len(params))(*[int(p) for p in params])
            err = ctypes.create_string_buffer(L.ERRCAP)
            _handle = L.zke_circuit_build(template.encode(), arr, len(params), err, L.ERRCAP)
            if not _handle:
                raise L.ZkeError(err.value.decode())
        self._h = _handle
        self.template, self.params = template, tuple(params)
        info = L.CircuitInfo()
        L.zke_circuit_get_info(self._h, ctypes.byref(info))
        self.info = info
        self.groups = {}
        name = ctypes.create_string_buffer(256)
        first, count, kind = L.c_u32(), L.c_u32(), L.c_int()
        for i in range(info.n_groups):
            L.zke_circuit_group(self._h, i, name, 256, ctypes.byref(first), ctypes.byref(count), ctypes.byref(kind))
            self.groups[name.value.decode()] = (first.value, count.value, kind.value)

    @classmethod
    def from_regex(cls, parts, msg_len: int):
        """zk-regex circuit of a decomposed regex: parts = [(regex fragment, is_public), ...]; signals msg[msg_len] ->
        out, reveal0[msg_len] (the generator behind BodyHashRegex, email-verifier.circom:5,126)."""
        arr = (L.c_char_p * len(parts))(*[p.encode() for p, _ in parts])
        pub = bytes(1 if q else 0 for _, q in parts)
        err = ctypes.create_string_buffer(L.ERRCAP)
        h = L.zke_circuit_build_regex(arr, pub, len(parts), msg_len, err, L.ERRCAP)
        if not h:
            raise L.ZkeError(err.value.decode())
        return cls("Regex", (msg_len,), _handle=h)

    def __del__(self):
        if getattr(self, "_h", None):
            L.zke_circuit_free(self._h)
            self._h = None

    @property
    def handle(self):
        return self._h

    @property
    def n_inputs(self):
        return self.info.n_pub_inputs + self.info.n_prv_inputs

    def pack_inputs(self, inputs: dict) -> bytes:
        """snarkjs input JSON ({name: decimal string | number | list}) -> [n_inputs][32] little-endian bytes,
        in witness order.  Mirrors the checks of circom_runtime's witness calculator: every declared input
        must be present with the right number of values ("Not all inputs have been set" / "Too many values")."""
        base = 1 + self.info.n_outputs
        buf = bytearray(32 * self.n_inputs)
        seen = set()
        for name, val in inputs.items():
            if name not in self.groups or self.groups[name][2] == 0:
                raise L.ZkeError(f"Signal not found: {name}")
            first, count, _ = self.groups[name]
            flat = _flatten(val)
            if len(flat) != count:
                raise L.ZkeError(f"{'Too many' if len(flat) > count else 'Not enough'} values for input signal {name}")
            for i, v in enumerate(flat):
                x = int(v) % FR_MODULUS
                off = 32 * (first - base + i)
                buf[off:off + 32] = x.to_bytes(32, "little")
            seen.add(name)
        for name, (_, _, kind) in self.groups.items():
            if kind != 0 and name not in seen:
                raise L.ZkeError(f"Not all inputs have been set. Missing: {name}")
        return bytes(buf)

    def array(self, which: int, ctype):
        n = L.c_size_t()
        p = L.zke_circuit_array(self._h, which, ctypes.byref(n))
        return p, n.value

    def scope_name(self, idx: int) -> str:
        s = L.zke_circuit_scope_name(self._h, idx)
        return s.decode() if s else "?"


def _flatten(v):
    if isinstance(v, (list, tuple)):
        out = []
        for x in v:
            out.extend(_flatten(x))
        return out
    return [v]
