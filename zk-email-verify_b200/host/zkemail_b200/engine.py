"""Proving key and GPU context wrappers over the C ABI (zke_setup / zke_ctx_* / zke_witness / zke_prove)."""
from __future__ import annotations
import ctypes
import json
from . import _lib as L
from .circuit import Circuit


class AssertFailed(L.ZkeError):
    """Raised when a witness violates a constraint; the message contains "Assert Failed" - the string the
    reference's tests match (/root/reference/packages/circuits/tests/email-verifier.test.ts:78)."""


def device_count() -> int:
    return L.zke_device_count()


class Zkey:
    """Groth16 proving key resident on one GPU.  `Zkey(circuit, seed)` runs the TOY seeded setup (known toxic waste:
    tests and benchmarks only); `Zkey.load(bytes)` / `Zkey.load_chunks([...])` ingest a real snarkjs `.zkey`."""

    def __init__(self, circuit: Circuit | None, seed: int = 1, device: int = 0, _handle=None):
        if _handle is None:
            err = ctypes.create_string_buffer(L.ERRCAP)
            _handle = L.zke_setup(circuit.handle, seed, device, err, L.ERRCAP)
            if not _handle:
                raise L.ZkeError(err.value.decode())
        self._h = _handle
        self.circuit, self.device = circuit, device

    @classmethod
    def load(cls, zkey_bytes: bytes, device: int = 0, circuit: Circuit | None = None):
        """`${circuitName}.zkey` as handed to snarkjs.groth16.fullProve (chunked-zkey.ts:80-84)."""
        err = ctypes.create_string_buffer(L.ERRCAP)
        h = L.zke_zkey_load(zkey_bytes, len(zkey_bytes), device, err, L.ERRCAP)
        if not h:
            raise L.ZkeError(err.value.decode())
        return cls(circuit, device=device, _handle=h)

    @classmethod
    def load_chunks(cls, chunks, device: int = 0, circuit: Circuit | None = None):
        """The fork's chunked key: chunks[i] = contents of `${circuitName}.zkey{b..k}[i]` (chunked-zkey.ts:9,35-37)."""
        keep = [bytes(c) for c in chunks]
        ptrs = (L.c_void_p * len(keep))(*[ctypes.cast(ctypes.c_char_p(c), L.c_void_p) for c in keep])
        lens = (L.c_size_t * len(keep))(*[len(c) for c in keep])
        err = ctypes.create_string_buffer(L.ERRCAP)
        h = L.zke_zkey_load_chunks(ptrs, lens, len(keep), device, err, L.ERRCAP)
        if not h:
            raise L.ZkeError(err.value.decode())
        return cls(circuit, device=device, _handle=h)

    @property
    def is_toy(self) -> bool:
        return L.zke_zkey_is_toy(self._h) == 1

    @property
    def info(self):
        a, b, c = L.c_u32(), L.c_u32(), L.c_u32()
        L.zke_zkey_info(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return {"n_vars": a.value, "n_public": b.value, "domain_log2": c.value}

    def write(self) -> bytes:
        """The key as a `.zkey` file image (iden3 binfile, what `snarkjs groth16 setup` writes)."""
        ch = self.circuit.handle if self.circuit is not None else None
        n = L.zke_zkey_write(self._h, ch, None, 0)
        if n < 0:
            raise L.ZkeError("zkey export failed (a key made by the toy setup needs its circuit)")
        buf = ctypes.create_string_buffer(n)
        if L.zke_zkey_write(self._h, ch, buf, n) != n:
            raise L.ZkeError("zkey export failed")
        return buf.raw

    def __del__(self):
        if getattr(self, "_h", None):
            L.zke_zkey_free(self._h)
            self._h = None

    @property
    def handle(self):
        return self._h

    def vkey(self) -> dict:
        """`snarkjs zkey export verificationkey` -> vkey.json object."""
        n = L.c_size_t(0)
        L.zke_zkey_vkey_json(self._h, None, ctypes.byref(n))
        buf = ctypes.create_string_buffer(n.value)
        n2 = L.c_size_t(n.value)
        if L.zke_zkey_vkey_json(self._h, buf, ctypes.byref(n2)) != 0:
            raise L.ZkeError("vkey export failed")
        return json.loads(buf.value.decode())

    def section(self, sec: int) -> bytes:
        cnt = L.zke_zkey_section(self._h, sec, None, 0)
        if cnt < 0:
            raise L.ZkeError("bad section")
        size = 128 if sec in (L.SEC_BETA2, L.SEC_GAMMA2, L.SEC_DELTA2, L.SEC_B2) else 64
        buf = ctypes.create_string_buffer(size * cnt)
        if L.zke_zkey_section(self._h, sec, buf, len(buf)) < 0:
            raise L.ZkeError("section read failed")
        return buf.raw


class Context:
    def __init__(self, circuit: Circuit | None, zkey: Zkey | None = None, device: int = 0, max_batch: int = 1):
        err = ctypes.create_string_buffer(L.ERRCAP)
        self._h = L.zke_ctx_open(circuit.handle if circuit is not None else None, zkey.handle if zkey else None, device,
                                 max_batch, err, L.ERRCAP)
        if not self._h:
            raise L.ZkeError(err.value.decode())
        self.circuit, self.zkey, self.device, self.max_batch = circuit, zkey, device, max_batch
        self._pending = []
        self.n_vars = circuit.info.n_vars if circuit is not None else zkey.info["n_vars"]
        self.n_public = circuit.info.n_public if circuit is not None else zkey.info["n_public"]

    def close(self):
        if getattr(self, "_h", None):
            L.zke_ctx_close(self._h)
            self._h = None

    __del__ = close

    @property
    def handle(self):
        return self._h

    @property
    def stream(self) -> int:
        return L.zke_ctx_stream(self._h) or 0

    def upload_inputs(self, packed_inputs, batch: int):
        err = ctypes.create_string_buffer(L.ERRCAP)
        if L.zke_upload_inputs(self._h, packed_inputs, batch, err, L.ERRCAP) != 0:
            raise L.ZkeError(err.value.decode())

    def profile(self, enable: bool = True):
        L.zke_ctx_profile(self._h, 1 if enable else 0)

    def profile_get(self) -> dict:
        n = len(L.STAGES)
        ms, cnt = (ctypes.c_double * n)(), (L.c_u64 * n)()
        L.zke_ctx_profile_get(self._h, ms, cnt)
        return {name: {"ms": ms[i], "count": cnt[i]} for i, name in enumerate(L.STAGES)}

    def witness(self, packed_inputs, batch: int, want_witness: bool = True, raise_on_fail: bool = True):
        """calculateWitness + checkConstraints.  Returns (witness bytes | None, status list)."""
        m = self.n_vars
        out = ctypes.create_string_buffer(32 * m * batch) if want_witness else None
        status = (ctypes.c_int32 * batch)()
        err = ctypes.create_string_buffer(L.ERRCAP)
        rc = L.zke_witness(self._h, packed_inputs, batch, out, status, err, L.ERRCAP)
        if rc < 0:
            raise L.ZkeError(err.value.decode())
        if rc > 0 and raise_on_fail:
            raise AssertFailed(err.value.decode())
        return (out.raw if out is not None else None), list(status)

    def load_witness(self, wtns: bytes, batch: int):
        err = ctypes.create_string_buffer(L.ERRCAP)
        if L.zke_load_witness(self._h, wtns, batch, err, L.ERRCAP) != 0:
            raise L.ZkeError(err.value.decode())

    def _prove_call(self, fn, head_args, batch, rs, raise_on_fail):
        npub = self.n_public
        proofs = ctypes.create_string_buffer(256 * batch)
        publics = ctypes.create_string_buffer(max(1, 32 * npub * batch))
        status = (ctypes.c_int32 * batch)()
        err = ctypes.create_string_buffer(L.ERRCAP)
        rc = fn(self._h, *head_args, batch, rs, proofs, publics, status, err, L.ERRCAP)
        if rc < 0:
            raise L.ZkeError(err.value.decode())
        if rc > 0 and raise_on_fail:
            raise AssertFailed(err.value.decode())
        return proofs.raw, publics.raw[: 32 * npub * batch], list(status)

    def prove(self, batch: int, rs: bytes | None = None, raise_on_fail: bool = True):
        return self._prove_call(L.zke_prove, (), batch, rs, raise_on_fail)

    def wtns_prove(self, wtns_file: bytes, rs: bytes | None = None):
        """`snarkjs groth16 prove zkey wtns`: one `.wtns` file image -> (proof bytes, public signal bytes)."""
        npub = self.n_public
        proof = ctypes.create_string_buffer(256)
        publics = ctypes.create_string_buffer(max(1, 32 * npub))
        err = ctypes.create_string_buffer(L.ERRCAP)
        rc = L.zke_wtns_prove(self._h, wtns_file, len(wtns_file), rs, proof, publics, err, L.ERRCAP)
        if rc != 0:
            raise L.ZkeError(err.value.decode())
        return proof.raw, publics.raw[: 32 * npub]

    def submit(self, packed_inputs, batch: int, rs: bytes | None = None):
        """Pipelined fullprove: enqueue one batch (at most two in flight); pair with collect()."""
        err = ctypes.create_string_buffer(L.ERRCAP)
        if L.zke_fullprove_submit(self._h, packed_inputs, batch, rs, err, L.ERRCAP) != 0:
            raise L.ZkeError(err.value.decode())
        self._pending.append(batch)

    def collect(self, raise_on_fail: bool = True):
        batch = self._pending.pop(0)
        npub = self.n_public
        proofs = ctypes.create_string_buffer(256 * batch)
        publics = ctypes.create_string_buffer(max(1, 32 * npub * batch))
        status = (ctypes.c_int32 * batch)()
        err = ctypes.create_string_buffer(L.ERRCAP)
        rc = L.zke_fullprove_collect(self._h, proofs, publics, status, err, L.ERRCAP)
        if rc < 0:
            raise L.ZkeError(err.value.decode())
        if rc > 0 and raise_on_fail:
            raise AssertFailed(err.value.decode())
        return proofs.raw, publics.raw[: 32 * npub * batch], list(status)

    def fullprove(self, packed_inputs, batch: int, rs: bytes | None = None, raise_on_fail: bool = True):
        npub = self.n_public
        proofs = ctypes.create_string_buffer(256 * batch)
        publics = ctypes.create_string_buffer(max(1, 32 * npub * batch))
        status = (ctypes.c_int32 * batch)()
        err = ctypes.create_string_buffer(L.ERRCAP)
        rc = L.zke_fullprove(self._h, packed_inputs, batch, rs, proofs, publics, status, err, L.ERRCAP)
        if rc < 0:
            raise L.ZkeError(err.value.decode())
        if rc > 0 and raise_on_fail:
            raise AssertFailed(err.value.decode())
        return proofs.raw, publics.raw[: 32 * npub * batch], list(status)


def proof_to_json(proof256: bytes, publics: bytes, n_public: int):
    """256-byte proof + packed public signals -> (proof.json object, publicSignals list) in snarkjs format."""
    pl, sl = L.c_size_t(4096), L.c_size_t(80 * max(1, n_public) + 16)
    pj, sj = ctypes.create_string_buffer(pl.value), ctypes.create_string_buffer(sl.value)
    if L.zke_proof_to_json(proof256, publics, n_public, pj, ctypes.byref(pl), sj, ctypes.byref(sl)) != 0:
        raise L.ZkeError("proof_to_json failed")
    return json.loads(pj.value.decode()), json.loads(sj.value.decode())


def verify_batch(vkey: dict, public_signals_list, proofs, rand: bytes | None = None) -> list:
    """n proofs under one verification key with ONE randomised product of pairings (n + 3 Miller loops, one final
    exponentiation) - SURVEY 8(f) rank 4; per-proof verdicts, identical to [verify(vkey, s, p) for ...] (a false proof slips
    through with probability ~2^-128 over `rand`, 16 fresh bytes per proof; default os.urandom)."""
    import os
    n = len(proofs)
    if len(public_signals_list) != n:
        raise ValueError("one public-signal list per proof")
    if n == 0:
        return []
    rand = os.urandom(16 * n) if rand is None else rand
    if len(rand) != 16 * n:
        raise ValueError("rand must hold 16 bytes per proof")
    err = ctypes.create_string_buffer(L.ERRCAP)
    ok = ctypes.create_string_buffer(n)
    rbuf = ctypes.create_string_buffer(bytes(rand), 16 * n)
    rc = L.zke_verify_batch_json(json.dumps(vkey).encode(),
                                 json.dumps([[str(s) for s in sig] for sig in public_signals_list]).encode(),
                                 json.dumps(list(proofs)).encode(), ctypes.cast(rbuf, ctypes.c_void_p), ctypes.cast(ok, ctypes.c_void_p),
                                 err, L.ERRCAP)
    if rc < 0:
        raise L.ZkeError(err.value.decode())
    return [b == 1 for b in ok.raw]


def verify(vkey: dict, public_signals, proof: dict) -> bool:
    """snarkjs.groth16.verify(vkey, publicSignals, proof) (/root/reference/packages/helpers/src/chunked-zkey.ts:101)."""
    err = ctypes.create_string_buffer(L.ERRCAP)
    rc = L.zke_verify_json(json.dumps(vkey).encode(), json.dumps([str(s) for s in public_signals]).encode(),
                           json.dumps(proof).encode(), err, L.ERRCAP)
    if rc < 0:
        raise L.ZkeError(err.value.decode())
    return rc == 1
