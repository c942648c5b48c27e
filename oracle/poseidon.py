"""Poseidon over BN254 Fr, circomlib parameterisation (x^5 S-box, R_F = 8, R_P by width).

TEST INFRASTRUCTURE (oracle) - never imported by the product path.

Restates circomlib 2.0.5 ``Poseidon(nInputs)`` / circomlibjs 0.1.7 ``buildPoseidon`` (un-vendored;
pins at /root/reference/yarn.lock:3619-3633; call sites /root/reference/packages/circuits/utils/hash.circom:38
and /root/reference/packages/helpers/src/hash.ts:1-25).  The round constants and the MDS matrix are
regenerated with the Poseidon reference Grain-LFSR procedure (field = prime field, S-box = x^alpha,
n = 254 bits, t, R_F, R_P) - circomlib's constants file was produced by exactly this procedure.
circomlib evaluates an algebraically optimised schedule (sparse partial rounds); its *output* is
identical to the plain permutation restated here.

Known answers (published circomlib test vectors): poseidon([1, 2]), poseidon([1]) - see
tests/test_oracle_poseidon.py.
"""
from __future__ import annotations
from functools import lru_cache

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
N_BITS = 254
R_F = 8
# circomlib poseidon.circom: N_ROUNDS_P indexed by t - 2
N_ROUNDS_P = [56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64, 68]


class _Grain:
    def __init__(self, t: int, r_f: int, r_p: int):
        bits = []

        def put(v, n):
            bits.extend(int(c) for c in bin(v)[2:].zfill(n))

        put(1, 2)        # field: prime
        put(0, 4)        # s-box: x^alpha
        put(N_BITS, 12)
        put(t, 12)
        put(r_f, 10)
        put(r_p, 10)
        bits.extend([1] * 30)
        assert len(bits) == 80
        self.s = bits
        for _ in range(160):
            self._step()

    def _step(self) -> int:
        s = self.s
        nb = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(nb)
        return nb

    def _bit(self) -> int:
        while True:
            b1 = self._step()
            b2 = self._step()
            if b1:
                return b2

    def bits(self, n: int) -> int:
        v = 0
        for _ in range(n):
            v = (v << 1) | self._bit()
        return v

    def field_rejection(self) -> int:
        while True:
            v = self.bits(N_BITS)
            if v < R:
                return v


@lru_cache(maxsize=None)
def params(t: int):
    """(round constants [(R_F+R_P)*t], MDS matrix t x t) for state width t."""
    r_p = N_ROUNDS_P[t - 2]
    g = _Grain(t, R_F, r_p)
    rc = [g.field_rejection() for _ in range((R_F + r_p) * t)]
    # Cauchy MDS: M[i][j] = 1 / (x_i + y_j), x/y drawn from the same stream (no rejection, reduced mod r)
    while True:
        vals = [g.bits(N_BITS) % R for _ in range(2 * t)]
        if len(set(vals)) != 2 * t:
            continue
        xs, ys = vals[:t], vals[t:]
        if any((x + y) % R == 0 for x in xs for y in ys):
            continue
        mds = [[pow((x + y) % R, -1, R) for y in ys] for x in xs]
        return rc, mds


def permute(state):
    t = len(state)
    rc, mds = params(t)
    r_p = N_ROUNDS_P[t - 2]
    st = [s % R for s in state]
    k = 0
    for rnd in range(R_F + r_p):
        st = [(s + rc[k + i]) % R for i, s in enumerate(st)]
        k += t
        if rnd < R_F // 2 or rnd >= R_F // 2 + r_p:
            st = [pow(s, 5, R) for s in st]
        else:
            st[0] = pow(st[0], 5, R)
        st = [sum(mds[i][j] * st[j] for j in range(t)) % R for i in range(t)]
    return st


def poseidon(inputs) -> int:
    """circomlib ``Poseidon(n)(inputs)``: state = [0, inputs...], output = state[0]."""
    assert 1 <= len(inputs) <= 16
    return permute([0] + [int(x) for x in inputs])[0]


def poseidon_large(value: int, num_chunks: int, bits_per_chunk: int) -> int:
    """helpers ``poseidonLarge(input, numChunks, bitsPerChunk)``
    (/root/reference/packages/helpers/src/hash.ts:17-25): chunk little-endian, hash once."""
    mask = (1 << bits_per_chunk) - 1
    return poseidon([(value >> (i * bits_per_chunk)) & mask for i in range(num_chunks)])


def poseidon_modular(inputs) -> int:
    """helpers ``poseidonModular(inputs)`` (/root/reference/packages/helpers/src/hash.ts:27-59): hash chunks of 16,
    fold the chunk hashes left to right with Poseidon(2)."""
    if not inputs:
        raise ValueError("No inputs provided")
    out = None
    for start in range(0, len(inputs), 16):
        h = poseidon(inputs[start:start + 16])
        out = h if out is None else poseidon([out, h])
    return out
