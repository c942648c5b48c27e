"""The native-evaluation table of the Sha256compression gadget (circuit.hpp: ShaBlock, ZKE_ARR_SHA_BLOCKS): every signal the
gadget creates must be the recorded bit of the recorded 64-bit quantity of a plain SHA-256 compression.  Checked here on
the CPU: the oracle walks the ordinary witness program, and each descriptor is compared with an independent Python
compression - the same quantities the GPU's native op (witness.cu: sha_coop) computes.  Role in the reference: the
circomlib sha256compression template inside Sha256General (/root/reference/packages/circuits/lib/sha.circom:158,247)."""
import ctypes
import hashlib

import zkemail_b200 as z
from zkemail_b200 import _lib as L
from zkutil import oracle_witness

K = [0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
     0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
     0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
     0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
     0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
     0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2]
M = 0xffffffff
rotr = lambda x, r: ((x >> r) | (x << (32 - r))) & M
(S1MID, S1, S0MID, S0, WSUM, BS1MID, BS1, CH, T1SUM, BS0MID, BS0, MAJMID, MAJ, T2SUM, SUME, SUMA, FS) = range(17)


def quantities(hin_words, w16):
    q = {}
    w = list(w16)
    for t in range(16, 64):
        x, y = w[t - 2], w[t - 15]
        q[S1MID, t] = rotr(x, 19) & (x >> 10); q[S1, t] = rotr(x, 17) ^ rotr(x, 19) ^ (x >> 10)
        q[S0MID, t] = rotr(y, 18) & (y >> 3); q[S0, t] = rotr(y, 7) ^ rotr(y, 18) ^ (y >> 3)
        q[WSUM, t] = q[S1, t] + w[t - 7] + q[S0, t] + w[t - 16]
        w.append(q[WSUM, t] & M)
    a, b, c, d, e, f, g, h = hin_words
    for t in range(64):
        q[BS1MID, t] = rotr(e, 11) & rotr(e, 25); q[BS1, t] = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)
        q[CH, t] = (e & f) ^ (~e & g & M)
        q[T1SUM, t] = h + q[BS1, t] + q[CH, t] + K[t] + w[t]
        q[BS0MID, t] = rotr(a, 13) & rotr(a, 22); q[BS0, t] = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)
        q[MAJMID, t] = b & c; q[MAJ, t] = (a & b) ^ (a & c) ^ (b & c)
        q[T2SUM, t] = q[BS0, t] + q[MAJ, t]
        t1 = q[T1SUM, t] & M
        q[SUME, t] = d + t1; q[SUMA, t] = t1 + (q[T2SUM, t] & M)
        h, g, f, e, d, c, b, a = g, f, e, q[SUME, t] & M, c, b, a, q[SUMA, t] & M
    for i, (hi, s) in enumerate(zip(hin_words, (a, b, c, d, e, f, g, h))):
        q[FS, i] = hi + s
    return q


def _blocks(circuit):
    n = L.c_size_t()
    p = L.zke_circuit_array(circuit.handle, L.ARR_SHA_BLOCKS, ctypes.byref(n))
    flat = list((ctypes.c_uint32 * n.value).from_address(p)) if n.value else [0]
    out, pos = [], 1
    for _ in range(flat[0]):
        vb, ve, tb, te, nd = flat[pos:pos + 5]; pos += 5
        inputs = flat[pos:pos + 768]; pos += 768
        desc = flat[pos:pos + 2 * nd]; pos += 2 * nd
        out.append((vb, ve, tb, te, inputs, desc))
    assert pos == len(flat)
    return out


def _check(circuit, inputs):
    w = oracle_witness(circuit, inputs)
    blocks = _blocks(circuit)
    total = 0
    for vb, ve, tb, te, ins, desc in blocks:
        assert len(desc) == 2 * (ve - vb) and sorted(desc[0::2]) == list(range(vb, ve)), "one descriptor per created signal"
        bit = lambda s: 0 if s == 0xfffffffe else (1 if s == 0xffffffff else w[s])
        hin = [sum(bit(ins[32 * i + k]) << k for k in range(32)) for i in range(8)]                    # LSB first
        w16 = [sum(bit(ins[256 + 32 * t + 31 - k]) << k for k in range(32)) for t in range(16)]         # MSB first
        q = quantities(hin, w16)
        for var, qk in zip(desc[0::2], desc[1::2]):
            grp, idx, k = (qk >> 8) // 64, (qk >> 8) % 64, qk & 255
            assert w[var] == (q[grp, idx] >> k) & 1, (var, grp, idx, k)
            total += 1
    return blocks, total, w


def test_sha_blocks_cover_every_signal_of_the_gadget():
    c = z.Circuit("Sha256Bytes", [128])
    msg = bytes(range(97))
    padded, plen = z.sha256_pad(msg, 128)
    blocks, total, w = _check(c, {"paddedIn": list(padded), "paddedInLength": plen})
    assert len(blocks) == 2 and total > 50000
    # the first block's chaining input is the constant IV, the second one's are the first block's output signals
    assert all(s in (0xfffffffe, 0xffffffff) for s in blocks[0][4][:256])
    assert all(blocks[0][0] <= s < blocks[0][1] for s in blocks[1][4][:256])
    first, count, _ = c.groups["out"]
    bits = [w[first + i] for i in range(count)]
    assert bits == [(b >> (7 - j)) & 1 for b in hashlib.sha256(msg).digest() for j in range(8)]


def test_sha_blocks_with_a_prehash_input():
    c = z.Circuit("Sha256BytesPartial", [64])
    body = b"a" * 64 + b"partial tail"
    pre = z.partial_sha(body[:64], 64)
    padded, plen = z.sha256_pad(body[64:], 64)
    blocks, total, _ = _check(c, {"paddedIn": list(padded), "paddedInLength": plen, "preHash": list(pre)})
    assert len(blocks) == 1 and all(s < 0xfffffffe for s in blocks[0][4][:256])      # chaining input = preHash bit signals
