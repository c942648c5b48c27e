"""Mirror of /root/reference/packages/helpers/src/sha-utils.ts and of the `cacheState()` midstate export of
/root/reference/packages/helpers/src/lib/fast-sha256.ts:240-251."""
from __future__ import annotations
import hashlib
import struct
from .binary_format import int64_to_bytes

_K = [
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2,
]
_IV = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]
_M = 0xFFFFFFFF


def _rotr(x, n):
    return ((x >> n) | (x << (32 - n))) & _M


def sha256_compress(state, block: bytes):
    w = list(struct.unpack(">16I", block))
    for t in range(16, 64):
        s0 = _rotr(w[t - 15], 7) ^ _rotr(w[t - 15], 18) ^ (w[t - 15] >> 3)
        s1 = _rotr(w[t - 2], 17) ^ _rotr(w[t - 2], 19) ^ (w[t - 2] >> 10)
        w.append((w[t - 16] + s0 + w[t - 7] + s1) & _M)
    a, b, c, d, e, f, g, h = state
    for t in range(64):
        t1 = (h + (_rotr(e, 6) ^ _rotr(e, 11) ^ _rotr(e, 25)) + ((e & f) ^ (~e & g)) + _K[t] + w[t]) & _M
        t2 = ((_rotr(a, 2) ^ _rotr(a, 13) ^ _rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c))) & _M
        h, g, f, e, d, c, b, a = g, f, e, (d + t1) & _M, c, b, a, (t1 + t2) & _M
    return [(x + y) & _M for x, y in zip(state, (a, b, c, d, e, f, g, h))]


def find_index_in_uint8array(array: bytes, selector: bytes) -> int:
    """findIndexInUint8Array (sha-utils.ts:5-20) - deliberately the reference's scan (no back-tracking on a
    partial match), not bytes.find."""
    i = j = 0
    while i < len(array):
        if array[i] == selector[j]:
            j += 1
            if j == len(selector):
                return i - j + 1
        else:
            j = 0
        i += 1
    return -1


def pad_uint8array_with_zeros(array: bytes, length: int) -> bytes:
    """padUint8ArrayWithZeros (sha-utils.ts:22-28)."""
    return array + b"\x00" * max(0, length - len(array))


def partial_sha(msg: bytes, msg_len: int) -> bytes:
    """partialSha (sha-utils.ts:82-85): SHA-256 state after `msg_len` bytes (a multiple of 64), serialised as
    8 big-endian words (fast-sha256.ts cacheState)."""
    assert msg_len % 64 == 0 and msg_len <= len(msg)
    st = list(_IV)
    for off in range(0, msg_len, 64):
        st = sha256_compress(st, msg[off:off + 64])
    return struct.pack(">8I", *st)


def generate_partial_sha(body: bytes, body_length: int, selector_string: str | None, max_remaining_body_length: int):
    """generatePartialSHA (sha-utils.ts:30-76).  Returns (precomputedSha, bodyRemaining, bodyRemainingLength)."""
    selector_index = 0
    if selector_string:
        selector = selector_string.encode("utf-8")
        selector_index = find_index_in_uint8array(body, selector)
        if selector_index == -1:
            raise ValueError(f'SHA precompute selector "{selector_string}" not found in the body')
    sha_cutoff_index = (selector_index // 64) * 64
    precompute_text = body[:sha_cutoff_index]
    body_remaining = body[sha_cutoff_index:]
    body_remaining_length = body_length - len(precompute_text)
    if body_remaining_length > max_remaining_body_length:
        raise ValueError(
            f"Remaining body {body_remaining_length} after the selector is longer than max ({max_remaining_body_length})")
    if len(body_remaining) % 64 != 0:
        raise ValueError("Remaining body was not padded correctly with int64s")
    body_remaining = pad_uint8array_with_zeros(body_remaining, max_remaining_body_length)
    precomputed_sha = partial_sha(precompute_text, sha_cutoff_index)
    return precomputed_sha, body_remaining, body_remaining_length


def sha_hash(data: bytes) -> bytes:
    """shaHash (sha-utils.ts:78-80)."""
    return hashlib.sha256(data).digest()


def sha256_pad(message: bytes, max_sha_bytes: int):
    """sha256Pad (sha-utils.ts:88-111): message || 0x80 || zeros || 8-byte length (low 32 bits only), then 8-byte zero
    groups up to maxShaBytes.  Returns (padded, messageLen) with messageLen the SHA-padded length."""
    msg_len_bits = len(message) * 8
    msg_len_bytes = int64_to_bytes(msg_len_bits)
    res = message + b"\x80"
    while (len(res) * 8 + len(msg_len_bytes) * 8) % 512 != 0:
        res += b"\x00"
    res += msg_len_bytes
    assert (len(res) * 8) % 512 == 0, "Padding did not complete properly!"
    message_len = len(res)
    while len(res) < max_sha_bytes:
        res += int64_to_bytes(0)
    assert len(res) == max_sha_bytes, (
        f"Padding to max length did not complete properly! Your padded message is {len(res)} long but max is {max_sha_bytes}!")
    return res, message_len
