"""Mirror of /root/reference/packages/helpers/src/binary-format.ts (the functions on the EmailVerifier path)."""
from __future__ import annotations
from .constants import CIRCOM_BIGINT_K, CIRCOM_BIGINT_N


def uint8array_to_char_array(a: bytes) -> list[str]:
    """Uint8ArrayToCharArray (binary-format.ts:44-46)."""
    return [str(x) for x in a]


def bytes_to_bigint(b: bytes) -> int:
    """bytesToBigInt (binary-format.ts:63-69): big-endian."""
    return int.from_bytes(b, "big")


def bigint_to_chunked_bytes(num: int, bits_per_chunk: int, num_chunks: int) -> list[str]:
    """bigIntToChunkedBytes (binary-format.ts:71-79): little-endian limbs, decimal strings."""
    mask = (1 << bits_per_chunk) - 1
    return [str((num >> (i * bits_per_chunk)) & mask) for i in range(num_chunks)]


def to_circom_bigint_bytes(num: int) -> list[str]:
    """toCircomBigIntBytes (binary-format.ts:81-83)."""
    return bigint_to_chunked_bytes(num, CIRCOM_BIGINT_N, CIRCOM_BIGINT_K)


def int64_to_bytes(num: int) -> bytes:
    """int64toBytes (binary-format.ts:141-146): only the low 32 bits are written ("Works only on 32 bit sha text
    lengths"), big-endian in the last four of eight bytes; DataView.setInt32 wraps modulo 2^32."""
    return b"\x00\x00\x00\x00" + (num & 0xFFFFFFFF).to_bytes(4, "big")


def int8_to_bytes(num: int) -> bytes:
    """int8toBytes (binary-format.ts:149-154)."""
    return bytes([num & 0xFF])
