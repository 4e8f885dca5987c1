"""Host mirror of the transcript ezkl proves with: snark-verifier's `EvmTranscript` (Keccak-256 Fiat-Shamir), the type selected
at /root/reference/src/execute.rs:1608-1609 (`EvmTranscript<G1Affine, _, _, _>`) and driven from
/root/reference/src/pfsys/mod.rs:435-463 (`TranscriptWriterBuffer::init`, `create_proof`, `transcript.finalize()`).

Rules restated (snark-verifier 0.1.1 @ zkonduit/snark-verifier#508d96b2, system/halo2/transcript/evm.rs — not vendored; the byte
encodings are confirmed by the reference's own `tests/assets/proof.json`):
  * the sponge is a byte buffer; `common_ec_point` appends x || y, `common_scalar` appends the scalar, all 32-byte BIG-endian
    canonical integers (`fe_to_bytes32`, /root/reference/src/pfsys/mod.rs:114-126);
  * `squeeze_challenge`: hash = Keccak-256(buffer, plus one 0x01 byte when the buffer is exactly 32 bytes long, i.e. when nothing
    was absorbed since the previous squeeze); the buffer becomes the 32 hash bytes; the challenge is the hash read as a big-endian
    integer reduced modulo r;
  * `write_ec_point` / `write_scalar` absorb and also append the same bytes to the proof stream.
The transcript stays on the CPU in the Rust prover too: it hashes a few kilobytes per proof.
Keccak-256 here is the original Keccak padding (0x01 ... 0x80), not SHA3-256's (hashlib has only the latter).
"""
from __future__ import annotations

import numpy as np

from . import fields as F

_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001, 0x8000000080008081,
       0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B,
       0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A, 0x8000000080008081,
       0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M = (1 << 64) - 1


def _rol(v, s):
    return ((v << s) | (v >> (64 - s))) & _M if s else v


def _keccak_f(a):
    for rc in _RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    return a


def keccak256(data: bytes) -> bytes:
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i: off + 8 * i + 8], "little")
        a = _keccak_f(a)
    return b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))


def fq_from_limbs(a) -> int:
    a = np.asarray(a, dtype=np.uint64).reshape(4)
    m = sum(int(a[i]) << (64 * i) for i in range(4))
    return m * pow(1 << 256, -1, F.FQ_MODULUS) % F.FQ_MODULUS


def point_bytes(p) -> bytes:
    """G1 point in wire form (affine [8] or normalised Jacobian [12], Montgomery limbs) -> x || y, 32-byte big-endian each;
    the identity encodes as 64 zero bytes."""
    p = np.asarray(p, dtype=np.uint64).reshape(-1)
    if p.size == 12 and not p[8:].any():
        return bytes(64)
    return fq_from_limbs(p[:4]).to_bytes(32, "big") + fq_from_limbs(p[4:8]).to_bytes(32, "big")


def scalar_bytes(s) -> bytes:
    """Fr element (python int, or wire limbs) -> 32-byte big-endian canonical."""
    if not isinstance(s, int):
        s = F.fr_from_limbs(s)
    return (s % F.FR_MODULUS).to_bytes(32, "big")


class EvmTranscriptWrite:
    """`EvmTranscript` on the prover side (TranscriptWrite + TranscriptWriterBuffer): absorbs, squeezes, accumulates proof bytes."""

    def __init__(self):
        self.buf = bytearray()
        self.stream = bytearray()

    def common_ec_point(self, p):
        self.buf += point_bytes(p)

    def common_scalar(self, s):
        self.buf += scalar_bytes(s)

    def write_ec_point(self, p):
        b = point_bytes(p)
        self.buf += b
        self.stream += b

    def write_scalar(self, s):
        b = scalar_bytes(s)
        self.buf += b
        self.stream += b

    def squeeze_challenge(self) -> int:
        data = bytes(self.buf) + (b"\x01" if len(self.buf) == 32 else b"")
        h = keccak256(data)
        self.buf = bytearray(h)
        return int.from_bytes(h, "big") % F.FR_MODULUS

    def finalize(self) -> bytes:
        return bytes(self.stream)


class EvmTranscriptRead:
    """The verifier side over the proof bytes (TranscriptRead): same sponge, points / scalars come from the stream."""

    def __init__(self, proof: bytes):
        self.buf = bytearray()
        self.proof, self.pos = bytes(proof), 0

    def common_scalar(self, s):
        self.buf += scalar_bytes(s)

    def common_ec_point(self, p):
        self.buf += point_bytes(p)

    def read_ec_point(self):
        b = self.proof[self.pos:self.pos + 64]
        if len(b) != 64:
            raise ValueError("proof truncated")
        self.pos += 64
        self.buf += b
        x, y = int.from_bytes(b[:32], "big"), int.from_bytes(b[32:], "big")
        if x >= F.FQ_MODULUS or y >= F.FQ_MODULUS or ((x, y) != (0, 0) and (y * y - x * x * x - 3) % F.FQ_MODULUS):
            raise ValueError("invalid point in proof")
        return None if (x, y) == (0, 0) else (x, y)

    def read_scalar(self) -> int:
        b = self.proof[self.pos:self.pos + 32]
        if len(b) != 32:
            raise ValueError("proof truncated")
        self.pos += 32
        v = int.from_bytes(b, "big")
        if v >= F.FR_MODULUS:
            raise ValueError("non-canonical scalar in proof")
        self.buf += b
        return v

    squeeze_challenge = EvmTranscriptWrite.squeeze_challenge
