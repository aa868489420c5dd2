"""Deterministic synthetic blobs (SURVEY.md §8d "Synthetic blob generator"): element i of blob b is
SHA-256(u64_le(seed) || u32_le(b) || u32_le(i)) with the top two bits cleared, so every element is a
canonical Fr (< 2^254 < r) and all 255-bit windows are dense.  Plus the edge blobs every run includes."""
import hashlib
import struct

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
SEED = 0xB200


def blob(b, seed=SEED):
    out = bytearray(131072)
    pre = struct.pack("<Q", seed) + struct.pack("<I", b)
    for i in range(4096):
        h = bytearray(hashlib.sha256(pre + struct.pack("<I", i)).digest())
        h[0] &= 0x3F
        out[32 * i:32 * i + 32] = h
    return bytes(out)


def blobs(n, first=0, seed=SEED):
    return b"".join(blob(first + i, seed) for i in range(n))


def edge_blobs():
    """all-zero, all-(r-1), constant (every proof is the identity), and the reference bench's
    dummy_blob = BE32(-i mod r) (crates/eip7594/benches/benchmark-mt.rs:10-17)"""
    zero = bytes(131072)
    rm1 = (R - 1).to_bytes(32, "big") * 4096
    const = (0x1234567).to_bytes(32, "big") * 4096
    dummy = b"".join(((-i) % R).to_bytes(32, "big") for i in range(4096))
    return [zero, rm1, const, dummy]
