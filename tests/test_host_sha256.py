"""The verifiers' transcript hasher (rust-eth-kzg_b200/csrc/host_sha256.cpp: x86 SHA extensions with a portable
fallback) against hashlib.  Host code only -- runs in the GPU-less container through the C-ABI library's test hook."""
import ctypes
import hashlib
import random

import os


def test_sha256_matches_hashlib(pkg):
    path = pkg.library_path()
    if not os.path.exists(path):
        pkg.build_library()
    lib = ctypes.CDLL(path)
    out = ctypes.create_string_buffer(32)
    rng = random.Random(5)
    sizes = list(range(0, 200)) + [1000, 4096, 65537, 2048 * 33 + 17]
    for n in sizes:
        d = rng.randbytes(n)
        for split in (0, 1, 63, 64, 65, n // 2, n):
            for portable in (0, 1):
                lib.eth_kzg_b200_debug_sha256(d, ctypes.c_uint64(n), ctypes.c_uint64(split), portable, out)
                assert out.raw == hashlib.sha256(d).digest(), (n, split, portable)
