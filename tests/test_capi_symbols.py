"""CPU-side boundary checks: the C-ABI library loads without a GPU and exports every symbol
include/c_eth_kzg.h declares (no compute calls here); context creation fails loudly without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(pkg):
    path = pkg.library_path()
    if not os.path.exists(path):
        pkg.build_library()
    return ctypes.CDLL(path)


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "c_eth_kzg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(eth_kzg_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(lib):
    syms = _declared_symbols()
    # the 13 functions + 3 constants of bindings/c/src/lib.rs:79-569 must all be there
    reference = ["eth_kzg_das_context_new", "eth_kzg_das_context_free", "eth_kzg_free_error_message", "eth_kzg_blob_to_kzg_commitment",
                 "eth_kzg_compute_cells_and_kzg_proofs", "eth_kzg_compute_cells", "eth_kzg_verify_cell_kzg_proof_batch",
                 "eth_kzg_recover_cells_and_proofs", "eth_kzg_compute_kzg_proof", "eth_kzg_compute_blob_kzg_proof", "eth_kzg_verify_kzg_proof",
                 "eth_kzg_verify_blob_kzg_proof", "eth_kzg_verify_blob_kzg_proof_batch", "eth_kzg_constant_bytes_per_cell",
                 "eth_kzg_constant_bytes_per_proof", "eth_kzg_constant_cells_per_ext_blob"]
    assert set(reference) <= set(syms)
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s


def test_constants(lib):
    for name, v in (("eth_kzg_constant_bytes_per_cell", 2048), ("eth_kzg_constant_bytes_per_proof", 48), ("eth_kzg_constant_cells_per_ext_blob", 128)):
        f = getattr(lib, name)
        f.restype = ctypes.c_uint64
        assert f() == v


def test_null_safe_frees(lib):
    lib.eth_kzg_das_context_free.argtypes = [ctypes.c_void_p]
    lib.eth_kzg_free_error_message.argtypes = [ctypes.c_void_p]
    lib.eth_kzg_das_context_free(None)      # bindings/c/src/lib.rs:109 null-safe
    lib.eth_kzg_free_error_message(None)    # bindings/c/src/lib.rs:171 null-safe


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.KzgError):
        pkg.DASContext()


def _build_consumer(pkg):
    import subprocess
    libdir = os.path.dirname(pkg.library_path())
    exe = os.path.join(ROOT, "tests", "capi", "consumer")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "capi", "consumer.c"), "-L" + libdir, "-lc_eth_kzg_b200", "-Wl,-rpath," + libdir])
    return exe


def test_c_consumer_builds_and_runs_without_gpu(lib, pkg):
    """the header is valid C99 and a C program links against the library; without a device the context is refused (no fallback)"""
    import subprocess
    exe = _build_consumer(pkg)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_c_consumer_on_gpu(lib, pkg):
    import subprocess
    exe = _build_consumer(pkg)
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "gpu path ok" in r.stdout, r.stdout + r.stderr


def test_device_shard_bounds(lib):
    """csrc/kzg_multi.cu: the shards of a multi-device batch cover every item once, in order, in whole blob groups (32 items; 8 where
    a device's share is at most 80 items, the range of the cooperative latency-mode G1-NTT kernel)"""
    f = lib.eth_kzg_b200_debug_shard_bounds
    f.argtypes = [ctypes.c_uint64] * 3 + [ctypes.POINTER(ctypes.c_uint64)] * 2
    f.restype = None
    for n in (0, 1, 8, 9, 31, 32, 33, 64, 100, 128, 640, 641, 1000, 1024, 4097):
        for parts in (1, 2, 3, 4, 8):
            gw = 8 if -(-n // parts) <= 80 else 32
            spans = []
            for i in range(parts):
                lo, cnt = ctypes.c_uint64(), ctypes.c_uint64()
                f(n, parts, i, ctypes.byref(lo), ctypes.byref(cnt))
                spans.append((lo.value, cnt.value))
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (lo, c), (lo2, _) in zip(spans, spans[1:]):
                assert lo + c == lo2
            for lo, c in spans[:-1]:
                assert lo % gw == 0
            full = [c for _, c in spans if c and c % gw == 0]
            if full:
                assert max(full) - min(full) <= gw
            if n == 64 and parts == 8:
                assert [c for _, c in spans] == [8] * 8
