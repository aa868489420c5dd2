"""B200-native KZG backend: Python host-side mirror of the reference's `DASContext`
(crates/eip7594/src/lib.rs:41-88) over the C ABI of libc_eth_kzg_b200.so (include/c_eth_kzg.h).

The directory name is not an importable identifier; load it with
    importlib.util.spec_from_file_location("eth_kzg_b200", "<repo>/rust-eth-kzg_b200/__init__.py")
(tests/conftest.py and __graft_entry__.py do exactly that)."""
from .eth_kzg import (BYTES_PER_BLOB, BYTES_PER_CELL, BYTES_PER_COMMITMENT, BYTES_PER_PROOF, CELLS_PER_EXT_BLOB,  # noqa: F401
                      DASContext, KzgError, build_library, library_path, load_library)
