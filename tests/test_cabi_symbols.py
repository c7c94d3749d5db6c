"""CPU-side check of the drop-in boundary: libvdbcuda.so loads and exports every symbol that
include/vdb_cuda.h declares (no compute calls - there is no GPU here)."""
import os
import re

from vectordb_retrieval_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "vdb_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vdb_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    build.build()
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in vdb_cuda.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES out of sync with the header"
    assert lib.vdb_abi_version() == 1
    # pure host helpers are safe to call without a device
    assert lib.vdb_flat_kpad(50) == 64 and lib.vdb_flat_kpad(128) == 128
    assert lib.vdb_flat_npad(1000) == 1024 and lib.vdb_flat_nqpad(1) == 256
    assert lib.vdb_ivf_d4(50) == 13
