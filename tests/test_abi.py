"""The C-ABI library loads and exports every symbol include/spparks_cuda.h declares (CPU-only check:
no compute calls)."""
import ctypes
import os
import re

import pytest

from spparks_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "spparks_cuda.h")).read()
    declared = sorted(set(re.findall(r"\b(spk_cuda_[a-z_]+)\s*\(", hdr)))
    assert declared == sorted(_lib.SYMBOLS)
    lib = _built()
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_tables():
    lib = _built()
    assert lib.spk_cuda_abi_version() == 1
    # Lattice::ncolors (src/lattice.cpp:187-228) and maxneigh (src/create_sites.cpp:535-540)
    assert [lib.spk_cuda_ncolors(l) for l in (2, 3, 4, 5, 6)] == [2, 4, 4, 2, 8]
    assert [lib.spk_cuda_maxneigh(l) for l in (2, 3, 4, 5, 6)] == [4, 8, 6, 6, 26]
    assert [lib.spk_cuda_nbasis(l) for l in (2, 3, 4, 5, 6)] == [1, 1, 2, 1, 1]
    assert lib.spk_cuda_ncolors(7) == 0


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from spparks_b200 import LatticeSweep, SpkCudaError
    with pytest.raises(SpkCudaError):
        LatticeSweep("potts", "sc/26n", 8, 8, 8, nspins=5)
