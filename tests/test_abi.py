"""C-ABI: the shared library loads on a CPU-only machine and exports every symbol include/b2me.h declares;
the ctypes table mirrors the header."""
import ctypes
import os
import re

from conftest import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "b2me.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2me_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    syms = _header_symbols()
    assert len(syms) >= 29
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b2me.h but not exported"


def test_ctypes_table_matches_header(built_lib):
    from MinkowskiEngine import _lib
    assert sorted(_lib.SIGNATURES) == _header_symbols()
    assert _lib.lib.b2me_version() >= 100
    assert _lib.lib.b2me_strerror(-2).decode().startswith("workspace")


def test_sizes_without_gpu(built_lib):
    from MinkowskiEngine._lib import lib
    assert lib.b2me_table_slots(1000) == 2048
    assert lib.b2me_table_bytes(1000) == 2048 * 16
    assert lib.b2me_unique_workspace_bytes(1000, 3) > 1000 * (16 + 4 + 4 + 24)
    assert lib.b2me_tc_supported(27, 384, 32, 384) == 1
    assert lib.b2me_tc_supported(27, 3, 0, 32) == 0          # stem goes to the SIMT kernel
    assert lib.b2me_tc_supported(1, 1024, 0, 3) == 0
    assert lib.b2me_tc_packed_bytes(27, 384, 32, 384, 1) == 27 * 7 * 384 * 128
    assert lib.b2me_tc_packed_bytes(1, 256, 0, 1024, 1) == 4 * 1024 * 128
    assert lib.b2me_tc_packed_bytes(27, 384, 32, 384, 2) == 27 * 13 * 384 * 128   # tf32: 32 channels per chunk


def test_no_cpu_fallback(built_lib):
    """the CUDA-backed package refuses CPU tensors instead of silently computing elsewhere."""
    import pytest
    import torch
    import MinkowskiEngine as ME
    with pytest.raises(ME.B2MEError):
        ME.TensorField(features=torch.zeros(4, 3), coordinates=torch.zeros(4, 4))
    with pytest.raises(ME.B2MEError):
        ME.SparseTensor(torch.zeros(4, 3), coordinates=torch.zeros(4, 4, dtype=torch.int32))
