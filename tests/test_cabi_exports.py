"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/pink_b200.h declares (no compute calls here)."""

import ctypes
import os
import re

from pink_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "pink_b200.h")) as fh:
        text = fh.read()
    return sorted(set(re.findall(r"\b(pk_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__

    __graft_entry__.build()
    lib = ctypes.CDLL(_cabi.library_path())
    declared = _declared_symbols()
    assert set(declared) == set(_cabi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    lib.pk_abi_version.restype = ctypes.c_int
    assert lib.pk_abi_version() == 1


def test_struct_layouts_match_the_header_constants():
    with open(os.path.join(ROOT, "include", "pink_b200.h")) as fh:
        text = fh.read()
    for name in ["PK_MAX_JOINTS", "PK_MAX_NV", "PK_MAX_FRAMES", "PK_MAX_TASKS", "PK_MAX_SHARED"]:
        value = int(re.search(rf"#define {name} (\d+)", text).group(1))
        assert getattr(_cabi, name) == value
    # sizeof(PkProblemDesc): 4 + 12*52 + 4*4 + 4 + 5*64*4 + 192*4
    assert ctypes.sizeof(_cabi.PkTaskDesc) == 52
    assert ctypes.sizeof(_cabi.PkProblemDesc) == 4 + 12 * 52 + 16 + 4 + 5 * 64 * 4 + 192 * 4


def test_compute_without_gpu_fails_loudly():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pink_b200.engine import require_cuda

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        require_cuda()
