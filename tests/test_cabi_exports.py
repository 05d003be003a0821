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
    assert lib.pk_abi_version() == _cabi.PK_ABI_VERSION == 3


def test_struct_layouts_match_the_header_constants():
    with open(os.path.join(ROOT, "include", "pink_b200.h")) as fh:
        text = fh.read()
    for name in ["PK_MAX_JOINTS", "PK_MAX_NV", "PK_MAX_FRAMES", "PK_MAX_TASKS", "PK_MAX_SHARED", "PK_MAX_INEQ_ROWS",
                 "PK_MAX_EQ_ROWS", "PK_MAX_BARRIERS", "PK_MAX_CONSTRAINTS", "PK_MAX_PAIRS", "PK_ABI_VERSION"]:
        value = int(re.search(rf"#define {name} (\d+)", text).group(1))
        assert getattr(_cabi, name) == value
    # the ctypes mirrors have the sizes the compiler gave the C structs
    import __graft_entry__

    __graft_entry__.build()
    lib = ctypes.CDLL(_cabi.library_path())
    for which, struct in enumerate((_cabi.PkModelDesc, _cabi.PkTaskDesc, _cabi.PkBarrierDesc, _cabi.PkProblemDesc)):
        assert lib.pk_struct_size(which) == ctypes.sizeof(struct), struct.__name__
    assert ctypes.sizeof(_cabi.PkTaskDesc) == 60


def test_compute_without_gpu_fails_loudly():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pink_b200.engine import require_cuda

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        require_cuda()


def test_c_entry_point_without_gpu_reports_an_error():
    """Straight through the C ABI on a CPU-only box: a non-zero return code and a message,
    never a silent fallback."""
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pink_b200.robots import load_robot_description

    lib = _cabi.load()
    holder = _cabi.ModelDescHolder(load_robot_description("ur5_description").model.table())
    handle = ctypes.c_void_p()
    rc = lib.pk_model_create(ctypes.byref(holder.desc), 0, ctypes.byref(handle))
    assert rc != 0 and not handle.value
    lib.pk_last_error.restype = ctypes.c_char_p
    assert len(lib.pk_last_error()) > 0
