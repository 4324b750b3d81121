"""CPU: the C-ABI library builds, loads, and exports every symbol include/bruno_sm100.h declares."""
import ctypes


def test_library_builds_and_exports_declared_symbols():
    from bruno_b200 import build, _lib
    path = build.build()
    protos = _lib.declared_prototypes()
    assert len(protos) >= 8
    lib = ctypes.CDLL(path)
    missing = [n for n in protos if not hasattr(lib, n)]
    assert not missing, missing
    _lib.load()
    assert _lib.query('bruno_abi_version') >= 1
    assert _lib.query('bruno_process_table_floats', 0, 20, 784, 0) == (20 * 7 + 3) * 784


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing on the host."""
    import pytest
    import torch
    from bruno_b200 import _lib
    with pytest.raises(RuntimeError):
        _lib.ptr(torch.zeros(4))
