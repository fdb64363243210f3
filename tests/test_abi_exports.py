"""CPU: the C-ABI shared library builds/loads and exports every symbol include/avmoe_b200.h declares."""
import ctypes

from avmoe_b200 import _lib


def test_library_loads_and_reports_version():
    lib = _lib.load()
    assert lib.avmoe_abi_version() == 1
    assert b"sm_100a" in lib.avmoe_build_info()


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    names = _lib.exported_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/avmoe_b200.h but not exported"


def test_no_cpu_path():
    import pytest
    import torch
    from avmoe_b200 import ops
    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(4, 4), torch.zeros(4, 4))
