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


def test_resample_workspace_planning_is_host_only():
    """avmoe_resample_workspace_bytes is pure arithmetic (no GPU): sizes grow with the batch, cover the padded operands, and bad
    geometry is reported instead of planned."""
    lib = _lib.load()
    lib.avmoe_resample_workspace_bytes.restype = ctypes.c_int64

    def need(BT, N, Nx, C, Cx, dtype, bwd):
        a = _lib.ResampleArgs()
        a.dtype, a.BT, a.N, a.Nx, a.C, a.Cx = dtype, BT, N, Nx, C, Cx
        return lib.avmoe_resample_workspace_bytes(ctypes.byref(a), int(bwd))

    # stage-4 audio adapter (N=64 audio tokens <- Nx=36 visual tokens, bf16): the 36-token contraction is padded to 40
    fwd = need(2, 64, 36, 768, 1536, 1, False)
    assert fwd >= 64 * 40 * 2 + 768 * 1536 * 2 + 2 * 768 * 40 * 2
    assert need(4, 64, 36, 768, 1536, 1, False) > fwd
    assert need(2, 64, 36, 768, 1536, 1, True) > 0
    assert need(2, 64, 36, 768, 1536, 0, False) > fwd          # fp32 operands are twice as wide
    assert need(0, 64, 36, 768, 1536, 1, False) == -1


def test_token_update_support_probe():
    """avmoe_rf_tok_supported / avmoe_rf_tok_parts: geometry probes of the row-fused latent-token update, host only"""
    lib = _lib.load()
    assert lib.avmoe_rf_tok_supported(4096, 96, 32) == 1 and lib.avmoe_rf_tok_supported(1024, 192, 87) == 1
    assert lib.avmoe_rf_tok_supported(36, 768, 32) == 0, "Nk must be a multiple of 16"
    assert lib.avmoe_rf_tok_supported(4096, 768, 32) == 0, "rows wider than 384 channels do not fit the double-buffered tiles"
    assert lib.avmoe_rf_tok_supported(4096, 96, 100) == 0
    assert 1 <= lib.avmoe_rf_tok_parts(160, 4096) <= 64 and lib.avmoe_rf_tok_parts(2, 16) == 1


def test_graph_adapters_rejects_cpu_inputs():
    import pytest
    import torch
    from avmoe_b200.adapters import graph_adapters
    with pytest.raises(RuntimeError):
        graph_adapters(torch.nn.Identity(), (torch.zeros(2, 4, 8, 1), torch.zeros(2, 4, 8, 1)))
    with pytest.raises(ValueError):
        graph_adapters((torch.nn.Identity(), torch.nn.Identity()), ((torch.zeros(1),),))
