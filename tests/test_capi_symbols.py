"""CPU: the C-ABI library loads and exports every symbol include/fp8_b200.h declares.
No compute entry point is called (there is no GPU here)."""
import os

from _util import LIB_PATH, capi, declared_symbols


def test_library_is_built():
    assert os.path.exists(LIB_PATH), "run python fp8-mps-metal_b200/build.py"


def test_every_declared_symbol_is_exported():
    names = declared_symbols()
    assert len(names) >= 11 and len(set(names)) == len(names)
    L = capi()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/fp8_b200.h but not exported"


def test_version_and_status_strings():
    L = capi()
    assert L.fp8b_version() == 100
    assert L.fp8b_status_string(0) == b"ok"
    for code in (-1, -2, -3, -4):
        assert len(L.fp8b_status_string(code)) > 3
    assert L.fp8b_launch_count() == 0 or L.fp8b_launch_count() > 0
    assert L.fp8b_scaled_mm_workspace_bytes(1, 4096, 14336) == 0


def test_extension_imports_and_refuses_cpu_tensors():
    """The torch extension loads without a GPU and fails loudly instead of computing on the CPU."""
    import pytest
    import torch
    import fp8_mps_native
    lib = fp8_mps_native._get_lib()
    assert lib.version() == 100
    with pytest.raises(RuntimeError):
        lib.fp8_encode(torch.zeros(8))
    with pytest.raises(RuntimeError):
        lib.fp8_scaled_mm(torch.zeros(2, 16, dtype=torch.uint8), torch.zeros(3, 16, dtype=torch.uint8),
                          torch.ones(1), torch.ones(1))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            fp8_mps_native.fp8_encode(torch.zeros(8))          # no CUDA device -> error, not a CPU path
