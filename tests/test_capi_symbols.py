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


def test_compute_entry_points_fail_loudly_without_a_device():
    """No GPU in this process: every compute entry point returns a negative status (and launches nothing)
    instead of computing somewhere else.  Argument validation does not need a device either."""
    import ctypes
    import torch
    L = capi()
    if torch.cuda.is_available():
        import pytest
        pytest.skip("covered by the GPU suite")
    buf = (ctypes.c_uint8 * 256)()
    ptr = ctypes.cast(buf, ctypes.c_void_p)
    n0 = L.fp8b_launch_count()
    assert L.fp8b_encode(ptr, 0, ptr, 16, None, None) < 0
    assert L.fp8b_dequant_f16(ptr, ptr, 16, None, None) < 0
    assert L.fp8b_dequant(ptr, ptr, 2, 16, None) < 0
    assert L.fp8b_amax_scale(ptr, 0, 16, ptr, ptr, ptr, None) < 0
    assert L.fp8b_quantize_rows(ptr, 0, 2, 8, ptr, ptr, None) < 0
    assert L.fp8b_scaled_mm(ptr, ptr, ptr, 0, 2, 2, 16, 2, ptr, 1, ptr, 1, None, 0, None, None, 0, 0, None) < 0
    assert L.fp8b_launch_count() == n0
    # pure validation errors
    assert L.fp8b_encode(None, 0, ptr, 16, None, None) == -1
    assert L.fp8b_encode(ptr, 7, ptr, 16, None, None) == -1
    assert L.fp8b_scaled_mm(ptr, ptr, ptr, 0, 2, 2, 16, 1, ptr, 1, ptr, 1, None, 0, None, None, 0, 0, None) == -1      # ldc < N
    assert L.fp8b_scaled_mm(ptr, ptr, ptr, 0, 2, 2, 16, 2, ptr, 3, ptr, 1, None, 0, None, None, 0, 0, None) == -1      # bad scale length
    assert L.fp8b_set_option(99, 1) == -1 and L.fp8b_set_option(1, 0) == 0 and L.fp8b_get_option(0) == 1
    # empty problems are accepted and do nothing
    assert L.fp8b_encode(None, 0, None, 0, None, None) == 0
    assert L.fp8b_scaled_mm(None, None, None, 0, 0, 4, 16, 4, None, 1, None, 1, None, 0, None, None, 0, 0, None) == 0


def test_bridge_errors_are_python_exceptions():
    """Failures inside the torch extension surface as RuntimeError (the reference raises std::runtime_error /
    TORCH_CHECK, fp8_bridge.cpp:96-101,174-177) -- including the ones with a formatted message, which once took
    the interpreter down."""
    import pytest
    import torch
    import fp8_metal
    with pytest.raises(RuntimeError, match="invalid argument"):
        fp8_metal.set_option(99, 1)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        fp8_metal.fp8_dequantize(torch.zeros(4, dtype=torch.uint8), torch.ones(1))
    with pytest.raises(RuntimeError, match="uint8"):
        fp8_metal.fp8_dequantize(torch.zeros(4, dtype=torch.int32), torch.ones(1))


def test_ctypes_bindings_match_the_header_prototypes():
    """Guards the drop-in boundary against drift: for every prototype in include/fp8_b200.h the binding the tests
    (and INTEGRATION.md's example) use must pass the same number of arguments, with pointer / integer kinds in
    the same positions."""
    import ctypes
    import re
    from _util import HEADER
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    L = capi()
    checked = 0
    for m in re.finditer(r"FP8B_API\s+[\w\s\*]+?\b(fp8b_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        plist = [] if params in ("void", "") else [p.strip() for p in params.split(",")]
        fn = getattr(L, name)
        if fn.argtypes is None:
            assert not plist or name in ("fp8b_scaled_mm_multicast", "fp8b_scaled_mm_peers"), f"{name}: no argtypes set"
            continue
        assert len(fn.argtypes) == len(plist), f"{name}: header has {len(plist)} parameters, binding {len(fn.argtypes)}"
        for decl, ct in zip(plist, fn.argtypes):
            is_ptr = "*" in decl
            assert is_ptr == (ct is ctypes.c_void_p or ct is ctypes.c_char_p), f"{name}: '{decl}' bound as {ct}"
            if not is_ptr:
                want = ctypes.c_size_t if "size_t" in decl else ctypes.c_int64 if "int64_t" in decl else ctypes.c_int
                assert ct is want, f"{name}: '{decl}' bound as {ct}"
        checked += 1
    assert checked >= 18
