"""CPU: the interception layer's host logic (fp8_mps_patch), mirroring the reference's
install/uninstall and symbol tests (test_fp8_metal.py:318-349, test_mps_limits_patch.py:21-70,128-153)."""
import os

import pytest
import torch

import fp8_mps_patch


@pytest.fixture(autouse=True)
def _clean():
    if fp8_mps_patch.is_installed():
        fp8_mps_patch.uninstall()
    yield
    if fp8_mps_patch.is_installed():
        fp8_mps_patch.uninstall()


def test_required_symbols_exist():
    for name in ["install", "uninstall", "is_installed", "patch_vae_decode_for_mps_limits",
                 "_metal_scaled_mm", "_metal_tensor_to", "_metal_tensor_copy",
                 "MPS_TENSOR_SIZE_THRESHOLD", "VAE_UPSCALE_FACTOR"]:
        assert hasattr(fp8_mps_patch, name), name


def test_install_uninstall_idempotent_and_restores():
    orig_mm, orig_to, orig_copy = torch._scaled_mm, torch.Tensor.to, torch.Tensor.copy_
    assert not fp8_mps_patch.is_installed()
    fp8_mps_patch.install()
    assert fp8_mps_patch.is_installed()
    assert torch._scaled_mm is fp8_mps_patch._metal_scaled_mm
    assert torch.Tensor.to is fp8_mps_patch._metal_tensor_to
    assert torch.Tensor.copy_ is fp8_mps_patch._metal_tensor_copy
    assert fp8_mps_patch._original_scaled_mm is orig_mm
    assert fp8_mps_patch._original_tensor_to is orig_to
    assert fp8_mps_patch._original_tensor_copy is orig_copy
    fp8_mps_patch.install()                                   # second install is a no-op
    assert fp8_mps_patch._original_scaled_mm is orig_mm
    fp8_mps_patch.uninstall()
    assert not fp8_mps_patch.is_installed()
    assert torch._scaled_mm is orig_mm and torch.Tensor.to is orig_to and torch.Tensor.copy_ is orig_copy
    assert fp8_mps_patch._original_scaled_mm is None
    fp8_mps_patch.uninstall()                                 # second uninstall is a no-op


def test_install_sets_fallback_env_var():
    os.environ.pop("PYTORCH_ENABLE_MPS_FALLBACK", None)
    fp8_mps_patch.install()
    assert os.environ.get("PYTORCH_ENABLE_MPS_FALLBACK") == "1"


def test_cpu_tensors_pass_through_unchanged():
    """Off the accelerator every wrapper forwards to the original op (fp8_mps_patch.py:64-72)."""
    x = torch.tensor([0.5, 1.0, 2.0, 10.0, 100.0, -3.0])
    ref_bytes = x.to(torch.float8_e4m3fn).view(torch.uint8)
    fp8_mps_patch.install()
    q = x.to(torch.float8_e4m3fn)
    assert q.dtype == torch.float8_e4m3fn and torch.equal(q.view(torch.uint8), ref_bytes)
    assert q.to(torch.float8_e4m3fn) is q or torch.equal(q.to(torch.float8_e4m3fn).view(torch.uint8), ref_bytes)
    back = q.to(torch.float32)
    assert back.dtype == torch.float32 and back[1] == 1.0
    dst = torch.empty(6, dtype=torch.float8_e4m3fn)
    assert dst.copy_(x) is dst
    assert torch.equal(dst.view(torch.uint8), ref_bytes)
    d2 = torch.empty(6, dtype=torch.float8_e4m3fn)
    d2.copy_(q)
    assert torch.equal(d2.view(torch.uint8), ref_bytes)
    t = torch.zeros(3)
    t.copy_(torch.ones(3))
    assert t.sum() == 3
    assert torch.ones(2).to("cpu", torch.float64).dtype == torch.float64
    assert torch.ones(2).to(dtype=torch.int32).dtype == torch.int32
    e = torch.empty(0).to(torch.float8_e4m3fn)               # empty tensor (test_fp8_metal.py:352-579)
    assert e.numel() == 0 and e.dtype == torch.float8_e4m3fn


def test_to_argument_forms_parse():
    parse = fp8_mps_patch._parse_to_args
    assert parse((torch.float16,), {}) == (torch.float16, None)
    assert parse(("cuda",), {}) == (None, "cuda")
    assert parse(("cuda", torch.float8_e4m3fn), {}) == (torch.float8_e4m3fn, "cuda")
    assert parse((), {"dtype": torch.bfloat16, "device": torch.device("cpu")}) == (torch.bfloat16, torch.device("cpu"))
    o = torch.zeros(1, dtype=torch.float64)
    assert parse((o,), {}) == (torch.float64, o.device)
    assert parse((1, torch.float8_e4m3fn), {}) == (torch.float8_e4m3fn, 1)          # to(ordinal, dtype)
    assert parse((torch.float16, True), {}) == (torch.float16, None)                # to(dtype, non_blocking)
    assert parse((True,), {}) == (None, None)
    assert fp8_mps_patch._device_type(2) == "cuda" and fp8_mps_patch._device_type(True) is None
    assert fp8_mps_patch._as_device(3) == torch.device("cuda", 3)
    assert fp8_mps_patch._as_device("cuda:1") == torch.device("cuda", 1)
    assert fp8_mps_patch._device_type("cuda:3") == "cuda"
    assert fp8_mps_patch._device_type(torch.device("cpu")) == "cpu"
    assert fp8_mps_patch._is_fp8_dtype(torch.float8_e4m3fn) and fp8_mps_patch._is_fp8_dtype(torch.float8_e5m2)
    assert not fp8_mps_patch._is_fp8_dtype(torch.uint8) and not fp8_mps_patch._is_fp8_dtype(None)


def test_scaled_mm_wrapper_accepts_positional_and_keyword_scales():
    """The reference wrapper is keyword-only (fp8_mps_patch.py:53); the aten schema is positional."""
    calls = []

    def fake(input, other, **kw):
        calls.append(kw)
        return "orig"

    fp8_mps_patch.install()
    saved = fp8_mps_patch._original_scaled_mm
    fp8_mps_patch._original_scaled_mm = fake
    try:
        a = torch.zeros(2, 16)
        b = torch.zeros(16, 3)
        s = torch.ones(1)
        assert torch._scaled_mm(a, b, s, s, None, None, torch.bfloat16) == "orig"
        assert calls[-1]["scale_a"] is s and calls[-1]["out_dtype"] == torch.bfloat16
        assert torch._scaled_mm(a, b, scale_a=s, scale_b=s, out_dtype=torch.float16) == "orig"
        assert calls[-1]["scale_b"] is s and calls[-1]["out_dtype"] == torch.float16
        with pytest.raises(TypeError):
            torch._scaled_mm(a, b, s, s, None, None, None, False, 1)
    finally:
        fp8_mps_patch._original_scaled_mm = saved


def test_vae_stub_is_callable():
    fp8_mps_patch.patch_vae_decode_for_mps_limits()


def test_repo_root_loads_as_a_comfyui_custom_node(tmp_path, capsys):
    """ComfyUI imports custom_nodes/<clone>/__init__.py by path (reference: __init__.py:22-61).  The repo root must
    install the patches and export the (empty) node mappings, exactly like the reference's root module."""
    import importlib.util
    import sys
    import types
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fp8_mps_patch.uninstall()
    fake = types.ModuleType("comfy")
    had = sys.modules.get("comfy")
    sys.modules["comfy"] = fake
    try:
        spec = importlib.util.spec_from_file_location("fp8_b200_custom_node", os.path.join(root, "__init__.py"),
                                                      submodule_search_locations=[root])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert mod.NODE_CLASS_MAPPINGS == {} and mod.NODE_DISPLAY_NAME_MAPPINGS == {}
        assert mod.__all__ == ["NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS"]
        assert fp8_mps_patch.is_installed()
        assert torch._scaled_mm is fp8_mps_patch._metal_scaled_mm
        out = capsys.readouterr().out
        assert "installed" in out.lower()
        spec.loader.exec_module(mod)                           # second load: idempotent (reference __init__.py:39-40)
        assert "already installed" in capsys.readouterr().out
    finally:
        fp8_mps_patch.uninstall()
        if had is None:
            sys.modules.pop("comfy", None)
        else:
            sys.modules["comfy"] = had


def test_pyproject_has_comfy_registry_fields():
    import tomllib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "pyproject.toml"), "rb") as f:
        meta = tomllib.load(f)
    comfy = meta["tool"]["comfy"]                                # reference pyproject.toml:12-15
    assert comfy["PublisherId"] and comfy["DisplayName"]
    assert meta["project"]["name"] and meta["project"]["version"]
