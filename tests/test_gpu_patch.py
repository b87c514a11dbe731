"""GPU: the drop-in boundary -- fp8_mps_patch.install() routes torch._scaled_mm / Tensor.to /
Tensor.copy_ through the B200 kernels (mirrors test_fp8_metal.py:318-705, validate_fix.py:50-160)."""
import numpy as np
import pytest
import torch

import fp8_mps_patch
import fp8_oracle as o
from _util import capi, to_np

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _installed():
    fp8_mps_patch.install()
    yield
    fp8_mps_patch.uninstall()


def _launches():
    return capi().fp8b_launch_count()


def test_to_fp8_uses_reference_codec_not_torch(golden):
    x32 = torch.from_numpy(golden["codec"]["enc_f32_in"])
    ref = golden["codec"]["enc_f32_out"]
    ok = ref != 0xFF
    n0 = _launches()
    q = x32.to(DEV).to(torch.float8_e4m3fn)
    assert _launches() > n0                                     # our kernel ran, not torch's cast
    assert q.dtype == torch.float8_e4m3fn and q.device.type == "cuda"
    assert np.array_equal(q.view(torch.uint8).cpu().numpy()[ok], ref[ok])
    q2 = x32.to(DEV, torch.float8_e4m3fn)                       # CPU -> device + dtype in one call
    assert np.array_equal(q2.view(torch.uint8).cpu().numpy()[ok], ref[ok])
    q3 = x32.to(device=DEV).to(dtype=torch.float8_e4m3fn)
    assert torch.equal(q3.view(torch.uint8), q.view(torch.uint8))
    for dt in (torch.float16, torch.bfloat16):
        xb = x32.to(dt)
        qb = xb.to(DEV).to(torch.float8_e4m3fn)
        assert np.array_equal(qb.view(torch.uint8).cpu().numpy(), o.encode(xb.float().numpy()))


def test_fp8_to_float_and_noops():
    b = torch.arange(256, dtype=torch.int32).to(torch.uint8)
    q = b.view(torch.float8_e4m3fn).to(DEV)                     # scenario 1: byte-preserving transfer
    assert q.dtype == torch.float8_e4m3fn and torch.equal(q.view(torch.uint8).cpu(), b)
    assert q.to(torch.float8_e4m3fn) is q                       # scenario 3 no-op (fp8_mps_patch.py:201-203)
    assert q.to(DEV) is q
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        n0 = _launches()
        f = q.to(dt)
        assert _launches() > n0 and f.dtype == dt
        assert np.array_equal(to_np(f).view(np.uint32), o.DECODE_TABLE.view(np.uint32))
    e5 = q.to(torch.float8_e5m2)                                # FP8 -> FP8 reinterpretation
    assert e5.dtype == torch.float8_e5m2 and torch.equal(e5.view(torch.uint8), q.view(torch.uint8))
    assert torch.empty(0, device=DEV).to(torch.float8_e4m3fn).numel() == 0
    back = q.to("cpu")                                          # device -> CPU stays with torch
    assert back.device.type == "cpu" and torch.equal(back.view(torch.uint8), b)


def test_value_preservation_no_autoscale():
    """validate_fix.py:50-160 / test_fp8_metal.py:582-705: .to(fp8) does not rescale; rel err < 15 %."""
    vals = torch.tensor([1.0, 2.0, 5.0, 10.0, 50.0, 100.0, 3.0, 440.0], device=DEV)
    back = vals.to(torch.float8_e4m3fn).to(torch.float32)
    assert ((back - vals).abs() / vals).max().item() < 0.15
    assert abs(back[6].item() - 3.0) < 0.26


def test_copy_scenarios():
    x = torch.tensor([0.5, 1.0, 2.0, 10.0, 100.0, -0.001, 0.0186, 500.0])
    want = o.encode(x.numpy())
    dst = torch.empty(8, dtype=torch.float8_e4m3fn, device=DEV)
    assert dst.copy_(x.to(DEV)) is dst                          # float (device) -> FP8
    assert np.array_equal(dst.view(torch.uint8).cpu().numpy(), want)
    dst2 = torch.empty(8, dtype=torch.float8_e4m3fn, device=DEV)
    dst2.copy_(x)                                               # float (CPU) -> FP8 on device
    assert np.array_equal(dst2.view(torch.uint8).cpu().numpy(), want)
    dst3 = torch.empty(8, dtype=torch.float8_e4m3fn, device=DEV)
    dst3.copy_(dst)                                             # FP8 -> FP8 byte copy (stochastic-rounding path)
    assert torch.equal(dst3.view(torch.uint8), dst.view(torch.uint8))
    dst4 = torch.empty(2, 8, dtype=torch.float8_e4m3fn, device=DEV)
    dst4.copy_(x.to(DEV))                                       # broadcasting copy
    assert np.array_equal(dst4.view(torch.uint8).cpu().numpy(), np.stack([want, want]))
    d5 = torch.empty(8, dtype=torch.float8_e5m2, device=DEV)
    d5.copy_(dst)                                               # e4m3fn -> e5m2 destination: BYTE copy (fp8_mps_patch.py:245-264)
    assert torch.equal(d5.view(torch.uint8), dst.view(torch.uint8))
    xi = torch.tensor([0, 1, -3, 7, 100, 449, -1000, 2], dtype=torch.int32)
    d6 = torch.empty(8, dtype=torch.float8_e4m3fn, device=DEV)
    d6.copy_(xi.to(DEV))                                        # non-float source: float32 then the codec (fp8_mps_patch.py:266-290)
    assert np.array_equal(d6.view(torch.uint8).cpu().numpy(), o.encode(xi.numpy().astype(np.float32)))
    d7 = torch.empty(8, dtype=torch.float8_e4m3fn, device=DEV)
    d7.copy_(x.double())                                        # float64 from the CPU
    assert np.array_equal(d7.view(torch.uint8).cpu().numpy(), want)
    f = torch.zeros(8, device=DEV)
    f.copy_(x)                                                  # unrelated copies untouched
    assert torch.equal(f.cpu(), x)
    f.copy_(3)
    assert float(f[0]) == 3.0


@pytest.mark.parametrize("M,K,N,odt", [(1, 4096, 4096, torch.float16), (4, 4096, 4096, torch.bfloat16),
                                       (300, 512, 640, torch.bfloat16), (64, 256, 128, None)])
def test_scaled_mm_drop_in(M, K, N, odt):
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g)
    qa, sa = o.fp8_quantize(a.numpy())
    qw, sw = o.fp8_quantize(w.numpy())
    bias = torch.randn(N, generator=g).to(odt or torch.float32)
    x8 = torch.from_numpy(qa).to(DEV).view(torch.float8_e4m3fn)
    w8 = torch.from_numpy(qw).to(DEV).view(torch.float8_e4m3fn)       # (N,K) row-major
    tsa, tsw = torch.from_numpy(sa).to(DEV), torch.from_numpy(sw).to(DEV)
    n0 = _launches()
    y_kw = torch._scaled_mm(x8, w8.t(), scale_a=tsa, scale_b=tsw, bias=bias.to(DEV), out_dtype=odt)
    y_pos = torch._scaled_mm(x8, w8.t(), tsa, tsw, bias.to(DEV), None, odt)             # north-star call form
    assert _launches() >= n0 + 2
    assert y_kw.dtype == (odt or torch.float32) and y_kw.shape == (M, N)                # out_dtype=None -> fp32
    assert torch.equal(y_kw, y_pos)
    name = {torch.float16: "f16", torch.bfloat16: "bf16", None: "f32"}[odt]
    ref = o.scaled_mm(qa, qw, sa, sw, bias=to_np(bias), out_dtype=name)
    tol = {"f32": 1e-4, "f16": 5e-4, "bf16": 3e-3}[name]
    assert o.rel_rmse(to_np(y_kw), ref) <= tol
    # uint8 operands and a row-major `other` (needs the .contiguous() transpose, fp8_mps_patch.py:84)
    y_u8 = torch._scaled_mm(x8.view(torch.uint8), w8.view(torch.uint8).t().contiguous(), scale_a=tsa, scale_b=tsw,
                            bias=bias.to(DEV), out_dtype=odt)
    assert torch.equal(y_u8, y_kw)
    # default scales (fp8_mps_patch.py:87-90)
    y_def = torch._scaled_mm(x8, w8.t(), out_dtype=torch.float32)
    assert o.rel_rmse(to_np(y_def), o.scaled_mm(qa, qw, np.ones(1, np.float32), np.ones(1, np.float32))) <= 1e-4


def test_non_fp8_calls_reach_the_original_op():
    a = torch.randn(32, 64, device=DEV).to(torch.bfloat16)
    assert a.to(torch.float32).dtype == torch.float32
    n0 = _launches()
    _ = a.to(torch.float16)
    assert _launches() == n0
    fp8_mps_patch.uninstall()
    assert torch.Tensor.to is not fp8_mps_patch._metal_tensor_to
    fp8_mps_patch.install()


def test_to_with_device_ordinal_uses_reference_codec(golden):
    """`.to(0, float8_e4m3fn)`: a bare device ordinal must reach the codec kernel, not torch's cast (which gives NaN
    codes above 448 where the reference saturates)."""
    x = torch.tensor([500.0, -1000.0, 1.0, 0.001])
    n0 = _launches()
    q = x.to(0, torch.float8_e4m3fn)
    assert _launches() > n0 and q.device == torch.device("cuda", 0)
    assert np.array_equal(q.view(torch.uint8).cpu().numpy(), o.encode(x.numpy()))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_to_other_gpu_lands_on_the_requested_device():
    """`x_cuda0.to('cuda:1', float8_e4m3fn)` must return a tensor on cuda:1."""
    x = torch.tensor([500.0, -2.5, 1.0, 0.0186], device="cuda:0")
    for target in ("cuda:1", torch.device("cuda", 1), 1):
        q = x.to(target, torch.float8_e4m3fn)
        assert q.device == torch.device("cuda", 1)
        assert np.array_equal(q.view(torch.uint8).cpu().numpy(), o.encode(x.cpu().numpy()))
    q0 = x.to("cuda:0", torch.float8_e4m3fn)
    assert q0.device == torch.device("cuda", 0)
    moved = q0.to("cuda:1")                                     # FP8 bytes between GPUs
    assert moved.device == torch.device("cuda", 1) and torch.equal(moved.view(torch.uint8).cpu(), q0.view(torch.uint8).cpu())
