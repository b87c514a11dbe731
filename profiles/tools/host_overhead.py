"""Host-side cost per call of each API layer (no device sync inside the loop): how many microseconds of CPU time a
decode-path GEMV costs before the kernel is even enqueued.  Usage: python profiles/tools/host_overhead.py"""
import ctypes, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "fp8-mps-metal_b200"))
import torch
import fp8_mps_native as nat, fp8_mps_patch
from _util import capi

dev = torch.device("cuda:0")
L = capi()
K, N = 4096, 4096
x = torch.randint(0, 120, (1, K), dtype=torch.uint8, device=dev)
W = torch.randint(0, 120, (N, K), dtype=torch.uint8, device=dev)
sa = torch.tensor([0.01], device=dev); sb = torch.tensor([0.02], device=dev)
out = torch.empty(1, N, dtype=torch.bfloat16, device=dev)
x8, W8 = x.view(torch.float8_e4m3fn), W.view(torch.float8_e4m3fn)
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def bench(name, fn, n=3000):
    for _ in range(50): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:46s} host {1e6 * (t1 - t0) / n:7.2f} us/call   (with drain {1e6 * (t2 - t0) / n:7.2f})", flush=True)

bench("C ABI via ctypes (fp8b_scaled_mm)", lambda: L.fp8b_scaled_mm(P(x), P(W), P(out), 2, 1, N, K, N, P(sa), 1, P(sb), 1, None, 0, None, None, 0, 0, st))
bench("bridge fp8_metal.fp8_scaled_mm_fused", lambda: nat._get_lib().fp8_scaled_mm_fused(x, W, sa, sb, None, None, torch.bfloat16, 0, None, 0, 0))
bench("fp8_mps_native.fp8_scaled_mm_fused", lambda: nat.fp8_scaled_mm_fused(x, W, sa, sb, None, None, torch.bfloat16))
bench("fp8_mps_native.fp8_scaled_mm (fp32 out)", lambda: nat.fp8_scaled_mm(x, W, sa, sb))
fp8_mps_patch.install()
bench("patched torch._scaled_mm (positional)", lambda: torch._scaled_mm(x8, W8.t(), sa, sb, None, None, torch.bfloat16))
bench("patched torch._scaled_mm (keywords)", lambda: torch._scaled_mm(x8, W8.t(), scale_a=sa, scale_b=sb, out_dtype=torch.bfloat16))
xb = torch.randn(1 << 16, device=dev, dtype=torch.bfloat16)
bench("patched Tensor.to(float8_e4m3fn), 64K elems", lambda: xb.to(torch.float8_e4m3fn))
q = xb.to(torch.float8_e4m3fn)
bench("patched Tensor.to(float16) from fp8", lambda: q.to(torch.float16))
bench("patched Tensor.to(float32) on a float tensor", lambda: xb.to(torch.float32))
fp8_mps_patch.uninstall()
bench("unpatched Tensor.to(float32) on a float tensor", lambda: xb.to(torch.float32))
