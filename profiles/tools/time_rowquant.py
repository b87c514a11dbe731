"""Per-row dynamic quantisation (fp8_quantize_rowwise) and the M > 16 dynamic linear on C4's shape.
Usage: python profiles/tools/time_rowquant.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "fp8-mps-metal_b200"))
import torch
import fp8_mps_native as nat

dev = torch.device("cuda:0")

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps

for (M, K) in ((4096, 3072), (4096, 12288), (16, 14336), (128, 4096), (65536, 1024)):
    xs = [torch.randn(M, K, device=dev, dtype=torch.bfloat16) for _ in range(max(1, min(8, (1 << 30) // (M * K * 2))))]
    us = timeit(lambda: [nat.fp8_quantize_rowwise(x) for x in xs]) / len(xs)
    print(f"quantize_rowwise bf16 ({M:6d},{K:6d}): {us:8.2f} us  {3.0 * M * K / us / 1e3:7.0f} GB/s (3 B/elem)", flush=True)

M, K, N = 4096, 3072, 12288
x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
W = torch.randint(0, 120, (N, K), dtype=torch.uint8, device=dev)
sb = torch.tensor([0.01], device=dev)
q, inv = nat.fp8_quantize_rowwise(x)
t_mm = timeit(lambda: nat.fp8_scaled_mm_fused(q, W, inv, sb, None, None, torch.bfloat16))
t_dyn = timeit(lambda: nat.fp8_linear_dynamic(x, W, sb, None, torch.bfloat16))
print(f"C4 shape: pre-quantised GEMM {t_mm:.1f} us, dynamic linear (quantise + GEMM) {t_dyn:.1f} us")
