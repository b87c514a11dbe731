"""Times the decode-path linear with dynamic activation quantisation on C2's shape:
   (a) two-step: fp8_quantize_rowwise(x) then fp8_scaled_mm_fused  (what the reference composes, minus its host sync)
   (b) fused:    fp8_linear_dynamic(x, W)  -- one launch
Each is captured in a CUDA graph over a 16-weight rotation (inputs > L2).  Usage: python profiles/tools/time_dynamic.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "fp8-mps-metal_b200"))
import torch
import fp8_mps_native as nat

dev = torch.device("cuda:0")
K, N, R = 14336, 4096, 16
for M in (1, 4):
    Ws = [torch.randint(0, 256, (N, K), dtype=torch.uint8, device=dev) for _ in range(R)]
    for W in Ws:
        W[(W & 0x7F) == 0x7F] = 0x3C
    x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    sb = torch.tensor([0.01], device=dev)

    def two_step():
        for W in Ws:
            q, inv = nat.fp8_quantize_rowwise(x)
            nat.fp8_scaled_mm_fused(q, W, inv, sb, None, None, torch.bfloat16)

    def fused():
        for W in Ws:
            nat.fp8_linear_dynamic(x, W, sb, None, torch.bfloat16)

    def fused1():
        for W in Ws:
            nat.fp8_linear_dynamic(x, W, sb, None, torch.bfloat16, single_kernel=True)

    def prequant():
        q, inv = nat.fp8_quantize_rowwise(x)
        for W in Ws:
            nat.fp8_scaled_mm_fused(q, W, inv, sb, None, None, torch.bfloat16)

    for name, fn in (("prequantised (floor)", prequant), ("two_step", two_step), ("chained (PDL)", fused), ("single kernel", fused1)):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(8):
                    fn()
            g.replay(); s.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(5):
                g.replay()
            e1.record(s); s.synchronize()
        print(f"M={M} {name:22s} {e0.elapsed_time(e1) * 1e3 / (5 * 8 * R):8.2f} us per linear")
