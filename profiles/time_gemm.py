#!/usr/bin/env python3
"""Time one scaled-matmul shape with CUDA events over a 4-set (A,B,C) rotation -- quick A/B harness
for kernel variants (FP8B_GEMM_CFG / FP8B_GEMM_DEBUG / FP8B_GEMV_IMPL knobs).

    python profiles/time_gemm.py [M K N [algo]]
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "fp8-mps-metal_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from _util import capi
L = capi(); dev = torch.device("cuda", 0)
L.fp8b_set_option(0, int(os.environ.get("PDL", "1"))); L.fp8b_set_option(1, int(os.environ.get("STATICW", "0")))
M, K, N = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (4096, 3072, 12288)))
algo = int(sys.argv[4]) if len(sys.argv) > 4 else 2
g = torch.Generator(device=dev).manual_seed(0)
P = lambda t: ctypes.c_void_p(t.data_ptr())
sets = []
for _ in range(4):
    A = torch.randint(0, 120, (M, K), dtype=torch.uint8, device=dev, generator=g)
    B = torch.randint(0, 120, (N, K), dtype=torch.uint8, device=dev, generator=g)
    C = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    sets.append((A, B, C))
one = torch.full((1,), 0.01, device=dev)
def run():
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)      # the capture stream inside torch.cuda.graph
    for A, B, C in sets:
        rc = L.fp8b_scaled_mm(P(A), P(B), P(C), 2, M, N, K, N, P(one), 1, P(one), 1, None, 0, None, None, 0, algo, st)
        assert rc == 0, rc
for _ in range(3): run()
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    run()
for _ in range(3): gr.replay()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): gr.replay()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 40
env = {k: v for k, v in os.environ.items() if k.startswith("FP8B_") or k in ("PDL", "STATICW")}
print(f"{env} M{M} K{K} N{N} algo{algo}: {us:.2f} us  {2.0*M*N*K/us/1e6:.0f} TFLOP/s  {(M*K+N*K)/us/1e3:.0f} GB/s(in)")
