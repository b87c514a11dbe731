#!/usr/bin/env python3
"""One GPU: C4 (or M K N from argv) with the st.global epilogue vs the TMA-store epilogue (fp8b_scaled_mm_push with
1, 2 destinations in local HBM), CUDA events over a 4-set rotation in one CUDA graph.

    python profiles/tools/time_push.py [M K N]
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "fp8-mps-metal_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from _util import capi
L = capi(); dev = torch.device("cuda", 0)
M, K, N = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (4096, 3072, 12288)))
P = lambda t: ctypes.c_void_p(t.data_ptr())
import fp8_mps_native
g = torch.Generator(device=dev).manual_seed(0)
sets = []
for _ in range(4):
    A, ia = fp8_mps_native.fp8_quantize(torch.randn(M, K, device=dev, generator=g))
    B, ib = fp8_mps_native.fp8_quantize(torch.randn(N, K, device=dev, generator=g))
    C = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    C2 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    sets.append((A, B, C, C2, ia, ib))


def direct():
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for A, B, C, C2, ia, ib in sets:
        rc = L.fp8b_scaled_mm(P(A), P(B), P(C), 2, M, N, K, N, P(ia), 1, P(ib), 1, None, 0, None, None, 0, 2, st)
        assert rc == 0, rc


def push(nd):
    def run():
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for A, B, C, C2, ia, ib in sets:
            arr = (ctypes.c_void_p * nd)(*([C.data_ptr(), C2.data_ptr()][:nd]))
            rc = L.fp8b_scaled_mm_push(P(A), P(B), arr, nd, 2, M, N, K, N, P(ia), 1, P(ib), 1, None, 0, None, st)
            assert rc == 0, rc
    return run


def timeit(fn, name):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    for _ in range(3): gr.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): gr.replay()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / 40)
    print(f"{name:28s} M{M} K{K} N{N}: {best:8.2f} us  {2.0*M*N*K/best/1e6:7.0f} TFLOP/s", flush=True)


for raster in (1, 2):
    L.fp8b_set_option(24, raster)
    for cfg in ([0, 3, 5, 4] if len(sys.argv) <= 4 else [0]):
        L.fp8b_set_option(16, cfg if cfg else -1)
        L.fp8b_set_option(20, 1)
        timeit(direct, f"raster{raster} cfg{cfg} st.global epilogue")
        L.fp8b_set_option(20, -1)
        timeit(push(1), f"raster{raster} cfg{cfg} TMA store x1")
        timeit(push(2), f"raster{raster} cfg{cfg} TMA store x2 (local)")
L.fp8b_set_option(16, -1); L.fp8b_set_option(24, -1)
torch.cuda.synchronize()
sets[0][2].zero_(); sets[0][3].zero_()
direct(); ref = sets[0][2].clone(); sets[0][2].zero_()
push(2)(); torch.cuda.synchronize()
print("push == direct:", bool(torch.equal(ref, sets[0][2])), bool(torch.equal(ref, sets[0][3])))
