#!/usr/bin/env python3
"""Where does the persistent TMA-ring GEMV lose bandwidth?  Needs the profiling library:
    FP8B_LIB=profiles/tools/bin/libfp8_b200_profile.so python profiles/tools/ring_probe.py
FP8B_RING_DEBUG (read per call by the profiling build): 0 = full kernel, 1 = consumers only wait/arrive (the TMA
ring alone), 2 = + LDS of the weights, 3 = + F2FP decode (no MMA).  Results of modes 1-3 are garbage by design."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "fp8-mps-metal_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from _util import capi
L = capi(); dev = torch.device("cuda", 0)
P = lambda t: ctypes.c_void_p(t.data_ptr())
g = torch.Generator(device=dev).manual_seed(0)
for name, M, K, N, rot in (("C2", 1, 14336, 4096, 16), ("SQ", 1, 14336, 14336, 4), ("C1", 1, 4096, 4096, 32), ("C3", 4, 4096, 4096, 32)):
    x = torch.randint(0, 120, (M, K), dtype=torch.uint8, device=dev, generator=g)
    Ws = [torch.randint(0, 120, (N, K), dtype=torch.uint8, device=dev, generator=g) for _ in range(rot)]
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    one = torch.full((1,), 0.01, device=dev)
    nbytes = N * K
    def run():
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for W in Ws:
            rc = L.fp8b_scaled_mm(P(x), P(W), P(out), 2, M, N, K, N, P(one), 1, P(one), 1, None, 0, None, None, 0, 1, st)
            assert rc == 0, rc
    res = []
    for impl, dbg, cps in ((1, 0, 1), (4, 0, 1), (4, 1, 1), (4, 2, 1), (4, 3, 1), (4, 1, 2), (4, 0, 2)):
        if impl == 1 and M > 1: impl = 2
        L.fp8b_set_option(17, impl); L.fp8b_set_option(21, cps)
        os.environ["FP8B_RING_DEBUG"] = str(dbg)
        for _ in range(2): run()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            run()
        for _ in range(3): gr.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): gr.replay()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / (10 * rot))
        res.append(f"impl{impl} dbg{dbg} x{cps}/SM {best:6.2f} us {nbytes / best / 1e3:5.0f} GB/s")
    os.environ["FP8B_RING_DEBUG"] = "0"
    L.fp8b_set_option(17, -1); L.fp8b_set_option(21, -1)
    print(f"{name} M{M} K{K} N{N}:\n   " + "\n   ".join(res), flush=True)
    del Ws
    torch.cuda.empty_cache()
