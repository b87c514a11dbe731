"""Per-step latency of copy-in -> patched torch._scaled_mm -> copy-out -> sync, with and without PDL launches."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "fp8-mps-metal_b200"))
import torch
import fp8_mps_native as nat, fp8_mps_patch
dev = torch.device("cuda:0")
K, N = 14336, 4096
W = torch.randint(0, 120, (N, K), dtype=torch.uint8, device=dev).view(torch.float8_e4m3fn)
hx = torch.randint(0, 120, (1, K), dtype=torch.uint8).pin_memory()
hout = torch.empty(1, N, dtype=torch.bfloat16).pin_memory()
xd = torch.empty(1, K, dtype=torch.uint8, device=dev)
sa = torch.tensor([0.01], device=dev); sb = torch.tensor([0.02], device=dev)
fp8_mps_patch.install()
lib = nat._get_lib()
for pdl in (1, 0, 1, 0):
    lib.set_option(lib.OPT_PDL, pdl)
    def step():
        xd.copy_(hx, non_blocking=True)
        y = torch._scaled_mm(xd.view(torch.float8_e4m3fn), W.t(), sa, sb, None, None, torch.bfloat16)
        hout.copy_(y, non_blocking=True)
        torch.cuda.synchronize()
    for _ in range(20): step()
    t0 = time.perf_counter()
    for _ in range(300): step()
    print(f"PDL={pdl}: {(time.perf_counter() - t0) / 300 * 1e6:.1f} us per step", flush=True)
