import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "fp8-mps-metal_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from _util import mm_capi
A = torch.randint(0, 120, (2, 37), dtype=torch.uint8, device="cuda")
B = torch.randint(0, 120, (5, 37), dtype=torch.uint8, device="cuda")
rc, C = mm_capi(A, B, torch.ones(1), torch.ones(1))
torch.cuda.synchronize()
print(rc, C)
