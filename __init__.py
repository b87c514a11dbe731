"""
ComfyUI custom node: FP8 (float8_e4m3fn) kernels for NVIDIA B200 behind the fp8-mps-metal API.

Installation, as for the reference (its __init__.py:1-11): clone this repository into
ComfyUI/custom_nodes/ and build the native library once (`python fp8-mps-metal_b200/build.py`).
When ComfyUI imports this folder the patches are installed: torch._scaled_mm, Tensor.to and
Tensor.copy_ route FP8 work on CUDA devices to the sm_100a kernels in fp8-mps-metal_b200/.

This file mirrors the reference's root module (__init__.py:13-61): put the implementation directory
on sys.path, install the patch unless it already is, report the outcome without ever breaking the
host application, and export empty node mappings (the node adds no graph nodes).
"""

import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(_ROOT, "fp8-mps-metal_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

try:
    import fp8_mps_patch

    if not fp8_mps_patch.is_installed():
        fp8_mps_patch.install()
        print("\n" + "=" * 70)
        print("FP8 B200 patch installed")
        print("=" * 70)
        print("float8_e4m3fn operations on CUDA devices now run on the sm_100a kernels:")
        print("  - FP8 weight loading and Tensor.to(float8_e4m3fn) with the reference codec")
        print("  - FP8 stochastic rounding (.copy_() into FP8 tensors)")
        print("  - torch._scaled_mm: split-K GEMV (M <= 16), tcgen05 GEMM (large M)")
        print("=" * 70 + "\n")
    else:
        print("[fp8-mps-metal] Patch already installed")
except Exception as e:  # never take the host application down on load (reference __init__.py:43-53)
    print("\n" + "!" * 70)
    print("WARNING: failed to install the FP8 B200 patch")
    print("!" * 70)
    print(f"Error: {e}")
    print("FP8 operations will use PyTorch's stock CUDA paths.")
    print("Build the native library with: python fp8-mps-metal_b200/build.py")
    print("!" * 70 + "\n")
    import traceback
    traceback.print_exc()

# ComfyUI needs these to recognise the folder as a custom node; no graph nodes are added.
NODE_CLASS_MAPPINGS = {}
NODE_DISPLAY_NAME_MAPPINGS = {}

__all__ = ["NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS"]
