"""
fp8-mps-metal (B200 build) -- ComfyUI custom-node entry point.

Same contract as the reference's __init__.py (:22-61): importing the folder installs the
FP8 patches and exports empty node mappings.  Import is side-effect free outside ComfyUI
unless FP8_B200_AUTOINSTALL=1 is set, so tests and benchmarks control install() themselves.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

NODE_CLASS_MAPPINGS = {}
NODE_DISPLAY_NAME_MAPPINGS = {}
__all__ = ["NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS"]

if "comfy" in sys.modules or os.environ.get("FP8_B200_AUTOINSTALL") == "1":
    try:
        import fp8_mps_patch

        if not fp8_mps_patch.is_installed():
            fp8_mps_patch.install()
            print("[fp8-mps-metal] FP8 patches installed (B200 / sm_100a kernels)")
    except Exception as e:  # mirror the reference: never break the host app on load (:43-53)
        print(f"[fp8-mps-metal] failed to install FP8 patches: {e}")
