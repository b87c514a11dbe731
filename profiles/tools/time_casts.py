"""Runs only bench.py's C5 leg (FLUX-sized cast sweeps: per-tensor launches, 4 streams, batched API).
Usage: python profiles/tools/time_casts.py"""
import json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import torch
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
gen = torch.Generator(device=dev).manual_seed(0)
print(json.dumps(bench.bench_casts_c5(torch, bench._capi(), gen, dev, bench.load_peaks(), 5, 3), indent=1))
