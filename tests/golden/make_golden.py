#!/usr/bin/env python3
"""
Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

What is imported from the reference (audiohacking/fp8-mps-metal), unmodified:
  * test_fp8_correctness.fp8_e4m3fn_decode_spec / fp8_e4m3fn_encode_spec
    (test_fp8_correctness.py:22-106) -- the reference's own pure-Python statement
    of the shader codec (fp8_matmul.metal:19-92);
  * test_fp8_metal.fp8_e4m3fn_decode_reference (test_fp8_metal.py:35-50).
The Metal kernels cannot run on Linux, so host-level goldens (fp8_quantize,
fp8_dequantize) are produced by running the reference's host arithmetic
(fp8_mps_native.py:121-122, :174-189) with torch CPU ops and the spec encoder in
place of the kernel launch.

float8_e5m2 has no codec in the reference (it mis-routes e5m2 tensors into the e4m3fn kernels); its
decode table is pinned to PyTorch's own CPU cast, the comparison target the reference itself uses for casts
(test_mps_vs_cpu.py:308):  e5m2_golden.npz.

Outputs (committed):  codec_golden.npz, host_golden.npz, e5m2_golden.npz, kat.json
This script never imports oracle/ -- the fixtures are independent of it.
"""

import json
import math
import os
import sys

import numpy as np
import torch

REF = os.environ.get("FP8_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

import test_fp8_correctness as ref_spec          # noqa: E402
import test_fp8_metal as ref_metal_test          # noqa: E402

enc = ref_spec.fp8_e4m3fn_encode_spec
dec = ref_spec.fp8_e4m3fn_decode_spec

NAN_SENTINEL = 0xFF   # the encoder can never emit 0xFF; marks "undefined in the reference"


def enc_array(x32: np.ndarray) -> np.ndarray:
    out = np.empty(x32.shape, dtype=np.uint8)
    flat = x32.reshape(-1)
    o = out.reshape(-1)
    for i, v in enumerate(flat.tolist()):
        if math.isnan(v):
            o[i] = NAN_SENTINEL
        elif math.isinf(v):
            o[i] = enc(math.copysign(1e30, v))     # shader: v >= 448 -> saturate (metal:53)
        else:
            o[i] = enc(v)
    return out


def main():
    rng = np.random.default_rng(20260118)

    # ---- decode: all 256 patterns, from both reference decoders
    dec_spec = np.array([dec(b) for b in range(256)], dtype=np.float64)
    dec_metal_test = np.array([ref_metal_test.fp8_e4m3fn_decode_reference(b) for b in range(256)],
                              dtype=np.float64)
    assert np.array_equal(dec_spec, dec_metal_test)
    # keep the sign of zero: the shader returns -0.0 for 0x80 (metal:39); the python
    # spec does too (-value with value == 0.0)
    decode_table = dec_spec.astype(np.float32)

    # ---- encode: exhaustive bf16 and fp16 inputs
    bf16_bits = np.arange(65536, dtype=np.uint32)
    bf16_vals = (bf16_bits << 16).view(np.float32)
    enc_bf16_all = enc_array(bf16_vals)
    fp16_vals = np.arange(65536, dtype=np.uint16).view(np.float16).astype(np.float32)
    enc_fp16_all = enc_array(fp16_vals)

    # ---- encode: fp32 inputs -- mixed-scale random, every rounding boundary +-1ulp,
    #      and a strided sweep of the exponent range
    parts = []
    parts.append(rng.standard_normal(12000).astype(np.float32))
    parts.append((rng.standard_normal(12000) * 0.02).astype(np.float32))
    parts.append((rng.standard_normal(6000) * 100).astype(np.float32))
    parts.append((rng.standard_normal(4000) * np.exp2(rng.integers(-14, 10, 4000))).astype(np.float32))
    # boundaries: midpoints between consecutive representable magnitudes, +-1,2 ulp
    mags = sorted({abs(dec(b)) for b in range(256)})
    mids = []
    for a, b in zip(mags[:-1], mags[1:]):
        mids.append((a + b) / 2)
    pts = np.array(mags + mids + [448.0, 464.0, 480.0, 1.0 / 512, 1.0 / 1024, 1.0 / 64,
                                  7.5 / 512, 0.0], dtype=np.float32)
    near = []
    for d in (-2, -1, 0, 1, 2):
        near.append((pts.view(np.int32) + d).view(np.float32))
    near = np.concatenate(near)
    near = near[np.isfinite(near)]
    parts.append(near)
    parts.append(-near)
    # strided sweep of all fp32 bit patterns (positive and negative, finite + inf)
    sweep = np.arange(0, 0x7F800001, 0x1F3F7, dtype=np.uint32)
    parts.append(sweep.view(np.float32))
    parts.append((sweep | 0x80000000).view(np.float32))
    # the reference's own test inputs
    parts.append(np.array([0.5, 1.0, 2.0, 10.0, 100.0, -0.001, 0.0186, 3.0, 5.0, 50.0, 440.0,
                           0.001953125, 0.013671875, 0.015625, 448.0, 500.0, -0.0, 0.0,
                           np.inf, -np.inf, 1e-45, -1e-45, 1.1754944e-38], dtype=np.float32))
    enc_f32_in = np.concatenate(parts).astype(np.float32)
    enc_f32_out = enc_array(enc_f32_in)

    np.savez_compressed(os.path.join(HERE, "codec_golden.npz"),
                        decode_table=decode_table,
                        enc_bf16_all=enc_bf16_all,
                        enc_fp16_all=enc_fp16_all,
                        enc_f32_in=enc_f32_in,
                        enc_f32_out=enc_f32_out)

    # ---- host-level goldens: dequantize (fp16 scale multiply) and quantize
    u8_all = torch.arange(256, dtype=torch.int32).to(torch.uint8)
    h = torch.tensor(decode_table).to(torch.float16)              # fp8_to_half_kernel, metal:222
    deq_scales = np.array([1.0, 0.5, 0.01, 0.0123456, 3.0, 1.0 / 448.0, 100.0, 7e-5, 200.0],
                          dtype=np.float32)
    deq_out = []
    for s in deq_scales:
        scale_val = torch.tensor([float(s)], dtype=torch.float32).to(torch.float16)  # native.py:121
        deq_out.append((h * scale_val).view(torch.int16).numpy().view(np.uint16))    # native.py:122
    deq_out = np.stack(deq_out)

    q_inputs, q_bytes, q_inv = [], [], []
    cases = [
        torch.tensor([0.0, 1.0, -1.0, 0.5, -0.5, 100.0, -100.0, 448.0]),     # test_fp8_metal.py:175
        torch.tensor([1.0, 2.0, 5.0, 10.0, 50.0, 100.0]),                    # validate_fix.py
        torch.zeros(16),
        torch.tensor(rng.standard_normal(2048).astype(np.float32)),
        torch.tensor((rng.standard_normal(2048) * 0.02).astype(np.float32)),
        torch.tensor((rng.standard_normal(1000) * 37.0).astype(np.float32)),
    ]
    for inp in cases:
        inp = inp.to(torch.float32).contiguous()                  # native.py:170
        amax = inp.abs().max().item()                             # :174
        scale = 448.0 / amax if amax > 0 else 1.0                 # :175-176
        scaled = (inp * scale).contiguous()                       # :179
        q = enc_array(scaled.numpy())                             # :183-187 (kernel -> spec)
        inv = torch.tensor([1.0 / scale], dtype=torch.float32)    # :189
        q_inputs.append(inp.numpy())
        q_bytes.append(q)
        q_inv.append(inv.numpy())
    np.savez_compressed(os.path.join(HERE, "host_golden.npz"),
                        deq_scales=deq_scales, deq_out=deq_out,
                        **{f"q{i}_in": a for i, a in enumerate(q_inputs)},
                        **{f"q{i}_bytes": a for i, a in enumerate(q_bytes)},
                        **{f"q{i}_inv": a for i, a in enumerate(q_inv)},
                        n_quant=np.array(len(cases)))

    # ---- known-answer tests quoted in the reference
    kat = {
        "encode": [
            # test_fp8_correctness.py:154-164
            [0.0, 0x00], [0.001953125, 0x01], [0.013671875, 0x07], [0.015625, 0x08],
            [1.0, 0x38], [448.0, 0x7E], [500.0, 0x7E],
            # FIX_DOCUMENTATION.md:33-41, :79-84
            [100.0, 0x6C], [-0.001, 0x80], [0.0186, 0x0A],
        ],
        # test_mps_vs_cpu.py:303 -- kernel bytes must equal torch CPU .to(float8_e4m3fn)
        "torch_cpu_equal": [0.5, 1.0, 2.0, 10.0, 100.0],
        "roundtrip_allow": [0x7F, 0xFF, 0x80],   # test_fp8_correctness.py:118-131 -> 0x00
    }
    for v, b in kat["encode"]:
        assert enc(v) == b, (v, b, enc(v))
    tc = torch.tensor(kat["torch_cpu_equal"]).to(torch.float8_e4m3fn).view(torch.uint8).tolist()
    kat["torch_cpu_bytes"] = tc
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)

    # ---- float8_e5m2 decode: torch CPU cast of all 256 bytes, as fp32 bit patterns (NaN payloads included)
    e5 = torch.arange(256, dtype=torch.int32).to(torch.uint8).view(torch.float8_e5m2)
    np.savez_compressed(os.path.join(HERE, "e5m2_golden.npz"),
                        decode_f32_bits=e5.to(torch.float32).view(torch.int32).numpy().astype(np.uint32),
                        decode_f16_bits=e5.to(torch.float16).view(torch.int16).numpy().view(np.uint16),
                        decode_bf16_bits=e5.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16))
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
