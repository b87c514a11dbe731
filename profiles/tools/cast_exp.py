"""Cast experiments: one launch over N elements (bf16 -> fp8, fp8 -> f16), sizes sweep, vs the batched entry point
over equal-size chunks.  Usage: python profiles/tools/cast_exp.py"""
import ctypes, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "fp8-mps-metal_b200"))
import torch
from _util import capi, make_spans
L = capi(); dev = torch.device("cuda:0")
total = 1 << 32
src = torch.empty(total, dtype=torch.bfloat16, device=dev)
for off in range(0, total, 1 << 28):
    src[off:off + (1 << 28)] = (torch.randn(1 << 28, device=dev) * 0.02).to(torch.bfloat16)
q = torch.empty(total, dtype=torch.uint8, device=dev)
h = torch.empty(total, dtype=torch.float16, device=dev)
sp = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = ctypes.c_void_p

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for n in (1 << 22, 1 << 24, 1 << 26, 1 << 28, 1 << 32):
    cnt = total // n
    def enc_single():
        for i in range(cnt):
            assert L.fp8b_encode(P(src.data_ptr() + 2 * i * n), 2, P(q.data_ptr() + i * n), n, None, sp()) == 0
    def dec_single():
        for i in range(cnt):
            assert L.fp8b_dequant_f16(P(q.data_ptr() + i * n), P(h.data_ptr() + 2 * i * n), n, None, sp()) == 0
    es = make_spans([(src.data_ptr() + 2 * i * n, q.data_ptr() + i * n, n) for i in range(cnt)])
    ds = make_spans([(q.data_ptr() + i * n, h.data_ptr() + 2 * i * n, n) for i in range(cnt)])
    def enc_batch(): assert L.fp8b_encode_batch(es, cnt, 2, sp()) == 0
    def dec_batch(): assert L.fp8b_dequant_batch(ds, cnt, 1, sp()) == 0
    r = []
    for fn in (enc_single, enc_batch, dec_single, dec_batch):
        ms = timeit(fn)
        r.append(3.0 * total / ms / 1e6)
    print(f"n=2^{n.bit_length()-1:2d} x{cnt:4d}: encode single {r[0]:7.0f}  batch {r[1]:7.0f} | dequant single {r[2]:7.0f}  batch {r[3]:7.0f}  GB/s", flush=True)

# launch-shape knob: 0 = library default (by size), 1 = always 256x4 tiles / 4 CTAs per SM, 2 = always big tiles
n = 1 << 28; cnt = total // n
es = make_spans([(src.data_ptr() + 2 * i * n, q.data_ptr() + i * n, n) for i in range(cnt)])
ds = make_spans([(q.data_ptr() + i * n, h.data_ptr() + 2 * i * n, n) for i in range(cnt)])
for shape in (1, 2):
    L.fp8b_set_option(19, shape)                 # FP8B_OPT_TUNE_CAST_SHAPE (round 2: knobs are library options)
    e = 3.0 * total / timeit(lambda: L.fp8b_encode_batch(es, cnt, 2, sp())) / 1e6
    d = 3.0 * total / timeit(lambda: L.fp8b_dequant_batch(ds, cnt, 1, sp())) / 1e6
    print(f"FP8B_CAST_SHAPE={shape}: encode {e:7.0f}  dequant {d:7.0f} GB/s", flush=True)
L.fp8b_set_option(19, -1)

# amax (read-only pass of fp8_quantize): 2 B/element
scales = torch.empty(2, device=dev); scratch = torch.zeros(1, dtype=torch.int32, device=dev)
for lib, tag in ((L, "current"),):
    for cap in (2, 3, 4, 6, 8):
        lib.fp8b_set_option(23, cap)             # FP8B_OPT_TUNE_AMAX_CAP
        ms = timeit(lambda: lib.fp8b_amax_scale(P(src.data_ptr()), 2, total, P(scales.data_ptr()), P(scales.data_ptr() + 4), P(scratch.data_ptr()), sp()))
        print(f"amax bf16 {tag} cap {cap}/SM: {2.0 * total / ms / 1e6:7.0f} GB/s", flush=True)
