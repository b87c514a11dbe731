"""Shape sweep of the large-M path: fp8b_scaled_mm (tcgen05, default tile selection) against stock torch._scaled_mm
(cuBLASLt FP8) on FLUX / DiT-shaped linears and awkward M.  Both run from a CUDA graph over a rotation of operand sets
larger than L2 (or 8 sets, whichever is more), bf16 out, per-tensor scales.  A yardstick, not part of the product.

    python profiles/tools/shape_sweep.py [--cfgs]      # --cfgs: also force every tile configuration (FP8B_OPT_TUNE_GEMM_CFG)
"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, os.path.join(ROOT, "fp8-mps-metal_b200"))
import torch
import fp8_mps_native as nat

dev = torch.device("cuda", 0)
lib = nat._get_lib()
SHAPES = [  # (M, K, N)
    (4096, 3072, 12288), (4096, 12288, 3072), (4096, 3072, 9216), (4096, 3072, 3072), (4096, 15360, 3072),
    (4608, 3072, 21504), (512, 3072, 12288), (1024, 3072, 12288), (2048, 3072, 12288), (8192, 3072, 12288),
    (256, 3072, 3072), (128, 3072, 12288), (64, 3072, 12288), (32, 3072, 12288), (17, 3072, 12288),
    (4096 + 77, 3072, 12288), (1000, 3072, 3072), (333, 4096, 4096), (4096, 4096, 4096), (8192, 8192, 8192),
    (16384, 1024, 1024), (2048, 512, 2048),
]


def timed(fn, sets, reps=5):
    fn(0); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(sets): fn(i)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * sets)


def main():
    cfgs = "--cfgs" in sys.argv
    gen = torch.Generator(device=dev).manual_seed(5)
    print(f"{'M':>6} {'K':>6} {'N':>6} | {'ours us':>9} {'TF/s':>7} | {'cublasLt':>9} {'TF/s':>7} | ours/lib" + ("  | per-cfg us (1..5)" if cfgs else ""))
    shapes = SHAPES
    if os.environ.get("SHAPES"):
        shapes = [tuple(int(v) for v in t.split(",")) for t in os.environ["SHAPES"].split(";")]
    for (M, K, N) in shapes:
        per_set = M * K + N * K + M * N * 2
        sets = max(4, min(16, int(200e6 // per_set) + 1))
        if os.environ.get("SETS"):
            sets = int(os.environ["SETS"])                         # SETS=1: operands L2-resident from the second replay on
        A = [torch.randn(M, K, device=dev, generator=gen) for _ in range(sets)]
        qa = [nat.fp8_quantize(a) for a in A]; del A
        W = [torch.randn(N, K, device=dev, generator=gen) for _ in range(sets)]
        qw = [nat.fp8_quantize(w) for w in W]; del W
        outs = [torch.empty(M, N, dtype=torch.bfloat16, device=dev) for _ in range(sets)]

        def ours(i):
            nat.fp8_scaled_mm_fused(qa[i][0], qw[i][0], qa[i][1], qw[i][1], out_dtype=torch.bfloat16, out=outs[i])

        def stock(i):
            torch._scaled_mm(qa[i][0].view(torch.float8_e4m3fn), qw[i][0].view(torch.float8_e4m3fn).t(), qa[i][1].reshape(()),
                             qw[i][1].reshape(()), None, None, torch.bfloat16, False)
        t_o = timed(ours, sets)
        try:
            t_l = timed(stock, sets)
        except Exception as e:
            t_l = float("nan")
        fl = 2.0 * M * K * N
        line = (f"{M:6d} {K:6d} {N:6d} | {t_o:9.2f} {fl / t_o / 1e6:7.0f} | {t_l:9.2f} {fl / t_l / 1e6:7.0f} | {t_o / t_l:6.3f}")
        if cfgs:
            per = []
            for c in (1, 2, 3, 4, 5):
                lib.set_option(16, c)
                try:
                    per.append(f"{timed(ours, sets):.1f}")
                except Exception:
                    per.append("n/a")
                finally:
                    lib.set_option(16, -1)
            line += "  | " + " ".join(per)
        print(line, flush=True)
        del qa, qw, outs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
