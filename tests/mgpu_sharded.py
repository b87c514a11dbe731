#!/usr/bin/env python3
"""Multi-GPU check of the N-sharded FP8 linear (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mgpu_sharded.py

Every rank compares (a) the NCCL all-gather path, row-major and rank-major layouts, and (b) the fused
multicast-store, peer-store and TMA-push paths against the un-sharded GEMM it computes locally with the same kernel, and a slab of
the result against the CPU oracle.  Prints one PASS/FAIL line per rank; exit code 0 only if all pass."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "fp8-mps-metal_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    import fp8_mps_native
    import fp8_oracle as o
    from fp8_sharded import ShardedScaledMM

    ok = True
    msgs = []
    for (M, K, N, per_row, with_bias) in [(4096, 3072, 12288, False, False), (512, 1024, 2048, True, True),
                                          (300, 512, 1000, True, True)]:
        g = torch.Generator().manual_seed(7)                    # same data on every rank
        A = torch.randint(0, 127, (M, K), dtype=torch.uint8, generator=g)
        W = torch.randint(0, 127, (N, K), dtype=torch.uint8, generator=g)
        sa = torch.tensor([0.01])
        sb = (torch.rand(N, generator=g) * 0.02 + 0.01) if per_row else torch.tensor([0.02])
        bias = torch.randn(N, generator=g).to(torch.bfloat16) if with_bias else None
        Ad, Wd = A.to(dev), W.to(dev)
        sad, sbd = sa.to(dev), sb.to(dev)
        biasd = bias.to(dev) if bias is not None else None
        full = fp8_mps_native.fp8_scaled_mm_fused(Ad, Wd, sad, sbd, bias=biasd, out_dtype=torch.bfloat16)
        lin = ShardedScaledMM(Wd, sbd, biasd)
        y = lin(Ad, sad, out_dtype=torch.bfloat16)
        same = bool(torch.equal(y, full)) and y.is_contiguous() and tuple(y.shape) == (M, N)
        gr = lin(Ad, sad, out_dtype=torch.bfloat16, layout="rank_major")
        relaid = gr.permute(1, 0, 2).reshape(M, -1)[:, :N]
        same_rank = bool(torch.equal(relaid, full))
        mc_state = "skipped"
        if N % (32 * world) == 0 and lin.width == N // world:
            try:
                ymc = lin(Ad, sad, out_dtype=torch.bfloat16, mode="multicast")
                torch.cuda.synchronize()
                mc_ok = bool(torch.equal(ymc, full))
                ymc2 = lin(Ad, sad, out_dtype=torch.bfloat16, mode="multicast")      # buffer reuse
                torch.cuda.synchronize()
                mc_ok = mc_ok and bool(torch.equal(ymc2, full))
                mc_state = "ok" if mc_ok else "MISMATCH"
                ok = ok and mc_ok
            except Exception as e:   # NVLS unavailable is reported, not hidden
                mc_state = f"unavailable: {type(e).__name__}: {str(e)[:120]}"
        pe_state = "skipped"
        if N % (32 * world) == 0 and lin.width == N // world and M > 128 and N // world > 128:
            try:
                ype = lin(Ad, sad, out_dtype=torch.bfloat16, mode="peers")
                torch.cuda.synchronize()
                pe_ok = bool(torch.equal(ype, full))
                ype2 = lin(Ad, sad, out_dtype=torch.bfloat16, mode="peers")          # buffer reuse
                torch.cuda.synchronize()
                pe_ok = pe_ok and bool(torch.equal(ype2, full))
                pe_state = "ok" if pe_ok else "MISMATCH"
                ok = ok and pe_ok
            except Exception as e:
                pe_state = f"FAILED: {type(e).__name__}: {str(e)[:160]}"
                ok = False
        pu_state = "skipped"
        if lin.push_supported(M, K, torch.bfloat16, dev):
            try:
                ypu = lin(Ad, sad, out_dtype=torch.bfloat16, mode="push")
                torch.cuda.synchronize()
                pu_ok = bool(torch.equal(ypu, full))
                ypu2 = lin(Ad, sad, out_dtype=torch.bfloat16, mode="push")           # the other buffer of the pair
                ypu3 = lin(Ad, sad, out_dtype=torch.bfloat16, mode="auto")           # auto = push; buffer reuse
                torch.cuda.synchronize()
                pu_ok = pu_ok and bool(torch.equal(ypu2, full)) and bool(torch.equal(ypu3, full))
                lin.fused_barrier = True                                             # (the default for two ranks only)
                # a burst of back-to-back calls without host synchronisation: the fused barrier (kernel-side signal +
                # PDL wait kernel, growing epochs, two alternating buffers) must keep every result intact
                outs = [lin(Ad, sad, out_dtype=torch.bfloat16, mode="push").clone() for _ in range(12)]
                torch.cuda.synchronize()
                pu_ok = pu_ok and all(bool(torch.equal(y_, full)) for y_ in outs)
                lin.fused_barrier = False                                            # the separate symmetric-memory barrier
                ypu4 = lin(Ad, sad, out_dtype=torch.bfloat16, mode="push")
                torch.cuda.synchronize()
                pu_ok = pu_ok and bool(torch.equal(ypu4, full))
                lin.fused_barrier = world <= 2
                # the call captured in CUDA graphs (one per buffer of the pair), replayed in turn as bench.py does; under
                # capture the call must close with the barrier kernel (the fused barrier's epoch is a host-side count)
                # (no name may drop its last reference to a symmetric buffer INSIDE a capture: freeing one is not a
                # capturable operation -- hence the fresh list per capture and the explicit clean-up afterwards)
                graphs, results = [], []
                for _ in range(2):
                    graphs.append(torch.cuda.CUDAGraph())
                    with torch.cuda.graph(graphs[-1]):
                        results.append(lin(Ad, sad, out_dtype=torch.bfloat16, mode="push"))
                for i in range(6):
                    results[i & 1].zero_()
                    torch.cuda.synchronize()
                    dist.barrier()                                                   # nobody pushes into a buffer still being cleared
                    graphs[i & 1].replay()
                    got = results[i & 1].clone()
                    torch.cuda.synchronize()
                    pu_ok = pu_ok and bool(torch.equal(got, full))
                del graphs, results
                pu_state = "ok" if pu_ok else "MISMATCH"
                ok = ok and pu_ok
            except Exception as e:
                pu_state = f"FAILED: {type(e).__name__}: {str(e)[:160]}"
                ok = False
        else:
            ok = False
            pu_state = "UNSUPPORTED (unexpected for these shapes)"
        rows = slice(0, min(M, 128))
        ref = o.scaled_mm(A[rows].numpy(), W.numpy(), sa.numpy(), sb.numpy(),
                          None if bias is None else bias.float().numpy(), None, "bf16", accum="f32")
        err = o.rel_rmse(y[rows].float().cpu().numpy(), ref)
        case_ok = same and same_rank and err <= 3e-3
        ok = ok and case_ok
        msgs.append(f"M{M} K{K} N{N}: allgather={'ok' if same else 'MISMATCH'} rank_major={'ok' if same_rank else 'MISMATCH'} "
                    f"multicast={mc_state} peers={pe_state} push={pu_state} oracle_rel_rmse={err:.2e}")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    for m in msgs:
        print(f"[rank {rank}/{world}] {m}", flush=True)
    print(f"[rank {rank}/{world}] {'PASS' if ok else 'FAIL'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
