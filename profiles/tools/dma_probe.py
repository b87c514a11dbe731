"""torchrun script: would a copy-engine exchange beat the fused TMA push for the N-sharded C4 linear?

Plan under test ("dma"): the shard GEMM runs in row chunks into this rank's column block of its own symmetric
(M, N) buffer; after each chunk one strided device-to-device copy per peer (cudaMemcpy2DAsync on a side stream,
i.e. a copy engine, rows of (N/w)*2 bytes at a pitch of N*2 bytes) forwards the chunk; a symmetric-memory barrier
closes the call.  Timed as a CUDA graph (the plan is 4..60 small host calls per linear) beside the push plan
captured the same way.

   python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P profiles/tools/dma_probe.py"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, os.path.join(ROOT, "fp8-mps-metal_b200"))
import torch, torch.distributed as dist
from cuda.bindings import runtime as cudart

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
from fp8_sharded import ShardedScaledMM, shard_bounds
M, K, N = 4096, 3072, 12288
g = torch.Generator(device=dev).manual_seed(3)
a = torch.randint(0, 120, (M, K), dtype=torch.uint8, device=dev, generator=g)
n0, n1, _ = shard_bounds(N, world, rank)
w = torch.randint(0, 120, (n1 - n0, K), dtype=torch.uint8, device=dev, generator=g)
sa = torch.tensor([0.01], device=dev); sb = torch.tensor([0.02], device=dev)
lin = ShardedScaledMM(w, sb, None, weight_is_shard=True, full_N=N)
lin.fused_barrier = False                                        # both plans close with the same barrier kernel
y_ref = lin(a, sa, torch.bfloat16, mode="push").clone()
key, pair, _ = lin._symm_buffers(M, torch.bfloat16, dev)
side = [torch.cuda.Stream(device=dev) for _ in range(world - 1)]
D2D = cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice


def copy_rows(turn, r0, r1, stream, peer):
    buf, hdl = pair[turn]
    ptrs = [int(p) for p in hdl.buffer_ptrs]
    off = (r0 * N + n0) * 2
    err, = cudart.cudaMemcpy2DAsync(ptrs[peer] + off, N * 2, ptrs[rank] + off, N * 2, (n1 - n0) * 2, r1 - r0, D2D,
                                    stream.cuda_stream)
    assert int(err) == 0, err


def dma_call(turn, nchunk, gemm=True, barrier=True):
    buf, hdl = pair[turn]
    cur = torch.cuda.current_stream()
    for c in range(nchunk):
        r0, r1 = c * M // nchunk, (c + 1) * M // nchunk
        if gemm:
            lin.local(a[r0:r1], sa, torch.bfloat16, out=buf[r0:r1, n0:n1])
        ev = torch.cuda.Event(); ev.record(cur)
        for d in range(1, world):
            s = side[d - 1]
            s.wait_event(ev)
            copy_rows(turn, r0, r1, s, (rank + d) % world)
    for s in side:
        cur.wait_stream(s)
    if barrier:
        hdl.barrier(channel=0)


def timed(fn, name, calls=4, reps=10):
    """fn(turn) captured `calls` times (alternating buffers) into one graph, replayed `reps` times."""
    try:
        for t in range(4): fn(t & 1)
        torch.cuda.synchronize(); dist.barrier()
        gr = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(cap):
            with torch.cuda.graph(gr, stream=cap, capture_error_mode="relaxed"):
                for t in range(calls): fn(t & 1)
        gr.replay(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): gr.replay()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / (reps * calls) * 1e3], device=dev, dtype=torch.float64)
        how = "graph"
    except Exception as e:                                        # capture refused: eager loop (host-bound for many chunks)
        if rank == 0: print(f"  [{name}] graph capture failed ({repr(e)[:100]}), eager timing", flush=True)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(20): fn(t & 1)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 20 * 1e3], device=dev, dtype=torch.float64)
        how = "eager"
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()), 1), how


out = {}
def push_call(turn):
    lin._symm = (key, pair, turn)
    lin(a, sa, torch.bfloat16, mode="push")
out["push"] = timed(push_call, "push")
for nchunk in (1, 2, 4, 8, 16):
    out[f"dma_copy_only_{nchunk}"] = timed(lambda t, n=nchunk: dma_call(t, n, gemm=False), f"copy{nchunk}")
for nchunk in (2, 4, 8, 16):
    out[f"dma_{nchunk}chunks"] = timed(lambda t, n=nchunk: dma_call(t, n), f"dma{nchunk}")
# parity of the dma plan
pair[0][0].zero_(); torch.cuda.synchronize(); dist.barrier()
pair[0][1].barrier(channel=0)
dma_call(0, 4)
torch.cuda.synchronize()
out["dma==push"] = bool(torch.equal(pair[0][0], y_ref))
egress = (world - 1) * M * (n1 - n0) * 2
if rank == 0:
    print(f"world {world}: egress per rank {egress / 1e6:.1f} MB", flush=True)
    for k, v in out.items():
        if isinstance(v, tuple):
            extra = f"  ({egress / v[0] / 1e3:.0f} GB/s egress)" if "copy_only" in k else ""
            print(f"  {k:24s} {v[0]:8.1f} us  [{v[1]}]{extra}", flush=True)
        else:
            print(f"  {k:24s} {v}", flush=True)
dist.barrier()
dist.destroy_process_group()
