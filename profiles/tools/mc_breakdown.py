#!/usr/bin/env python3
"""Break down the fused multicast N-sharded GEMM (torchrun, one rank per GPU): barriers alone, the kernel with
multicast stores alone, the local kernel, and the NCCL all-gather alone."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "fp8-mps-metal_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    import fp8_mps_native
    lib = fp8_mps_native._get_lib()
    M, K, N = 4096, 3072, 12288
    ns = N // world
    g = torch.Generator(device=dev).manual_seed(rank)
    A = torch.randint(0, 120, (M, K), dtype=torch.uint8, device=dev, generator=g)
    W = torch.randint(0, 120, (ns, K), dtype=torch.uint8, device=dev, generator=g)
    one = torch.full((1,), 0.01, device=dev)
    buf = symm_mem.empty((M, N), dtype=torch.bfloat16, device=dev)
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    mc = int(hdl.multicast_ptr)
    local = torch.empty(M, ns, dtype=torch.bfloat16, device=dev)
    gathered = torch.empty(world, M, ns, dtype=torch.bfloat16, device=dev)

    def timeit(fn, n=20):
        for _ in range(5): fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    res = {}
    res["local_gemm"] = timeit(lambda: lib.fp8_scaled_mm_fused(A, W, one, one, None, None, torch.bfloat16, 2, local))
    res["two_barriers"] = timeit(lambda: (hdl.barrier(channel=0), hdl.barrier(channel=1)))
    res["one_barrier"] = timeit(lambda: hdl.barrier(channel=0))
    res["mc_gemm_no_barrier"] = timeit(lambda: lib.fp8_scaled_mm_multicast(A, W, one, one, None, torch.bfloat16, mc, N, rank * ns))
    res["mc_gemm_one_barrier"] = timeit(lambda: (lib.fp8_scaled_mm_multicast(A, W, one, one, None, torch.bfloat16, mc, N, rank * ns), hdl.barrier(channel=0)))
    res["nccl_allgather_only"] = timeit(lambda: dist.all_gather_into_tensor(gathered.view(-1), gathered[rank].reshape(-1)))
    # peer-store variant of the exchange alone: copy the local block into every peer's buffer with plain copies
    peers = [hdl.get_buffer(r, (M, N), torch.bfloat16) for r in range(world)]
    def push():
        for r in range(world):
            peers[r][:, rank * ns:(rank + 1) * ns].copy_(local)
    res["p2p_copy_push_only"] = timeit(push)
    if rank == 0:
        print(f"world {world}: " + "  ".join(f"{k}={v:.1f}us" for k, v in res.items()), flush=True)
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
