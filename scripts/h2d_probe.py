"""H2D rate of pinned host buffers, default (cacheable) vs write-combined, with every rank copying at once.
   python scripts/h2d_probe.py            (or under torchrun --nproc-per-node N)"""
import os
import time

import torch
import torch.distributed as dist
from cuda import cudart

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1400 * 1000 * 1000  # bytes per copy (half a bench batch)
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
res = {}
for name, flags in (("default", cudart.cudaHostAllocPortable), ("write_combined", cudart.cudaHostAllocPortable | cudart.cudaHostAllocWriteCombined)):
    err, ptr = cudart.cudaHostAlloc(n, flags)
    assert err == cudart.cudaError_t.cudaSuccess, err
    err, st = cudart.cudaStreamCreate()
    # touch the pages by a D2H copy (as bench.py fills its buffers)
    cudart.cudaMemcpyAsync(ptr, dst.data_ptr(), n, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost, st)
    cudart.cudaStreamSynchronize(st)
    for it in range(2):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            cudart.cudaMemcpyAsync(dst.data_ptr(), ptr, n, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, st)
        cudart.cudaStreamSynchronize(st)
        dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[name] = 4 * n / float(t.item()) / 1e9
    cudart.cudaFreeHost(ptr)
if rank == 0:
    print({"world": world, "h2d_gbs_per_rank": res})
