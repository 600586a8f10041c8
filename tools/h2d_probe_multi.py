"""Aggregate host -> device copy rate with all ranks copying at once (torchrun --nproc-per-node N):
ordinary pinned memory vs write-combined pinned memory (cudaHostAllocWriteCombined), 605 MB per rank per iteration."""
import ctypes, os, sys, time, json
import torch
import torch.distributed as dist
from cuda.bindings import runtime as cudart

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", init_method="env://")
dev = torch.device("cuda", local)
nbytes = 2 * 73 * 720 * 1440 * 4
dst = torch.empty(nbytes // 4, device=dev)
res = {}
for mode in ("pinned", "write_combined"):
    if mode == "pinned":
        host = torch.empty(nbytes // 4).pin_memory()
        host.normal_()
        hptr = host.data_ptr()
    else:
        err, hptr = cudart.cudaHostAlloc(nbytes, cudart.cudaHostAllocWriteCombined)
        assert int(err) == 0, err
        buf = (ctypes.c_float * (nbytes // 4)).from_address(hptr)
        t = torch.frombuffer(buf, dtype=torch.float32)
        t.copy_(host)                      # the host only ever writes these buffers
    s = torch.cuda.current_stream().cuda_stream
    def once():
        e, = cudart.cudaMemcpyAsync(dst.data_ptr(), hptr, nbytes, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, s)
        assert int(e) == 0
    once(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(20):
        once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rate = torch.tensor([20 * nbytes / dt / 1e9], device=dev)
    if world > 1:
        dist.all_reduce(rate)
    res[mode] = round(float(rate), 1)
if rank == 0:
    print(json.dumps({"world": world, "aggregate_GBs": res}))
