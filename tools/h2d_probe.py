"""Host -> device copy rate of one step's inputs (two pinned 73 x 721 x 1440 fp32 fields cropped to 720 rows)."""
import os, sys, time, json
import torch
sys.path.insert(0, ".")
from swin_v2_weather_b200 import distributed as D
from swin_v2_weather_b200.utils.host_io import copy_cropped_async
res = {}
if len(sys.argv) > 1 and sys.argv[1] == "bind":
    res["bound"] = D.bind_to_gpu_numa_node(0)
dev = torch.device("cuda", 0)
host = torch.randn(1, 73, 721, 1440).pin_memory()
hostc = torch.randn(1, 73, 720, 1440).pin_memory()
dst = torch.empty(1, 73, 720, 1440, device=dev)
s = torch.cuda.Stream(dev)
def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n
nbytes = dst.numel() * 4
with torch.cuda.stream(s):
    res["contiguous_GBs"] = round(nbytes / timed(lambda: dst.copy_(hostc, non_blocking=True)) / 1e9, 1)
    res["strided_2d_GBs"] = round(nbytes / timed(lambda: copy_cropped_async(dst, host, s)) / 1e9, 1)
    def planes():
        for c in range(73):
            dst[0, c].copy_(host[0, c, :720], non_blocking=True)
    res["per_plane_GBs"] = round(nbytes / timed(planes) / 1e9, 1)
print(json.dumps(res))
