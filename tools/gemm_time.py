"""Times tcgen05 GEMM cases (CUDA events, preallocated outputs):  python tools/gemm_time.py gelu dgelu ...
SWINB200_GEMM_DEBUG: 1 = no staging wait, 2 = no TMA store, 4 = no tmem ld wait (bring-up experiments)."""
import os, sys, json
import torch
sys.path.insert(0, ".")
from swin_v2_weather_b200 import ops
from swin_v2_weather_b200._lib import EPI_ADD_F32, EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU, EPI_F32
T, C, HID = 64800, 768, 3072
bf = lambda *s: (torch.randn(*s, device="cuda") * 0.5).bfloat16()
m = ops.MODE_BF16
res = {"debug": os.environ.get("SWINB200_GEMM_DEBUG", "0")}
for case in sys.argv[1:]:
    if case == "gelu":
        x, w, b = bf(T, C), bf(HID, C), torch.randn(HID, device="cuda")
        o1, o2 = torch.empty(T, HID, device="cuda", dtype=torch.bfloat16), torch.empty(T, HID, device="cuda", dtype=torch.bfloat16)
        f = lambda: ops.gemm(m, x, 0, w, 0, EPI_BIAS_GELU, bias=b, out=o1, out2=o2)
        fl = 2.0 * T * HID * C
    elif case == "fc1bias":
        x, w, b = bf(T, C), bf(HID, C), torch.randn(HID, device="cuda")
        o1 = torch.empty(T, HID, device="cuda", dtype=torch.bfloat16)
        f = lambda: ops.gemm(m, x, 0, w, 0, EPI_BIAS, bias=b, out=o1)
        fl = 2.0 * T * HID * C
    elif case == "dgelu":
        dz, w, h = bf(T, C), bf(C, HID), bf(T, HID)
        o1 = torch.empty(T, HID, device="cuda", dtype=torch.bfloat16)
        f = lambda: ops.gemm(m, dz, 0, w, 1, EPI_DGELU, aux=h, out=o1)
        fl = 2.0 * T * HID * C
    elif case == "fc2dgrad_plain":
        dz, w = bf(T, C), bf(C, HID)
        o1 = torch.empty(T, HID, device="cuda", dtype=torch.bfloat16)
        f = lambda: ops.gemm(m, dz, 0, w, 1, EPI_BIAS, out=o1)
        fl = 2.0 * T * HID * C
    elif case == "addf32":
        dh, w, dxo = bf(T, HID), bf(HID, C), torch.randn(T, C, device="cuda")
        o1 = torch.empty(T, C, device="cuda")
        f = lambda: ops.gemm(m, dh, 0, w, 1, EPI_ADD_F32, aux=dxo, out=o1)
        fl = 2.0 * T * HID * C
    else:
        continue
    for _ in range(3):
        f()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(10):
        f()
    ev[1].record()
    torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[1]) * 100
    res[case] = {"us": round(us, 1), "tflops": round(fl / us / 1e6, 1)}
print(json.dumps(res))
