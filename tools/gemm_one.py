"""Runs one tcgen05 GEMM case a few times (ncu target):  python tools/gemm_one.py {qkv|qknorm|gelu|fc2|dgelu|addf32|wgrad}"""
import sys
import torch
sys.path.insert(0, ".")
from swin_v2_weather_b200 import ops
from swin_v2_weather_b200._lib import EPI_ADD_F32, EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU, EPI_F32
T, C, HID = 64800, 768, 3072
bf = lambda *s: (torch.randn(*s, device="cuda") * 0.5).bfloat16()
m = ops.MODE_BF16
case = sys.argv[1]
if case == "qkv":
    x, w, b = bf(T, C), bf(3 * C, C), torch.randn(3 * C, device="cuda")
    f = lambda: ops.gemm(m, x, 0, w, 0, EPI_BIAS, bias=b)
elif case == "qknorm":
    x, w, b = bf(T, C), bf(3 * C, C), torch.randn(3 * C, device="cuda")
    f = lambda: ops.qkv_projection(m, x, w, b, C, 8)
elif case == "gelu":
    x, w, b = bf(T, C), bf(HID, C), torch.randn(HID, device="cuda")
    f = lambda: ops.gemm(m, x, 0, w, 0, EPI_BIAS_GELU, bias=b)
elif case == "fc2":
    g, w, b = bf(T, HID), bf(C, HID), torch.randn(C, device="cuda")
    f = lambda: ops.gemm(m, g, 0, w, 0, EPI_BIAS, bias=b)
elif case == "dgelu":
    dz, w, h = bf(T, C), bf(C, HID), bf(T, HID)
    f = lambda: ops.gemm(m, dz, 0, w, 1, EPI_DGELU, aux=h)
elif case == "addf32":
    dh, w, dxo = bf(T, HID), bf(HID, C), torch.randn(T, C, device="cuda")
    f = lambda: ops.gemm(m, dh, 0, w, 1, EPI_ADD_F32, aux=dxo)
else:
    dh, x = bf(T, HID), bf(T, C)
    out = torch.zeros(HID, C, device="cuda")
    f = lambda: ops.gemm(m, dh, 1, x, 1, EPI_F32, out=out, accumulate=True, split_k=ops.wgrad_split_k(HID, C, T))
for _ in range(4):
    f()
torch.cuda.synchronize()
