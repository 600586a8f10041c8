"""Times the tcgen05 GEMM at the model's shapes (CUDA events, L2 flushed by rotating buffers)."""
import sys
import torch
sys.path.insert(0, ".")
from swin_v2_weather_b200 import ops
from swin_v2_weather_b200._lib import BACKEND_TCGEN05, EPI_ADD_F32, EPI_BIAS, EPI_BIAS_GELU, EPI_DGELU, EPI_F32
T, C, HID = 64800, 768, 3072
def t(f, n=20):
    for _ in range(3): f()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize(); ev[0].record()
    for _ in range(n): f()
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n
bf = lambda *s: (torch.randn(*s, device="cuda") * 0.5).bfloat16()
x, w_qkv, w_fc1, w_fc2 = bf(T, C), bf(3 * C, C), bf(HID, C), bf(C, HID)
g, h, dz, dh = bf(T, HID), bf(T, HID), bf(T, C), bf(T, HID)
b3, bh, bc = torch.randn(3 * C, device="cuda"), torch.randn(HID, device="cuda"), torch.randn(C, device="cuda")
dxo = torch.randn(T, C, device="cuda")
m = ops.MODE_BF16
cases = {
 "qkv fwd   (BIAS)     ": (2*T*3*C*C, lambda: ops.gemm(m, x, 0, w_qkv, 0, EPI_BIAS, bias=b3)),
 "qkv fwd   (BIAS_QKNORM)": (2*T*3*C*C, lambda: ops.qkv_projection(m, x, w_qkv, b3, C, 8)),
 "fc1 fwd   (BIAS_GELU)": (2*T*HID*C, lambda: ops.gemm(m, x, 0, w_fc1, 0, EPI_BIAS_GELU, bias=bh)),
 "fc2 fwd   (BIAS)     ": (2*T*HID*C, lambda: ops.gemm(m, g, 0, w_fc2, 0, EPI_BIAS, bias=bc)),
 "fc2 dgrad (DGELU)    ": (2*T*HID*C, lambda: ops.gemm(m, dz, 0, w_fc2, 1, EPI_DGELU, aux=h)),
 "fc1 dgrad (ADD_F32)  ": (2*T*HID*C, lambda: ops.gemm(m, dh, 0, w_fc1, 1, EPI_ADD_F32, aux=dxo)),
 "fc1 wgrad (F32 splitK)": (2*T*HID*C, lambda: ops.gemm(m, dh, 1, x, 1, EPI_F32, out=torch.zeros(HID, C, device="cuda"), accumulate=True, split_k=ops.wgrad_split_k(HID, C, T))),
}
wq = w_qkv.t().contiguous(); w1 = w_fc1.t().contiguous(); w2 = w_fc2.t().contiguous()
cases["cuBLAS x@Wqkv (no epilogue)"] = (2*T*3*C*C, lambda: torch.matmul(x, wq))
cases["cuBLAS x@W1   (no epilogue)"] = (2*T*HID*C, lambda: torch.matmul(x, w1))
cases["cuBLAS g@W2   (no epilogue)"] = (2*T*HID*C, lambda: torch.matmul(g, w2))
for k, (fl, f) in cases.items():
    ms = t(f)
    print(f"{k}: {ms*1e3:7.1f} us  {fl/ms/1e9:7.1f} TFLOP/s")
