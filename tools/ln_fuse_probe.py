"""Times swinb200_linear_ln_residual (LayerNorm + residual in the GEMM epilogue) against GEMM + stand-alone LayerNorm.
SWINB200_GEMM_DEBUG bits: 8 = bookkeeping only (no LayerNorm rows), 16 = LayerNorm rows without their stores."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swin_v2_weather_b200 import ops
from swin_v2_weather_b200._lib import EPI_BIAS

dev = torch.device("cuda", 0)
M, C = 64800, 768
out = {"debug": os.environ.get("SWINB200_GEMM_DEBUG", "0")}
for K in (768, 3072):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(C, K, device=dev) / K ** 0.5).bfloat16()
    bias, x_in = torch.randn(C, device=dev), torch.randn(M, C, device=dev)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    def fused():
        return ops.linear_ln_residual(ops.MODE_BF16, a, w, bias, x_in, gamma, beta, None, M, fuse=True)
    def gemm_only():
        return ops.gemm(ops.MODE_BF16, a, 0, w, 0, EPI_BIAS, bias=bias)
    z = gemm_only()
    def ln_only():
        return ops.ln_residual_fwd(z, x_in, gamma, beta, None, None, M, ops.MODE_BF16)
    for name, fn in (("fused", fused), ("gemm", gemm_only), ("ln", ln_only)):
        for _ in range(3):
            fn()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(20):
            fn()
        ev[1].record()
        torch.cuda.synchronize()
        out[f"K{K}_{name}_us"] = round(ev[0].elapsed_time(ev[1]) * 50, 1)
print(json.dumps(out))
