import os, sys, torch
os.environ["SWINB200_GEMM_DEBUG"] = "32"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swin_v2_weather_b200 import ops
dev = torch.device("cuda", 0)
M, C = 64800, 768
for K in (768, 3072):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(C, K, device=dev) / K ** 0.5).bfloat16()
    bias, x_in = torch.randn(C, device=dev), torch.randn(M, C, device=dev)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    for _ in range(3):
        z, xo, xb, st = ops.linear_ln_residual(ops.MODE_BF16, a, w, bias, x_in, gamma, beta, None, M, fuse=True)
    torch.cuda.synchronize()
    t = st.view(-1)[:148 * 8].view(148, 8).double()
    print(K, "kernel cyc", t[:, 0].mean().item(), "ln cyc", t[:, 1].mean().item(), "acc-wait cyc", t[:, 2].mean().item(), "slices", t[:, 3].mean().item(),
          "tiles", t[:, 4].mean().item(), "tail cyc", t[:, 5].mean().item(), "cyc/slice", (t[:, 1].sum() / t[:, 3].sum()).item())
