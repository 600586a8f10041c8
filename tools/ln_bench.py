import sys, torch
sys.path.insert(0, ".")
from swin_v2_weather_b200 import ops
T, C = 64800, 768
z = (torch.randn(T, C, device="cuda")).bfloat16(); dx = torch.randn(T, C, device="cuda")
gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
_, _, stats = ops.ln_residual_fwd(z, None, gamma, beta, None, None, T, ops.MODE_BF16)
def t(f, n=20):
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
print("ln_bwd  %.1f us" % t(lambda: ops.ln_residual_bwd(dx, z, stats, gamma, None, T, ops.MODE_BF16)))
x_in = torch.randn(T, C, device="cuda")
print("ln_fwd  %.1f us" % t(lambda: ops.ln_residual_fwd(z, x_in, gamma, beta, None, None, T, ops.MODE_BF16)))
a = torch.empty(T * C * 3 // 2, device="cuda"); b = torch.empty_like(a)
print("copy 398MB r+w %.1f us" % t(lambda: b.copy_(a)))
