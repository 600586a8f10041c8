"""bring-up: where does the single-pass attention backward disagree with fp32 autograd?  (per output, per key tile, per column group)"""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_kernels_gpu import oracle_attention, gen
from swin_v2_weather_b200 import ops
from swin_v2_weather_b200._lib import BACKEND_TCGEN05
from oracle import swinv2_oracle as O

B, H, W, C, heads, window, shift = 2, 18, 36, 192, 2, (9, 18), (0, 0)
L, T = 162, B * H * W
raw = gen(T, 3 * C, seed=20).to(torch.bfloat16)
scale = torch.tensor([10.0, 13.5], device="cuda")
qkv = raw.clone()
ops.qk_normalize_(qkv, C, heads)
o, lse = ops.window_attn_fwd(qkv, scale, None, B, H, W, C, heads, 9, 18, 0, 0, ops.MODE_BF16, backend=BACKEND_TCGEN05)
raw_f = raw.float().requires_grad_(True)
sc_f = scale.clone().requires_grad_(True)
o_r, _ = oracle_attention(raw_f, sc_f, None, B, H, W, C, heads, window, shift)
d_o = gen(T, C, seed=22).to(torch.bfloat16)
o_r.backward(d_o.float())
inv = 1.0 / raw.float().view(T, 3, heads, C // heads)[:, :2].norm(dim=-1).clamp_min(1e-12)
dqkv, dscale, _ = ops.window_attn_bwd(qkv, inv.contiguous(), scale, None, o, d_o, lse, B, H, W, C, heads, 9, 18, 0, 0, ops.MODE_BF16, backend=BACKEND_TCGEN05)
idx = O.window_token_index((H, W), window, shift).cuda()      # (nW, L) token of slot n
ref = raw_f.grad
for name, off in (("dq", 0), ("dk", C), ("dv", 2 * C)):
    a = dqkv[:, off:off + C].float().view(B, H * W, heads, 96)[:, idx]      # (B, nW, L, heads, 96)
    b = ref[:, off:off + C].view(B, H * W, heads, 96)[:, idx]
    print(name, "all %.4f" % O.rel_l2(a, b), "slots<128 %.4f" % O.rel_l2(a[:, :, :128], b[:, :, :128]), "slots>=128 %.4f" % O.rel_l2(a[:, :, 128:], b[:, :, 128:]),
          "colgroups", ["%.4f" % O.rel_l2(a[..., 24 * g:24 * g + 24], b[..., 24 * g:24 * g + 24]) for g in range(4)],
          "ratio |ours|/|ref| %.3f" % (a.norm() / b.norm()).item())
print("dscale", dscale.tolist(), sc_f.grad.tolist())
