"""One forward + backward of the tiled tcgen05 attention (attn_tc_gen.cu) at a config-5 geometry, for ncu:
   python tools/attn_gen_one.py [Wh Ww heads cpb]   (default 12 24 12 0: 288-token windows, head_dim 64)"""
import sys
import torch
sys.path.insert(0, ".")
from swin_v2_weather_b200 import ops

Wh, Ww, heads, cpb = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (12, 24, 12, 0)))
B, H, W, C = 1, 180, 360, 768
T, L = B * H * W, Wh * Ww
torch.manual_seed(0)
qkv = torch.randn(T, 3 * C, device="cuda").bfloat16()
inv = ops.qk_normalize_(qkv, C, heads)
scale = torch.full((heads,), 10.0, device="cuda")
bias = 0.5 * torch.randn(heads, L, L, device="cuda") if cpb else None
for _ in range(3):
    o, lse = ops.window_attn_fwd(qkv, scale, bias, B, H, W, C, heads, Wh, Ww, Wh // 2, Ww // 2, ops.MODE_BF16)
    d_o = torch.randn_like(o)
    ops.window_attn_bwd(qkv, inv, scale, bias, o, d_o, lse, B, H, W, C, heads, Wh, Ww, Wh // 2, Ww // 2, ops.MODE_BF16)
torch.cuda.synchronize()
print("ok")
