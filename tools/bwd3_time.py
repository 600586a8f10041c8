"""time the attention backward at full geometry (optionally under SWINB200_BWD3_DEBUG experiments)"""
import sys, os, torch
sys.path.insert(0, ".")
from swin_v2_weather_b200 import ops
from swin_v2_weather_b200._lib import BACKEND_TCGEN05
B, H, W, C, heads = 1, 180, 360, 768, 8
T = B * H * W
torch.manual_seed(0)
qkv = torch.randn(T, 3 * C, device="cuda").bfloat16()
inv = ops.qk_normalize_(qkv, C, heads)
scale = torch.full((heads,), 10.0, device="cuda")
for shift in ((0, 0), (4, 9)):
    o, lse = ops.window_attn_fwd(qkv, scale, None, B, H, W, C, heads, 9, 18, shift[0], shift[1], ops.MODE_BF16, backend=BACKEND_TCGEN05)
    d_o = torch.randn_like(o)
    for _ in range(2):
        ops.window_attn_bwd(qkv, inv, scale, None, o, d_o, lse, B, H, W, C, heads, 9, 18, shift[0], shift[1], ops.MODE_BF16, backend=BACKEND_TCGEN05)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(10):
        ops.window_attn_bwd(qkv, inv, scale, None, o, d_o, lse, B, H, W, C, heads, 9, 18, shift[0], shift[1], ops.MODE_BF16, backend=BACKEND_TCGEN05)
    ev[1].record()
    torch.cuda.synchronize()
    print("dbg", os.environ.get("SWINB200_BWD3_DEBUG", "0"), "shift", shift, "bwd %.1f us" % (ev[0].elapsed_time(ev[1]) * 100))
